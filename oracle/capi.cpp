// oracle/capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.hpp).
// C entry points over the CPU restatement, loaded by tests/ and bench.py via ctypes.
#include "oracle.hpp"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace fgo;

enum { FMT_RFC5424 = 0, FMT_LTSV = 1, FMT_GELF = 2, FMT_RFC3164 = 3 };

// `cfg` is the configuration object of the format: LtsvConfig for LTSV, Rfc3164Config for RFC3164, unused otherwise
static Decoded decode_one(int fmt, const LtsvConfig* cfg, std::string_view line) {
    static const LtsvConfig empty;
    static const Rfc3164Config empty3164;
    switch (fmt) {
        case FMT_RFC3164: return rfc3164_decode(cfg ? *(const Rfc3164Config*)(const void*)cfg : empty3164, line);
        case FMT_RFC5424: return rfc5424_decode(line);
        case FMT_LTSV: return ltsv_decode(cfg ? *cfg : empty, line);
        default: return gelf_decode(line);
    }
}

extern "C" {

void* fgo_ltsv_config_new() { return new LtsvConfig(); }
void fgo_ltsv_config_free(void* c) { delete (LtsvConfig*)c; }
// type: 0 string 1 bool 2 f64 3 i64 4 u64   (ltsv_decoder.rs:36-43)
void fgo_ltsv_config_add_schema(void* c, const char* name, int type) {
    auto* cfg = (LtsvConfig*)c;
    cfg->has_schema = true;
    cfg->schema[name] = (SDType)type;
}
void fgo_ltsv_config_enable_schema(void* c) { ((LtsvConfig*)c)->has_schema = true; }
void fgo_ltsv_config_set_suffix(void* c, int type, const char* suffix) {  // :68-79
    auto* cfg = (LtsvConfig*)c;
    switch (type) {
        case 1: cfg->suffix_bool = suffix; break;
        case 2: cfg->suffix_f64 = suffix; break;
        case 3: cfg->suffix_i64 = suffix; break;
        case 4: cfg->suffix_u64 = suffix; break;
    }
}

// RFC3164: current year + zone table (oracle.hpp: Rfc3164Config)
void* fgo_rfc3164_config_new(int year) {
    auto* c = new Rfc3164Config();
    c->year = year;
    return c;
}
void fgo_rfc3164_config_free(void* c) { delete (Rfc3164Config*)c; }
// offs has n_trans + 1 entries
void fgo_rfc3164_config_add_zone(void* c, const char* name, int n_trans, const int64_t* trans, const int32_t* offs) {
    TzZone z;
    z.trans.assign(trans, trans + n_trans);
    z.offs.assign(offs, offs + n_trans + 1);
    ((Rfc3164Config*)c)->zones[name] = std::move(z);
}

void fgo_free(void* p) { free(p); }

// Decode n lines (bytes + int32 offsets[n+1]) and return the canonical dumps:
// *out_buf (malloc'd arena) and *out_offsets (malloc'd int64[n+1]).
int fgo_decode_dump(int fmt, const void* cfg, const uint8_t* bytes, const int32_t* offsets, int64_t n,
                    int nthreads, uint8_t** out_buf, int64_t** out_offsets) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::string> parts((size_t)nthreads);
    std::vector<std::vector<int64_t>> lens((size_t)nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            std::string& o = parts[(size_t)t];
            lens[(size_t)t].reserve((size_t)(hi - lo));
            for (int64_t i = lo; i < hi; ++i) {
                size_t before = o.size();
                std::string_view line((const char*)bytes + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
                Decoded d = decode_one(fmt, (const LtsvConfig*)cfg, line);
                dump(d, o);
                lens[(size_t)t].push_back((int64_t)(o.size() - before));
            }
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    uint8_t* buf = (uint8_t*)malloc(total ? total : 1);
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    size_t pos = 0;
    int64_t li = 0;
    offs[0] = 0;
    for (int t = 0; t < nthreads; ++t) {
        memcpy(buf + pos, parts[(size_t)t].data(), parts[(size_t)t].size());
        for (int64_t l : lens[(size_t)t]) {
            offs[li + 1] = offs[li] + l;
            ++li;
        }
        pos += parts[(size_t)t].size();
    }
    *out_buf = buf;
    *out_offsets = offs;
    return 0;
}

// CPU baseline: decode every line into an owned Record (like the reference) on
// nthreads host threads over contiguous line shards; returns wall seconds.
double fgo_decode_bench(int fmt, const void* cfg, const uint8_t* bytes, const int32_t* offsets, int64_t n,
                        int nthreads, int64_t* n_ok) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> ok{0};
    std::atomic<uint64_t> sink{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            int64_t k = 0;
            uint64_t s = 0;
            for (int64_t i = lo; i < hi; ++i) {
                std::string_view line((const char*)bytes + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
                Decoded d = decode_one(fmt, (const LtsvConfig*)cfg, line);
                if (!d.err) {
                    ++k;
                    s += d.rec.hostname.size() + (d.rec.full_msg ? d.rec.full_msg->size() : 0);
                }
            }
            ok += k;
            sink += s;
        });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if (n_ok) *n_ok = ok.load();
    (void)sink.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// single line -> Rust-style rendering: "Ok(Record {...})" or "Err(\"...\")"; malloc'd
char* fgo_decode_debug(int fmt, const void* cfg, const uint8_t* line, int64_t len) {
    Decoded d = decode_one(fmt, (const LtsvConfig*)cfg, std::string_view((const char*)line, (size_t)len));
    std::string s = d.err ? std::string("Err(\"") + d.err + "\")" : "Ok(" + record_debug(d.rec) + ")";
    char* r = (char*)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}

int fgo_parse_f64(const uint8_t* s, int64_t len, double* out) {
    return rust_parse_f64(std::string_view((const char*)s, (size_t)len), *out) ? 1 : 0;
}
int fgo_rfc3339(const uint8_t* s, int64_t len, double* out) {
    return rfc3339_to_unix(std::string_view((const char*)s, (size_t)len), *out) ? 1 : 0;
}
int fgo_english(const uint8_t* s, int64_t len, int with_subsecond, double* out) {
    return english_to_unix(std::string_view((const char*)s, (size_t)len), with_subsecond != 0, *out) ? 1 : 0;
}
// record.rs:94-132 (G15): renders the two fixed structures of the reference's display tests
char* fgo_g15(int which) {
    std::string s;
    StructuredData data;
    data.sd_id = "someid";
    auto mk = [](SDTag t) { SDValue v; v.tag = t; return v; };
    SDValue a = mk(SDTag::String); a.s = "a string";
    SDValue b = mk(SDTag::U64); b.u = 123456;
    SDValue c = mk(SDTag::Bool); c.b = true;
    SDValue dd = mk(SDTag::F64); dd.f = 123.456;
    SDValue e = mk(SDTag::I64); e.i = -123456;
    SDValue f = mk(SDTag::Null);
    data.pairs = {{"a", a}, {"b", b}, {"c", c}, {"d", dd}, {"e", e}, {"_f", f}};
    if (which == 0) s = sd_display(data);
    else if (which == 1) s = sd_debug(data);
    else {
        Record r;
        r.ts = 123.456;
        r.hostname = "hostname";
        r.facility = 3;
        r.severity = 8;
        r.appname = "app";
        r.procid = "123";
        r.msg = "msg";
        s = record_debug(r);
    }
    char* r = (char*)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}

// decode + GelfEncoder::encode (line_splitter.rs:50-52) of n lines: JSON records concatenated into *out_buf with
// *out_offsets[n+1]; a line the decoder rejects contributes an empty record.  extra = output.gelf_extra pairs.
int fgo_decode_encode_gelf(int fmt, const void* cfg, const uint8_t* bytes, const int32_t* offsets, int64_t n, int nthreads,
                           int n_extra, const char* const* extra_keys, const char* const* extra_vals, uint8_t** out_buf,
                           int64_t** out_offsets) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::pair<std::string, std::string>> extra;
    for (int k = 0; k < n_extra; ++k) extra.emplace_back(extra_keys[k], extra_vals[k]);
    std::vector<std::string> parts((size_t)nthreads);
    std::vector<std::vector<int64_t>> lens((size_t)nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            std::string& o = parts[(size_t)t];
            for (int64_t i = lo; i < hi; ++i) {
                const size_t before = o.size();
                Decoded d = decode_one(fmt, (const LtsvConfig*)cfg,
                                       std::string_view((const char*)bytes + offsets[i], (size_t)(offsets[i + 1] - offsets[i])));
                if (!d.err) o += gelf_encode(d.rec, extra);
                lens[(size_t)t].push_back((int64_t)(o.size() - before));
            }
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    uint8_t* buf = (uint8_t*)malloc(total ? total : 1);
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    size_t pos = 0;
    int64_t li = 0;
    offs[0] = 0;
    for (int t = 0; t < nthreads; ++t) {
        memcpy(buf + pos, parts[(size_t)t].data(), parts[(size_t)t].size());
        for (const int64_t l : lens[(size_t)t]) {
            offs[li + 1] = offs[li] + l;
            ++li;
        }
        pos += parts[(size_t)t].size();
    }
    *out_buf = buf;
    *out_offsets = offs;
    return 0;
}

// the four Records of the reference's own encoder tests (gelf_encoder.rs:125-150, :152-173, :175-199, :215-250)
char* fgo_gelf_encoder_test(int which) {
    auto sval = [](const char* s) { SDValue v; v.tag = SDTag::String; v.s = s; return v; };
    Record r;
    r.ts = 1385053862.3072;
    r.severity = 1;
    r.msg = "A short message that helps you identify what is going on";
    std::vector<std::pair<std::string, std::string>> extra;
    if (which == 0 || which == 3) {
        r.hostname = "example.org";
        r.appname = "appname";
        r.procid = "44";
        r.full_msg = "Backtrace here\n\nmore stuff";
        StructuredData sd;
        sd.sd_id = "someid";
        sd.pairs = {{"_some_info", sval("foo")}};
        std::vector<StructuredData> v{sd};
        if (which == 3) {
            StructuredData sd2;
            sd2.sd_id = "someid2";
            SDValue f;
            f.tag = SDTag::F64;
            f.f = 123.456;
            sd2.pairs = {{"info", f}};
            v.push_back(sd2);
        }
        r.sd = v;
        extra = {{"secret-token", "secret"}};
    } else if (which == 2) {
        StructuredData sd;
        sd.pairs = {{"a_key", sval("foo")}};
        r.sd = std::vector<StructuredData>{sd};
        extra = {{"a_key", "bar"}};
    }
    const std::string s = gelf_encode(r, extra);
    char* out = (char*)malloc(s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return out;
}

char* fgo_format_f64(double v) {
    const std::string s = format_f64_json(v);
    char* out = (char*)malloc(s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return out;
}

}  // extern "C"
