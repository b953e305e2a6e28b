// gen.cpp — synthetic log generators for tests and bench (NOT product code, NOT the oracle).
// Workload shapes follow SURVEY.md §8(d): C2 RFC5424 (mean 180 B), C3 GELF (mean 512 B),
// C4 LTSV (20 key:value fields).  Every line is a pure function of (seed, line index), so shards can
// be generated in parallel and any sub-range can be regenerated on its own.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed, uint64_t idx) {
        s = seed * 0x9E3779B97F4A7C15ull ^ (idx + 0x632BE59BD9B4E019ull) * 0xD1B54A32D192ED03ull;
        next();
        next();
    }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
    int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }  // inclusive
    bool chance(double p) { return (double)(next() >> 11) * (1.0 / 9007199254740992.0) < p; }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char kHostChars[] = "abcdefghijklmnopqrstuvwxyz0123456789.-";
const char kAlnum[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789";
const char kWords[][16] = {"error", "warning", "request", "user", "failed", "connection", "timeout", "from", "to",
                           "session", "opened", "closed", "for", "id", "status", "GET", "POST", "/api/v1/items",
                           "200", "404", "500", "ms", "bytes", "cache", "miss", "hit", "retry", "queue", "worker",
                           "started", "stopped", "invalid", "token", "auth", "ok", "disk", "usage", "memory"};
const char* kUtf8Bits[] = {"caf\xC3\xA9", "na\xC3\xAFve", "\xE6\x97\xA5\xE6\x9C\xAC\xE8\xAA\x9E", "\xE2\x82\xAC" "42",
                           "\xF0\x9F\x9A\x80", "\xC3\x9Cml\xC3\xA4ut", "\xD0\xBB\xD0\xBE\xD0\xB3"};

void rand_chars(Rng& r, std::string& o, int n, const char* alphabet, int alen) {
    for (int i = 0; i < n; ++i) o.push_back(alphabet[r.below((uint32_t)alen)]);
}

void put2(std::string& o, int v) {
    o.push_back((char)('0' + v / 10));
    o.push_back((char)('0' + v % 10));
}

int dim(int y, int m) {
    static const int d[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    bool leap = (y % 4 == 0 && y % 100 != 0) || y % 400 == 0;
    return (m == 2 && leap) ? 29 : d[m - 1];
}

// RFC3339 timestamp: 2015-2035, fractional digits {0,3,6,9}, Z 70% / +-HH:MM 30%
void rfc3339(Rng& r, std::string& o) {
    int y = r.range(2015, 2035), m = r.range(1, 12), d = r.range(1, dim(y, m));
    char b[8];
    snprintf(b, sizeof b, "%04d", y);
    o += b;
    o.push_back('-');
    put2(o, m);
    o.push_back('-');
    put2(o, d);
    o.push_back('T');
    put2(o, r.range(0, 23));
    o.push_back(':');
    put2(o, r.range(0, 59));
    o.push_back(':');
    put2(o, r.range(0, 59));
    int fd = r.below(4) * 3;
    if (fd) {
        o.push_back('.');
        for (int i = 0; i < fd; ++i) o.push_back((char)('0' + r.below(10)));
    }
    if (r.chance(0.7)) o.push_back('Z');
    else {
        o.push_back(r.chance(0.5) ? '+' : '-');
        put2(o, r.range(0, 14));
        o.push_back(':');
        put2(o, r.below(4) * 15);
    }
}

void message_text(Rng& r, std::string& o, int target, bool utf8) {
    const size_t start = o.size();
    const int nw = (int)(sizeof kWords / sizeof kWords[0]);
    while ((int)(o.size() - start) < target) {
        if (o.size() > start) o.push_back(' ');
        if (utf8 && r.chance(0.25)) o += kUtf8Bits[r.below(7)];
        else if (r.chance(0.15)) rand_chars(r, o, r.range(2, 10), kAlnum, 62);
        else o += kWords[r.below((uint32_t)nw)];
    }
    if ((int)(o.size() - start) > target && !utf8) o.resize(start + (size_t)std::max(target, 1));
    while (o.size() > start && o.back() == ' ') o.pop_back();  // message must not end in a space here
    if (o.size() == start) o.push_back('x');
}

void sd_element(Rng& r, std::string& o) {
    o.push_back('[');
    rand_chars(r, o, r.range(3, 8), kAlnum, 26);
    o.push_back('@');
    rand_chars(r, o, r.range(1, 5), "0123456789", 10);
    int np = r.range(1, 4);
    for (int k = 0; k < np; ++k) {
        o.push_back(' ');
        rand_chars(r, o, r.range(2, 10), kAlnum, 52);
        o += "=\"";
        int vl = r.range(1, 24);
        bool esc = r.chance(0.05);
        for (int i = 0; i < vl; ++i) {
            if (esc && r.chance(0.2)) {
                static const char* e[4] = {"\\\"", "\\\\", "\\]", "\\x"};
                o += e[r.below(4)];
                ++i;
            } else {
                char c = (char)r.range(32, 126);
                if (c == '"' || c == '\\') c = '_';
                o.push_back(c);
            }
        }
        o.push_back('"');
    }
    o.push_back(']');
}

// C2: `<PRI>1 TS HOST APP PROCID MSGID SD MSG`
void gen_rfc5424(uint64_t seed, uint64_t idx, double mean_len, double bad_frac, std::string& o) {
    Rng r(seed, idx);
    const size_t start = o.size();
    int target = (int)std::lround(mean_len + 40.0 * r.normal());
    if (target < 60) target = 60;
    if (target > 1024) target = 1024;
    const bool bad = r.chance(bad_frac);
    const int bad_kind = bad ? (int)r.below(24) : -1;
    if (bad_kind == 0) { o += "\xEF\xBB\xBF"; }               // BOM (valid)
    if (bad_kind == 1) { /* no '<' */ } else o.push_back('<');
    if (bad_kind == 2) o += "256";
    else if (bad_kind == 3) { /* empty pri */ }
    else if (bad_kind == 4) o += "1x";
    else { char b[8]; snprintf(b, sizeof b, "%d", r.range(0, 191)); o += b; }
    if (bad_kind != 5) o.push_back('>');
    o += bad_kind == 6 ? "2" : (bad_kind == 7 ? "" : "1");
    if (bad_kind == 8) return;  // only "<PRI>1"
    o.push_back(' ');
    if (bad_kind == 9) o.push_back('-');                        // NILVALUE timestamp
    else if (bad_kind == 10) o += "2015-13-05T15:53:45Z";       // month 13
    else if (bad_kind == 11) o += "2015-02-30 15:53:45Z";       // no 'T' (adds a space -> shifts fields)
    else if (bad_kind == 12) o += "2017-02-29T00:00:00Z";       // not a leap year
    else rfc3339(r, o);
    int fields = 4;
    if (bad_kind == 13) fields = (int)r.below(4);               // truncated header
    for (int f = 0; f < fields; ++f) {
        o.push_back(' ');
        switch (f) {
            case 0: rand_chars(r, o, r.range(8, 24), kHostChars, 38); break;
            case 1: rand_chars(r, o, r.range(3, 16), kAlnum, 52); break;
            case 2: if (r.chance(0.2)) o.push_back('-'); else rand_chars(r, o, r.range(1, 6), "0123456789", 10); break;
            default: if (r.chance(0.3)) o.push_back('-'); else rand_chars(r, o, r.range(2, 8), kAlnum, 62); break;
        }
    }
    if (bad_kind == 13) return;
    if (bad_kind == 14) return;  // "Missing message data"
    o.push_back(' ');
    if (bad_kind == 15) return;  // "Missing log message"
    const double u = r.uniform();
    const int nsd = u < 0.40 ? 0 : (u < 0.85 ? 1 : 2);
    if (bad_kind == 16) { o += "[id]"; if (r.chance(0.5)) return; }
    else if (bad_kind == 17) o += "[id a=b]";
    else if (bad_kind == 18) { o += "[id a=\"1\""; if (r.chance(0.5)) return; }  // missing ]
    else if (bad_kind == 19) o += "[id =\"v\"]";
    else if (bad_kind == 20) o += "x";                           // Malformated
    else if (nsd == 0) o.push_back('-');
    else for (int k = 0; k < nsd; ++k) sd_element(r, o);
    if (bad_kind == 21) return;                                  // nothing after SD
    if (bad_kind == 22) { o += "junk"; }                         // junk right after ']' (or after '-')
    o.push_back(' ');
    if (bad_kind == 23) { o += "  "; return; }                   // whitespace-only message -> msg None
    const bool utf8 = r.chance(0.02);
    int room = target - (int)(o.size() - start);
    message_text(r, o, room < 1 ? 1 : room, utf8);
    if (r.chance(0.01)) o.append((size_t)r.range(1, 3), ' ');
}


// C4: 20 tab-separated key:value fields — time (1/3 decimal unix, 1/3 [RFC3339], 1/3 [D/Mon/YYYY:HH:MM:SS(.f) +-HHMM]),
// host, message, level and 16 free pairs (4 of them are the typed keys counter/score/mean/done so that the same
// lines exercise a u64/i64/f64/bool schema); keys 2-12 B, values 0-24 B and may contain ':'.
void gen_ltsv(uint64_t seed, uint64_t idx, double /*mean_len*/, double bad_frac, std::string& o) {
    static const char* MON[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
    Rng r(seed, idx);
    const size_t line_start = o.size();
    const bool bad = r.chance(bad_frac);
    const int bad_kind = bad ? (int)r.below(12) : -1;
    // field order is shuffled: position of the 4 special keys among 20 slots
    int order[20];
    for (int k = 0; k < 20; ++k) order[k] = k;
    for (int k = 19; k > 0; --k) { int j = (int)r.below((uint32_t)k + 1); int t = order[k]; order[k] = order[j]; order[j] = t; }
    char buf[64];
    for (int slot = 0; slot < 20; ++slot) {
        if (slot) o.push_back('\t');
        const int f = order[slot];
        if (f == 0) {  // time
            if (bad_kind == 0) { o += "tim:1"; continue; }
            o += "time:";
            if (bad_kind == 1) { o += "yesterday"; continue; }
            const int form = (int)r.below(3);
            if (form == 0) {
                snprintf(buf, sizeof buf, "%u", (unsigned)r.range(1420070400, 2051222400));
                o += buf;
                int fd = (int)r.below(4) * 3;
                if (fd) { o.push_back('.'); for (int i = 0; i < fd; ++i) o.push_back((char)('0' + r.below(10))); }
            } else if (form == 1) {
                o.push_back('[');
                rfc3339(r, o);
                o.push_back(']');
            } else {
                int y = r.range(2015, 2035), m = r.range(1, 12), d = r.range(1, dim(y, m));
                snprintf(buf, sizeof buf, "[%d/%s/%04d:%02d:%02d:%02d", d, MON[m - 1], y, r.range(0, 23), r.range(0, 59), r.range(0, 59));
                o += buf;
                if (r.chance(0.5)) { o.push_back('.'); int fd = r.range(1, 9); for (int i = 0; i < fd; ++i) o.push_back((char)('0' + r.below(10))); }
                snprintf(buf, sizeof buf, " %c%02d%02d]", r.chance(0.5) ? '+' : '-', r.range(0, 14), (int)r.below(4) * 15);
                o += buf;
            }
        } else if (f == 1) {  // host
            if (bad_kind == 2) { o += "hostname:x"; continue; }
            o += "host:";
            rand_chars(r, o, r.range(8, 24), kHostChars, 38);
        } else if (f == 2) {  // message
            o += "message:";
            message_text(r, o, r.range(10, 60), r.chance(0.02));
        } else if (f == 3) {  // level
            o += "level:";
            if (bad_kind == 3) o += "9";
            else if (bad_kind == 4) o += "high";
            else o.push_back((char)('0' + r.below(8)));
        } else if (f == 4) {
            o += "counter:";
            if (bad_kind == 5) o += "-5";
            else { snprintf(buf, sizeof buf, "%llu", (unsigned long long)(r.next() >> (r.below(60)))); o += buf; }
        } else if (f == 5) {
            o += "score:";
            if (bad_kind == 6) o += "1.5";
            else { snprintf(buf, sizeof buf, "%lld", (long long)(r.next() >> (1 + r.below(60))) * (r.chance(0.5) ? -1 : 1)); o += buf; }
        } else if (f == 6) {
            o += "mean:";
            if (bad_kind == 7) o += "n/a";
            else {
                const int form = (int)r.below(4);
                if (form == 0) snprintf(buf, sizeof buf, "%.*f", r.range(0, 9), r.uniform() * 1000.0);
                else if (form == 1) snprintf(buf, sizeof buf, "%.17g", (r.uniform() - 0.5) * std::pow(10.0, r.range(-30, 30)));
                else if (form == 2) snprintf(buf, sizeof buf, "%ue%d", (unsigned)r.below(100000), r.range(-40, 40));
                else snprintf(buf, sizeof buf, "%llu.%llu", (unsigned long long)(r.next() >> 10), (unsigned long long)(r.next() >> 20));
                o += buf;
            }
        } else if (f == 7) {
            o += "done:";
            if (bad_kind == 8) o += "TRUE";
            else o += r.chance(0.5) ? "true" : "false";
        } else {
            if (bad_kind == 9 && f == 8) { rand_chars(r, o, r.range(2, 12), kAlnum, 52); continue; }  // no ':' -> println!
            rand_chars(r, o, r.range(2, 12), kAlnum, 52);
            o.push_back(':');
            if (bad_kind == 10 && f == 9) continue;  // empty value
            const int vl = r.range(0, 24);
            const bool utf8 = r.chance(0.01);
            for (int i = 0; i < vl; ++i) {
                if (utf8 && r.chance(0.2)) { o += kUtf8Bits[r.below(7)]; continue; }
                char c = (char)r.range(32, 126);
                o.push_back(c);
            }
        }
    }
    if (bad_kind == 11) o.resize(line_start);  // empty line
}


void json_text(Rng& r, std::string& o, int target, bool rich) {
    // message text with JSON escapes: \n \t \" \\ and (1% of strings) \uXXXX incl. surrogate pairs
    const size_t start = o.size();
    const int nw = (int)(sizeof kWords / sizeof kWords[0]);
    const bool uni = r.chance(0.01);
    while ((int)(o.size() - start) < target) {
        if (o.size() > start) {
            const double u = r.uniform();
            if (rich && u < 0.06) o += "\\n";
            else if (rich && u < 0.09) o += "\\t";
            else if (rich && u < 0.11) o += "\\\"";
            else if (rich && u < 0.12) o += "\\\\";
            else if (uni && u < 0.20) o += r.chance(0.5) ? "\\u00e9" : "\\ud83d\\ude80";
            else o.push_back(' ');
        }
        if (r.chance(0.03)) o += kUtf8Bits[r.below(7)];
        else if (r.chance(0.15)) rand_chars(r, o, r.range(2, 10), kAlnum, 62);
        else o += kWords[r.below((uint32_t)nw)];
    }
}

// C3: GELF objects — version, host, short_message (40-120 B), full_message (150-300 B, escape rich), timestamp,
// level, 3-8 `_extra` members (strings 60 %, ints 25 %, floats <= 15 significant digits 10 %, bool/null 5 %),
// member order shuffled, random inter-token spaces; 0.5 % malformed.
void gen_gelf(uint64_t seed, uint64_t idx, double mean_len, double bad_frac, std::string& o) {
    Rng r(seed, idx);
    const double k = mean_len > 0 ? mean_len / 577.0 : 1.0;  // scales the message lengths (577 B = unscaled mean)
    const size_t line_start = o.size();
    const bool bad = r.chance(bad_frac);
    const int bad_kind = bad ? (int)r.below(16) : -1;
    const int nextra = r.range(3, 8);
    const int nmem = 6 + nextra;
    int order[16];
    for (int k = 0; k < nmem; ++k) order[k] = k;
    for (int k = nmem - 1; k > 0; --k) { int j = (int)r.below((uint32_t)k + 1); int t = order[k]; order[k] = order[j]; order[j] = t; }
    auto sp = [&]() { if (r.chance(0.3)) o.append((size_t)r.range(1, 2), ' '); };
    char buf[64];
    o.push_back('{');
    bool first = true;
    for (int slot = 0; slot < nmem; ++slot) {
        const int f = order[slot];
        if (f == 4 && bad_kind == 0) continue;  // no timestamp (valid: wall clock)
        if (f == 1 && bad_kind == 1) continue;  // Missing hostname
        if (!first) { o.push_back(','); }
        first = false;
        sp();
        if (f == 0) {
            o += "\"version\":"; sp();
            o += bad_kind == 2 ? "\"2.0\"" : (bad_kind == 3 ? "1.1" : (r.chance(0.5) ? "\"1.1\"" : "\"1.0\""));
        } else if (f == 1) {
            o += "\"host\":"; sp();
            if (bad_kind == 4) o += "42";
            else { o.push_back('"'); rand_chars(r, o, r.range(8, 24), kHostChars, 38); o.push_back('"'); }
        } else if (f == 2) {
            o += "\"short_message\":"; sp();
            if (bad_kind == 5) o += "null";
            else { o.push_back('"'); json_text(r, o, (int)(k * r.range(40, 120)), true); o.push_back('"'); }
        } else if (f == 3) {
            o += "\"full_message\":"; sp();
            if (bad_kind == 6) o += "[\"nested\",{\"a\":1}]";
            else {
                o.push_back('"');
                json_text(r, o, (int)(k * r.range(150, 300)), true);
                if (bad_kind == 7) o.push_back('\n');       // raw LF inside a string: newline retry path
                if (bad_kind == 8) o.push_back('\t');       // raw TAB inside a string: error
                o.push_back('"');
            }
        } else if (f == 4) {
            o += "\"timestamp\":"; sp();
            if (bad_kind == 9) o += "\"yesterday\"";
            else {
                snprintf(buf, sizeof buf, "%u", (unsigned)r.range(1420070400, 2051222400));
                o += buf;
                const int fd = r.range(0, 6);
                if (fd) { o.push_back('.'); for (int i = 0; i < fd; ++i) o.push_back((char)('0' + r.below(10))); }
            }
        } else if (f == 5) {
            o += "\"level\":"; sp();
            if (bad_kind == 10) o += "8";
            else if (bad_kind == 11) o += "-1";
            else o.push_back((char)('0' + r.below(8)));
        } else {
            o += "\"_";
            rand_chars(r, o, r.range(3, 12), kAlnum, 52);
            o += "\":"; sp();
            const double u = r.uniform();
            if (bad_kind == 12 && f == 6) o += "{\"nested\":{\"deep\":[1,2,3]}}";
            else if (u < 0.60) { o.push_back('"'); json_text(r, o, r.range(2, 14), r.chance(0.2)); o.push_back('"'); }
            else if (u < 0.85) { snprintf(buf, sizeof buf, "%lld", (long long)(r.next() >> (4 + r.below(56))) * (r.chance(0.2) ? -1 : 1)); o += buf; }
            else if (u < 0.95) {
                const int form = (int)r.below(3);
                if (form == 0) snprintf(buf, sizeof buf, "%.*f", r.range(1, 6), r.uniform() * 10000.0);
                else if (form == 1) snprintf(buf, sizeof buf, "%.15g", (r.uniform() - 0.5) * std::pow(10.0, r.range(-20, 20)));
                else snprintf(buf, sizeof buf, "%ue%d", (unsigned)r.below(100000), r.range(-22, 22));
                std::string t = buf;
                if (t.find_first_of(".e") == std::string::npos) t += ".0";
                if (t.find("inf") != std::string::npos || t.find("nan") != std::string::npos) t = "1.5";
                o += t;
            } else { o += r.chance(0.4) ? "true" : (r.chance(0.5) ? "false" : "null"); }
        }
        sp();
    }
    if (bad_kind == 13) o += ",";          // trailing comma
    o.push_back('}');
    if (bad_kind == 14) o += " x";         // trailing characters
    if (bad_kind == 15) { o.resize(line_start); o += "[1,2,3]"; }  // not an object
}



// RFC3164 (BSD syslog, decoder/rfc3164_decoder.rs): `[<PRI>][YYYY ]Mon D HH:MM:SS [Zone ]host tag: text` and the custom
// form `[<PRI>]host: [YYYY ]Mon D HH:MM:SS[ Zone]: text`; mean ~140 B.  Irregular spacing (double spaces, TABs, Unicode
// White_Space), non-ASCII text and trailing blanks appear at a few per cent each so that the re-join path is exercised.
void gen_rfc3164(uint64_t seed, uint64_t idx, double mean_len, double bad_frac, std::string& o) {
    Rng r(seed ^ 0x3164ull, idx);
    static const char* MON[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
    static const char* ZONES[] = {"UTC", "Europe/Paris", "America/Sao_Paulo", "America/New_York", "Asia/Tokyo", "Etc/GMT+5",
                                  "Australia/Lord_Howe", "GMT", "Africa/Casablanca", "America/Argentina/Buenos_Aires", "EST5EDT",
                                  "Europe/Dublin", "Pacific/Chatham", "Asia/Kolkata"};
    const bool bad = r.chance(bad_frac);
    const int bad_kind = bad ? (int)r.below(12) : -1;
    const bool custom = r.chance(0.08);
    const bool with_year = custom ? r.chance(0.8) : r.chance(0.12);
    const bool with_zone = r.chance(0.10);
    if (r.chance(0.9)) {
        o.push_back('<');
        if (bad_kind == 0) o += "256";
        else if (bad_kind == 1) o += "1x";
        else o += std::to_string(r.below(192));
        if (bad_kind != 2) o.push_back('>');
    }
    std::string host;
    rand_chars(r, host, r.range(4, 20), kHostChars, 38);
    if (host[0] == '.' || host[0] == '-') host[0] = 'h';
    std::string date;
    {
        const int y = r.range(1971, 2099), m = r.range(1, 12);
        int d = r.range(1, dim(with_year ? y : 2023, m));  // a day every year has (without a year the current one decides)
        if (with_year) {
            char b[8];
            snprintf(b, sizeof b, "%04d", y);
            date += b;
            date.push_back(' ');
        }
        date += bad_kind == 3 ? "aug" : (bad_kind == 4 ? "Sept" : MON[m - 1]);
        date.push_back(' ');
        if (!custom && d < 10 && r.chance(0.7)) date.push_back(' ');  // the BSD form pads the day with a space
        if (bad_kind == 5) d = r.chance(0.5) ? 0 : 32;
        date += std::to_string(d);
        date.push_back(' ');
        put2(date, bad_kind == 6 ? 24 : r.range(0, 23));
        date.push_back(':');
        put2(date, r.range(0, 59));
        if (bad_kind != 7) {
            date.push_back(':');
            put2(date, bad_kind == 8 ? 60 : r.range(0, 59));
        }
        if (with_zone) {
            date.push_back(' ');
            date += r.chance(0.06) ? "Mars/Phobos" : ZONES[r.below((uint32_t)(sizeof ZONES / sizeof ZONES[0]))];
        }
    }
    const int target = std::max(8, (int)(mean_len - 62.0 + r.normal() * 30.0));
    if (custom) {
        o += host;
        o += bad_kind == 9 ? ":" : ": ";
        o += date;
        o += bad_kind == 10 ? " " : ": ";
    } else {
        o += date;
        if (bad_kind == 11) return;  // the date alone: too few tokens, or a zone that swallows the last one
        o.push_back(' ');
        o += host;
        o.push_back(' ');
    }
    // tag[pid]: text
    rand_chars(r, o, r.range(3, 12), kAlnum, 52);
    if (r.chance(0.6)) {
        o.push_back('[');
        o += std::to_string(r.below(65536));
        o.push_back(']');
    }
    o += ": ";
    const bool utf8 = r.chance(0.04);
    const size_t m0 = o.size();
    message_text(r, o, target, utf8);
    const int style = (int)r.below(100);
    if (style < 6) {  // irregular spacing inside the text
        static const char* WS[] = {"  ", "\t", " \t ", "\xC2\xA0", "\xE3\x80\x80", "   ", "\xE2\x80\xA8", " \xC2\x85"};
        for (size_t q = m0; q < o.size(); ++q)
            if (o[q] == ' ' && r.chance(0.3)) {
                const char* w = WS[r.below(8)];
                o.replace(q, 1, w);
                q += strlen(w) - 1;
            }
    } else if (style < 10) {
        o += r.chance(0.5) ? " " : "  \t";  // trailing blanks: trimmed off full_msg, dropped by the re-join
    } else if (style < 12) {
        o += " \x01\x1b[0m x";  // control bytes that are not White_Space
    }
}

}  // namespace

extern "C" {

void fgen_free(void* p) { free(p); }

// kind 0 = RFC5424 (C2).  Returns malloc'd bytes + int32 offsets[n+1]; fails (-1) past 2 GiB.
static int g_terminate = 0;  // 1: every line is followed by '\n' (raw stream for fg_split_decode); offsets then include it
void fgen_set_terminator(int on) { g_terminate = on; }

int fgen_generate(int kind, uint64_t seed, int64_t first_index, int64_t n, double mean_len, double bad_frac,
                  int nthreads, uint8_t** out_bytes, int32_t** out_offsets, int64_t* out_total) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::string> parts((size_t)nthreads);
    std::vector<std::vector<int32_t>> lens((size_t)nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            std::string& o = parts[(size_t)t];
            o.reserve((size_t)((hi - lo) * (int64_t)(mean_len + 16)));
            lens[(size_t)t].reserve((size_t)(hi - lo));
            for (int64_t i = lo; i < hi; ++i) {
                const size_t before = o.size();
                switch (kind) {
                    case 3: gen_rfc3164(seed, (uint64_t)(first_index + i), mean_len, bad_frac, o); break;
                    case 2: gen_gelf(seed, (uint64_t)(first_index + i), mean_len, bad_frac, o); break;
                    case 1: gen_ltsv(seed, (uint64_t)(first_index + i), mean_len, bad_frac, o); break;
                    default: gen_rfc5424(seed, (uint64_t)(first_index + i), mean_len, bad_frac, o); break;
                }
                if (g_terminate) {
                    for (size_t q = before; q < o.size(); ++q) if (o[q] == '\n') o[q] = ' ';  // a raw LF would be a line break
                    o.push_back('\n');
                }
                lens[(size_t)t].push_back((int32_t)(o.size() - before));
            }
        });
    }
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    if (total > 0x7FFFFFC0ull) return -1;
    uint8_t* buf = (uint8_t*)malloc(total + 64);
    int32_t* offs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    size_t pos = 0;
    int64_t li = 0;
    offs[0] = 0;
    for (int t = 0; t < nthreads; ++t) {
        memcpy(buf + pos, parts[(size_t)t].data(), parts[(size_t)t].size());
        for (int32_t l : lens[(size_t)t]) {
            offs[li + 1] = offs[li] + l;
            ++li;
        }
        pos += parts[(size_t)t].size();
    }
    *out_bytes = buf;
    *out_offsets = offs;
    *out_total = (int64_t)total;
    return 0;
}

}  // extern "C"
