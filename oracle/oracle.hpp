// oracle/oracle.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of awslabs/flowgger's Decoder hot path (RFC5424 / LTSV /
// GELF bytes -> Record).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (flowgger_b200/csrc) never links, includes or calls anything in oracle/.
//
// Parity status: the Rust reference cannot be built in this environment (no
// cargo/rustc).  The restatement is pinned against every known-answer test
// the reference holds for this path (G1..G15) in
// tests/test_oracle_golden.py.  Behaviour that lives in un-vendored crates
// (`time` 0.3.x RFC3339/format-description parsing, `serde_json` ~0.8) is
// restated from their published algorithms; corners not covered by the
// reference's own tests are "parity unpinned" and listed in DESIGN.md.
#pragma once
#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace fgo {

// record.rs:4-11
enum class SDTag : uint8_t { String = 0, Bool = 1, F64 = 2, I64 = 3, U64 = 4, Null = 5 };

struct SDValue {
    SDTag tag = SDTag::Null;
    std::string s;
    bool b = false;
    double f = 0.0;
    int64_t i = 0;
    uint64_t u = 0;
};

// record.rs:24-27
struct StructuredData {
    std::optional<std::string> sd_id;
    std::vector<std::pair<std::string, SDValue>> pairs;
};

// record.rs:71-82
struct Record {
    double ts = 0.0;
    std::string hostname;
    std::optional<uint8_t> facility;
    std::optional<uint8_t> severity;
    std::optional<std::string> appname;
    std::optional<std::string> procid;
    std::optional<std::string> msgid;
    std::optional<std::string> msg;
    std::optional<std::string> full_msg;
    std::optional<std::vector<StructuredData>> sd;
};

struct Decoded {
    const char* err = nullptr;  // nullptr => Ok(record); else the reference's &'static str
    Record rec;
    bool ts_is_now = false;                  // gelf_decoder.rs:109 (wall clock) — not comparable
    std::vector<std::string> stdout_lines;   // ltsv_decoder.rs:99-101 println! side effects
};

// ltsv_decoder.rs:24-83 (host-side, once)
enum class SDType : uint8_t { String = 0, Bool = 1, F64 = 2, I64 = 3, U64 = 4 };
struct LtsvConfig {
    bool has_schema = false;
    std::map<std::string, SDType> schema;
    std::optional<std::string> suffix_bool, suffix_f64, suffix_i64, suffix_u64;
};

// rfc3164_decoder.rs: the two inputs the reference takes from its environment, made explicit so a decode is reproducible:
// `OffsetDateTime::now_utc().year()` (:175) and the zone database behind time_tz::timezones::get_by_name (:196).
struct TzZone {
    std::vector<int64_t> trans;  // UTC seconds of the transitions, ascending
    std::vector<int32_t> offs;   // trans.size() + 1 UTC offsets in seconds: offs[k] is in force on [trans[k-1], trans[k])
};
struct Rfc3164Config {
    int32_t year = 1970;
    std::map<std::string, TzZone> zones;
};

Decoded rfc5424_decode(std::string_view line);                       // rfc5424_decoder.rs:18-49
Decoded rfc3164_decode(const Rfc3164Config& cfg, std::string_view line);  // rfc3164_decoder.rs:31-48 (oracle/rfc3164.cpp)
Decoded ltsv_decode(const LtsvConfig& cfg, std::string_view line);   // ltsv_decoder.rs:87-221
Decoded gelf_decode(std::string_view line);                          // gelf_decoder.rs:34-125

// GelfEncoder::encode (encoder/gelf_encoder.rs:59-115) with output.gelf_extra = `extra`; serde_json 0.8 + dtoa text
std::string gelf_encode(const Record& record, const std::vector<std::pair<std::string, std::string>>& extra);
std::string format_f64_json(double v);  // the f64 writer alone

// Canonical, unambiguous text dump used by the parity tests (one line of
// bytes per record; same format is produced independently by the product's
// host materialiser in flowgger_b200/csrc/host/record.cpp).
void dump(const Decoded& d, std::string& out);
// Rust `{:?}` rendering of Record / `{}` of StructuredData (record.rs:42-68,94-132).
std::string record_debug(const Record& r);
std::string sd_display(const StructuredData& sd);
std::string sd_debug(const StructuredData& sd);

// primitives exposed for unit tests
bool rust_parse_f64(std::string_view s, double& out);
bool rust_parse_u64(std::string_view s, uint64_t& out);
bool rust_parse_i64(std::string_view s, int64_t& out);
bool rust_parse_u8(std::string_view s, uint8_t& out);
bool rfc3339_to_unix(std::string_view s, double& out);
bool english_to_unix(std::string_view s, bool with_subsecond, double& out);
std::string_view rust_trim(std::string_view s);
std::string_view rust_trim_end(std::string_view s);

}  // namespace fgo
