// fg_tz.h — host side of the zone database the RFC3164 decoder resolves names against
// (time_tz::timezones::get_by_name + assume_timezone, decoder/rfc3164_decoder.rs:196-203).
// fg_tz.cu reads the system's TZif files (RFC 8536) and packs them into the arrays fg::TzDeviceTable describes.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "fg_kernels.cuh"

namespace fg {

// one zone as a caller hands it over: offs.size() == trans.size() + 1, offs[k] in force on UTC seconds [trans[k-1], trans[k])
struct TzZoneSpans {
    std::vector<long long> trans;
    std::vector<int32_t> offs;
};

struct TzHostTable {
    std::vector<unsigned long long> hash;  // identifiers in FNV-1a order
    std::vector<int32_t> zone;
    std::vector<int32_t> name_off;
    std::vector<uint8_t> names;
    std::vector<int32_t> first;
    std::vector<long long> key;
    std::vector<int32_t> off;
    int32_t min_len = 0, max_len = 0;
    uint32_t first_mask[8] = {};
    bool loaded = false;
    int n_names() const { return (int)hash.size(); }
    TzDeviceTable view() const;  // pointers into the vectors above (host-side lookups)
};

// identifiers + their spans -> packed table (links with identical spans share one zone)
void tz_build(const std::vector<std::string>& names, const std::vector<TzZoneSpans>& zones, TzHostTable& out);
// every TZif file below `dir` (NULL: $TZDIR, else /usr/share/zoneinfo); false + err when no zone could be read
bool tz_load_dir(const char* dir, TzHostTable& out, std::string& err);
// one TZif file (version >= 2): explicit transitions + the POSIX TZ footer expanded up to kTzLastYear
bool tz_read_tzif(const std::string& path, TzZoneSpans& out);
constexpr int kTzLastYear = 2400;

}  // namespace fg
