// fg_tz.cu — zone database for the RFC3164 decoder: TZif reader + POSIX TZ footer + packing (host code only).
//
// The reference links time-tz, which compiles the IANA database in; `get_by_name` (rfc3164_decoder.rs:196) is an exact
// identifier lookup and `assume_timezone` (:202) takes the offset in force at the LOCAL time.  Here the database comes from
// the system's TZif files (RFC 8536, `zic` output), or from the caller (fg_set_tz_table), and is packed once per context
// into flat arrays the kernel searches (fg_rfc3164.cuh: tz_find, tz_offset_local):
//   identifiers   sorted by FNV-1a 64 hash; first-byte mask and length range reject ordinary hostnames before hashing
//   spans         per zone: key[j] = first LOCAL second span j answers for, off[j] = its UTC offset
// key[j] = transition + max(offset before, offset after): a local time that occurs twice resolves to the earlier
// offset, one that the jump skips to the offset before the jump (PARITY UNPINNED).
#include "fg_tz.h"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>

#include "fg_rfc3164.cuh"

namespace fg {

namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
long long be64(const uint8_t* p) { return (long long)(((unsigned long long)be32(p) << 32) | be32(p + 4)); }

bool leap_year(long long y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }
long long civil_days(long long y, int m, int d) {
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400, yoe = y - era * 400;
    const long long doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    return era * 146097 + yoe * 365 + yoe / 4 - yoe / 100 + doy - 719468;
}

// ---- POSIX TZ string (the TZif footer): std offset [dst [offset] , start[/time] , end[/time]] ---------------------------
struct Rule {
    char kind = 0;  // 'M' month.week.day, 'J' 1..365 without leap day, 'N' 0..365
    int a = 0, b = 0, c = 0;
    long long time = 7200;
};
struct PosixTz {
    long long std_off = 0, dst_off = 0;  // UTC offsets (east positive)
    bool has_rules = false;
    Rule start, end;
};

bool posix_name(const std::string& s, size_t& i) {
    if (i < s.size() && s[i] == '<') {
        const size_t k = s.find('>', i);
        if (k == std::string::npos) return false;
        i = k + 1;
        return true;
    }
    const size_t b = i;
    while (i < s.size() && ((s[i] >= 'A' && s[i] <= 'Z') || (s[i] >= 'a' && s[i] <= 'z'))) ++i;
    return i > b;
}
bool posix_number(const std::string& s, size_t& i, int max_digits, long long& v) {
    const size_t b = i;
    v = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9' && (int)(i - b) < max_digits) v = v * 10 + (s[i++] - '0');
    return i > b;
}
// [+-]hh[:mm[:ss]] in seconds
bool posix_offset(const std::string& s, size_t& i, long long& out) {
    long long sign = 1;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) sign = s[i++] == '-' ? -1 : 1;
    long long h, m = 0, sec = 0;
    if (!posix_number(s, i, 3, h)) return false;
    if (i < s.size() && s[i] == ':') {
        ++i;
        if (!posix_number(s, i, 2, m)) return false;
        if (i < s.size() && s[i] == ':') {
            ++i;
            if (!posix_number(s, i, 2, sec)) return false;
        }
    }
    out = sign * (h * 3600 + m * 60 + sec);
    return true;
}
bool posix_rule(const std::string& s, size_t& i, Rule& r) {
    long long v;
    if (i < s.size() && s[i] == 'M') {
        ++i;
        r.kind = 'M';
        if (!posix_number(s, i, 2, v)) return false;
        r.a = (int)v;
        if (i >= s.size() || s[i++] != '.') return false;
        if (!posix_number(s, i, 1, v)) return false;
        r.b = (int)v;
        if (i >= s.size() || s[i++] != '.') return false;
        if (!posix_number(s, i, 1, v)) return false;
        r.c = (int)v;
        if (r.a < 1 || r.a > 12 || r.b < 1 || r.b > 5 || r.c > 6) return false;
    } else if (i < s.size() && s[i] == 'J') {
        ++i;
        r.kind = 'J';
        if (!posix_number(s, i, 3, v)) return false;
        r.a = (int)v;
    } else {
        r.kind = 'N';
        if (!posix_number(s, i, 3, v)) return false;
        r.a = (int)v;
    }
    r.time = 7200;
    if (i < s.size() && s[i] == '/') {
        ++i;
        if (!posix_offset(s, i, r.time)) return false;
    }
    return true;
}
bool parse_posix_tz(const std::string& s, PosixTz& tz) {
    size_t i = 0;
    long long off;
    if (!posix_name(s, i) || !posix_offset(s, i, off)) return false;
    tz.std_off = -off;  // POSIX offsets count west of Greenwich
    tz.has_rules = false;
    if (i >= s.size()) return true;
    if (!posix_name(s, i)) return false;
    tz.dst_off = tz.std_off + 3600;
    if (i < s.size() && s[i] != ',') {
        if (!posix_offset(s, i, off)) return false;
        tz.dst_off = -off;
    }
    if (i >= s.size()) return true;  // a daylight name without rules: nothing can be generated
    if (s[i++] != ',' || !posix_rule(s, i, tz.start)) return false;
    if (i >= s.size() || s[i++] != ',' || !posix_rule(s, i, tz.end)) return false;
    tz.has_rules = i == s.size();
    return tz.has_rules;
}
// days since the epoch of the rule's date in `year`
long long rule_day(const Rule& r, long long year) {
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    const long long jan1 = civil_days(year, 1, 1);
    if (r.kind == 'J') return jan1 + r.a - 1 + ((leap_year(year) && r.a >= 60) ? 1 : 0);
    if (r.kind == 'N') return jan1 + r.a;
    const long long first = civil_days(year, r.a, 1);
    const int wd_first = (int)(((first + 4) % 7 + 7) % 7);  // 1970-01-01 was a Thursday; 0 = Sunday
    int day = 1 + ((r.c - wd_first) % 7 + 7) % 7 + (r.b - 1) * 7;
    const int md = mdays[r.a - 1] + ((r.a == 2 && leap_year(year)) ? 1 : 0);
    if (day > md) day -= 7;
    return first + day - 1;
}

bool read_file(const std::string& path, std::vector<uint8_t>& data) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t buf[8192];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) {
        data.insert(data.end(), buf, buf + k);
        if (data.size() > (8u << 20)) break;  // no zone file is anywhere near this large
    }
    fclose(f);
    return true;
}

}  // namespace

bool tz_read_tzif(const std::string& path, TzZoneSpans& out) {
    std::vector<uint8_t> d;
    if (!read_file(path, d) || d.size() < 44 || memcmp(d.data(), "TZif", 4) != 0 || d[4] == 0) return false;
    // skip the 32-bit block
    size_t p = 44 + (size_t)be32(&d[32]) * 5 + (size_t)be32(&d[36]) * 6 + be32(&d[40]) + (size_t)be32(&d[28]) * 8 + be32(&d[24]) + be32(&d[20]);
    if (p + 44 > d.size() || memcmp(&d[p], "TZif", 4) != 0) return false;
    const size_t isut = be32(&d[p + 20]), isstd = be32(&d[p + 24]), leap = be32(&d[p + 28]), timecnt = be32(&d[p + 32]),
                 typecnt = be32(&d[p + 36]), charcnt = be32(&d[p + 40]);
    p += 44;
    const size_t need = timecnt * 9 + typecnt * 6 + charcnt + leap * 12 + isstd + isut;
    if (typecnt == 0 || p + need > d.size()) return false;
    const uint8_t* times = &d[p];
    const uint8_t* idx = times + timecnt * 8;
    const uint8_t* types = idx + timecnt;
    auto utoff = [&](size_t k) { return (int32_t)be32(types + 6 * k); };
    out.trans.clear();
    out.offs.clear();
    out.offs.push_back(utoff(0));  // RFC 8536 3.2: before the first transition, time type 0
    for (size_t k = 0; k < timecnt; ++k) {
        if (idx[k] >= typecnt) return false;
        const long long t = be64(times + 8 * k);
        if (!out.trans.empty() && t <= out.trans.back()) return false;
        out.trans.push_back(t);
        out.offs.push_back(utoff(idx[k]));
    }
    p += need;
    if (p < d.size() && d[p] == '\n') {
        size_t e = p + 1;
        while (e < d.size() && d[e] != '\n') ++e;
        const std::string footer((const char*)&d[p + 1], e - (p + 1));
        PosixTz tz;
        if (!footer.empty() && parse_posix_tz(footer, tz) && tz.has_rules) {
            const long long last = out.trans.empty() ? LLONG_MIN : out.trans.back();
            long long y0 = 1900;
            if (!out.trans.empty()) y0 = std::max<long long>(1900, 1970 + (last / 86400) / 366 - 1);
            std::vector<std::pair<long long, int32_t>> extra;
            for (long long y = y0; y <= kTzLastYear; ++y) {
                extra.emplace_back(rule_day(tz.start, y) * 86400 + tz.start.time - tz.std_off, (int32_t)tz.dst_off);  // wall clock before: standard time
                extra.emplace_back(rule_day(tz.end, y) * 86400 + tz.end.time - tz.dst_off, (int32_t)tz.std_off);      // wall clock before: daylight time
            }
            std::sort(extra.begin(), extra.end());
            for (const auto& e2 : extra)
                if (e2.first > (out.trans.empty() ? last : out.trans.back())) {
                    out.trans.push_back(e2.first);
                    out.offs.push_back(e2.second);
                }
        }
    }
    return true;
}

TzDeviceTable TzHostTable::view() const {
    TzDeviceTable T{};
    T.n_names = n_names();
    T.min_len = min_len;
    T.max_len = max_len;
    memcpy(T.first_mask, first_mask, sizeof first_mask);
    T.hash = hash.data();
    T.zone = zone.data();
    T.name_off = name_off.data();
    T.names = names.data();
    T.first = first.data();
    T.key = key.data();
    T.off = off.data();
    return T;
}

void tz_build(const std::vector<std::string>& names, const std::vector<TzZoneSpans>& zones, TzHostTable& out) {
    out = TzHostTable();
    std::map<std::pair<std::vector<long long>, std::vector<int32_t>>, int32_t> seen;  // links share their target's spans
    std::vector<int32_t> zone_of(names.size());
    out.first.push_back(0);
    for (size_t k = 0; k < names.size(); ++k) {
        const TzZoneSpans& z = zones[k];
        auto it = seen.find({z.trans, z.offs});
        if (it != seen.end()) {
            zone_of[k] = it->second;
            continue;
        }
        const int32_t id = (int32_t)out.first.size() - 1;
        seen.emplace(std::make_pair(z.trans, z.offs), id);
        zone_of[k] = id;
        for (size_t j = 0; j < z.offs.size(); ++j) {
            long long key = LLONG_MIN;
            if (j > 0) {
                key = z.trans[j - 1] + std::max(z.offs[j - 1], z.offs[j]);
                key = std::max(key, out.key.back());  // (never needed for zic output: keys are strictly ascending)
            }
            out.key.push_back(key);
            out.off.push_back(z.offs[j]);
        }
        out.first.push_back((int32_t)out.key.size());
    }
    std::vector<size_t> order(names.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::vector<unsigned long long> h(names.size());
    for (size_t k = 0; k < names.size(); ++k) h[k] = tz_fnv1a((const uint8_t*)names[k].data(), 0, (int)names[k].size());
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return h[a] != h[b] ? h[a] < h[b] : names[a] < names[b]; });
    out.min_len = INT32_MAX;
    out.max_len = 0;
    out.name_off.push_back(0);
    for (const size_t k : order) {
        const std::string& n = names[k];
        if (n.empty()) continue;
        out.hash.push_back(h[k]);
        out.zone.push_back(zone_of[k]);
        out.names.insert(out.names.end(), n.begin(), n.end());
        out.name_off.push_back((int32_t)out.names.size());
        out.min_len = std::min<int32_t>(out.min_len, (int32_t)n.size());
        out.max_len = std::max<int32_t>(out.max_len, (int32_t)n.size());
        const uint8_t c0 = (uint8_t)n[0];
        out.first_mask[c0 >> 5] |= 1u << (c0 & 31u);
    }
    if (out.hash.empty()) out.min_len = 0;
    out.loaded = true;
}

namespace {
void walk(const std::string& root, const std::string& rel, std::vector<std::string>& names, std::vector<TzZoneSpans>& zones, int depth) {
    const std::string dir = rel.empty() ? root : root + "/" + rel;
    DIR* d = opendir(dir.c_str());
    if (!d) return;
    std::vector<std::string> entries;
    while (dirent* e = readdir(d)) {
        const std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        entries.push_back(n);
    }
    closedir(d);
    std::sort(entries.begin(), entries.end());
    for (const std::string& n : entries) {
        const std::string r = rel.empty() ? n : rel + "/" + n;
        const std::string path = root + "/" + r;
        struct stat ls, st;
        if (lstat(path.c_str(), &ls) != 0 || stat(path.c_str(), &st) != 0) continue;
        if (S_ISDIR(ls.st_mode)) {  // symlinked directories are not followed
            if (rel.empty() && (n == "posix" || n == "right")) continue;  // the same zones again, under other rules
            if (depth < 4) walk(root, r, names, zones, depth + 1);
            continue;
        }
        if (!S_ISREG(st.st_mode)) continue;
        if (r == "posixrules" || r == "localtime") continue;  // not IANA identifiers
        TzZoneSpans z;
        if (!tz_read_tzif(path, z)) continue;
        names.push_back(r);
        zones.push_back(std::move(z));
    }
}
}  // namespace

bool tz_load_dir(const char* dir, TzHostTable& out, std::string& err) {
    std::string root = dir && *dir ? dir : "";
    if (root.empty()) {
        const char* env = getenv("TZDIR");
        root = env && *env ? env : "/usr/share/zoneinfo";
    }
    std::vector<std::string> names;
    std::vector<TzZoneSpans> zones;
    walk(root, "", names, zones, 0);
    if (names.empty()) {
        err = "no TZif zone files under " + root + " (RFC3164 zone names need a zone database: set fg_config.tzdir or call fg_set_tz_table)";
        return false;
    }
    tz_build(names, zones, out);
    return true;
}

}  // namespace fg
