// prep.cpp -- host-only pieces of the Grid-B data preparation: what the
// reference does in Python before its drivers start (code/subset_bathymetry.py) restated in C++ so that a
// caller can go from a GEBCO download to a masked grid on the GPU without pandas, CSV text or a
// vector<vector<double>>:
//
//   auvi_netcdf3_find    locate a variable in a NetCDF-3 classic file image (GEBCO tiles are CDF-1:
//                        subset_bathymetry.py:8-14 reads 'lat', 'lon', 'elevation' through netCDF4)
//   auvi_netcdf3_read_f64  decode a (small) variable to doubles -- the coordinate axes
//   auvi_legacy_choice   numpy.random.seed(s); numpy.random.choice(total, n, replace=False)
//                        (subset_bathymetry.py:32-39), i.e. the legacy MT19937 permutation prefix
//
// Nothing here touches the GPU; the device side (decode + row flip, masking, masked metrics) is ingest.cu.
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/auvi.h"

namespace auvi { int set_error(const std::string& msg); }

namespace {

// ---- NetCDF-3 classic / 64-bit-offset header (network byte order) ------------------------------------------
struct Reader {
    const unsigned char* p;
    int64_t n, at;
    bool ok;
    uint32_t u32() {
        if (at + 4 > n) { ok = false; return 0; }
        const uint32_t v = (uint32_t(p[at]) << 24) | (uint32_t(p[at + 1]) << 16) | (uint32_t(p[at + 2]) << 8) | p[at + 3];
        at += 4;
        return v;
    }
    uint64_t u64() { const uint64_t hi = u32(); return (hi << 32) | u32(); }
    std::string name() {
        const uint32_t len = u32();
        if (!ok || at + len > n) { ok = false; return std::string(); }
        std::string s(reinterpret_cast<const char*>(p + at), len);
        at += (len + 3) / 4 * 4;
        return s;
    }
    void skip(int64_t bytes) { at += (bytes + 3) / 4 * 4; if (at > n) ok = false; }
};

int type_size(int t) { return t == 1 || t == 2 ? 1 : t == 3 ? 2 : t == 4 || t == 5 ? 4 : t == 6 ? 8 : 0; }

struct Attr { std::string name; int type; int64_t nelems; int64_t at; };

bool read_attrs(Reader& r, std::vector<Attr>* out) {
    const uint32_t tag = r.u32(), cnt = r.u32();
    if (!r.ok || (tag != 0 && tag != 0x0C) || (tag == 0 && cnt != 0)) return false;
    if (static_cast<int64_t>(cnt) * 12 > r.n - r.at) return false;   // an attribute takes at least 12 header bytes
    for (uint32_t k = 0; k < cnt; ++k) {
        Attr a;
        a.name = r.name();
        a.type = static_cast<int>(r.u32());
        a.nelems = r.u32();
        a.at = r.at;
        const int ts = type_size(a.type);
        if (!r.ok || ts == 0 || a.nelems * ts > r.n - r.at) return false;
        r.skip(a.nelems * ts);
        if (!r.ok) return false;
        if (out) out->push_back(a);
    }
    return true;
}

double decode_be(const unsigned char* q, int type) {
    switch (type) {
        case 1: return static_cast<double>(static_cast<int8_t>(q[0]));
        case 2: return static_cast<double>(q[0]);
        case 3: return static_cast<double>(static_cast<int16_t>((q[0] << 8) | q[1]));
        case 4: return static_cast<double>(static_cast<int32_t>((uint32_t(q[0]) << 24) | (uint32_t(q[1]) << 16) | (uint32_t(q[2]) << 8) | q[3]));
        case 5: { uint32_t u = (uint32_t(q[0]) << 24) | (uint32_t(q[1]) << 16) | (uint32_t(q[2]) << 8) | q[3]; float f; std::memcpy(&f, &u, 4); return f; }
        default: { uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | q[k]; double d; std::memcpy(&d, &u, 8); return d; }
    }
}

// ---- MT19937 exactly as numpy's legacy RandomState seeds and steps it --------------------------------------
struct MT19937 {
    uint32_t key[624];
    int pos;
    explicit MT19937(uint32_t seed) {                              // init_genrand (Knuth), numpy _legacy_seeding for ints
        key[0] = seed;
        for (int i = 1; i < 624; ++i) key[i] = 1812433253u * (key[i - 1] ^ (key[i - 1] >> 30)) + static_cast<uint32_t>(i);
        pos = 624;
    }
    void refill() {
        const uint32_t upper = 0x80000000u, lower = 0x7fffffffu, matrix = 0x9908b0dfu;
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (key[i] & upper) | (key[(i + 1) % 624] & lower);
            key[i] = key[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? matrix : 0u);
        }
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) refill();
        uint32_t y = key[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    uint64_t next64() { const uint64_t hi = next32(); return (hi << 32) | next32(); }
    // legacy random_interval: masked rejection sampling, 32-bit draws while max fits in 32 bits
    uint64_t interval(uint64_t max) {
        if (max == 0) return 0;
        uint64_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        uint64_t v;
        if (max <= 0xffffffffull) { while ((v = (next32() & mask)) > max) {} }
        else { while ((v = (next64() & mask)) > max) {} }
        return v;
    }
};

}  // namespace

extern "C" {

int auvi_netcdf3_find(const void* file_image, int64_t n_bytes, const char* var_name, auvi_nc_var* out) {
    if (!file_image || !var_name || !out) return auvi::set_error("null argument");
    std::memset(out, 0, sizeof *out);
    Reader r{static_cast<const unsigned char*>(file_image), n_bytes, 0, true};
    if (n_bytes < 8 || r.p[0] != 'C' || r.p[1] != 'D' || r.p[2] != 'F' || (r.p[3] != 1 && r.p[3] != 2))
        return auvi::set_error("not a NetCDF-3 classic / 64-bit-offset file (magic CDF\\x01 or CDF\\x02 expected)");
    const int version = r.p[3];
    r.at = 4;
    r.u32();                                                      // numrecs
    std::vector<int64_t> dim_len;
    {
        const uint32_t tag = r.u32(), cnt = r.u32();
        if (!r.ok || (tag != 0 && tag != 0x0A)) return auvi::set_error("corrupt NetCDF header (dimension list)");
        // counts come from the (untrusted) file: a dimension entry takes at least 8 header bytes
        if (static_cast<int64_t>(cnt) * 8 > r.n - r.at) return auvi::set_error("corrupt NetCDF header (dimension count)");
        for (uint32_t k = 0; k < cnt && r.ok; ++k) { r.name(); dim_len.push_back(r.u32()); }
    }
    if (!r.ok || !read_attrs(r, nullptr)) return auvi::set_error("corrupt NetCDF header (global attributes)");
    const uint32_t tag = r.u32(), n_vars = r.u32();
    if (!r.ok || (tag != 0 && tag != 0x0B)) return auvi::set_error("corrupt NetCDF header (variable list)");
    if (static_cast<int64_t>(n_vars) * 24 > r.n - r.at) return auvi::set_error("corrupt NetCDF header (variable count)");
    for (uint32_t v = 0; v < n_vars; ++v) {
        const std::string name = r.name();
        const uint32_t ndims = r.u32();
        if (!r.ok || static_cast<int64_t>(ndims) * 4 > r.n - r.at) return auvi::set_error("corrupt NetCDF header (variable entry)");
        std::vector<uint32_t> ids;
        for (uint32_t k = 0; k < ndims; ++k) ids.push_back(r.u32());
        std::vector<Attr> attrs;
        if (!r.ok || !read_attrs(r, &attrs)) return auvi::set_error("corrupt NetCDF header (variable attributes)");
        const int type = static_cast<int>(r.u32());
        r.u32();                                                  // vsize
        const int64_t begin = version == 1 ? static_cast<int64_t>(r.u32()) : static_cast<int64_t>(r.u64());
        if (!r.ok) return auvi::set_error("corrupt NetCDF header (variable entry)");
        if (name != var_name) continue;
        if (ndims > 4) return auvi::set_error("variable has more than 4 dimensions");
        out->nc_type = type;
        out->elem_bytes = type_size(type);
        out->ndims = static_cast<int32_t>(ndims);
        out->n_elems = 1;
        for (uint32_t k = 0; k < ndims; ++k) {
            if (ids[k] >= dim_len.size()) return auvi::set_error("corrupt NetCDF header (dimension id)");
            if (dim_len[ids[k]] == 0) return auvi::set_error("record (unlimited-dimension) variables are not supported");
            out->shape[k] = dim_len[ids[k]];
            if (out->shape[k] > (int64_t(1) << 40) / out->n_elems) return auvi::set_error("variable too large");   // no int64 overflow
            out->n_elems *= out->shape[k];
        }
        out->data_offset = begin;
        out->scale_factor = 1.0; out->add_offset = 0.0;
        for (const Attr& a : attrs) {
            if (a.nelems < 1) continue;
            if (a.name == "scale_factor") out->scale_factor = decode_be(r.p + a.at, a.type);
            if (a.name == "add_offset") out->add_offset = decode_be(r.p + a.at, a.type);
            if (a.name == "_FillValue") { out->has_fill = 1; out->fill_value = decode_be(r.p + a.at, a.type); }
        }
        if (out->elem_bytes == 0 || begin < 0 || begin > n_bytes || out->n_elems > (n_bytes - begin) / out->elem_bytes)
            return auvi::set_error("variable data lies outside the file image");
        return 0;
    }
    return auvi::set_error(std::string("no variable named '") + var_name + "' in the NetCDF file");
}

int auvi_netcdf3_read_f64(const void* file_image, int64_t n_bytes, const char* var_name, double* out, int64_t n_out) {
    auvi_nc_var v;
    if (auvi_netcdf3_find(file_image, n_bytes, var_name, &v)) return 1;
    if (!out || n_out < v.n_elems) return auvi::set_error("output buffer too small for the variable");
    const unsigned char* q = static_cast<const unsigned char*>(file_image) + v.data_offset;
    for (int64_t k = 0; k < v.n_elems; ++k) out[k] = decode_be(q + k * v.elem_bytes, v.nc_type) * v.scale_factor + v.add_offset;
    return 0;
}

int auvi_legacy_choice(int64_t total, int64_t n, uint32_t seed, int64_t* out_idx) {
    if (total < 0 || n < 0 || n > total) return auvi::set_error("choice needs 0 <= n <= total");
    if (n == 0) return 0;
    if (!out_idx) return auvi::set_error("null output");
    std::vector<int64_t> perm;
    try { perm.resize(static_cast<size_t>(total)); } catch (const std::bad_alloc&) { return auvi::set_error("out of host memory"); }
    for (int64_t k = 0; k < total; ++k) perm[k] = k;
    MT19937 rng(seed);
    for (int64_t i = total - 1; i >= 1; --i) {                    // legacy shuffle: Fisher-Yates from the top
        const int64_t j = static_cast<int64_t>(rng.interval(static_cast<uint64_t>(i)));
        const int64_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    std::memcpy(out_idx, perm.data(), sizeof(int64_t) * static_cast<size_t>(n));
    return 0;
}

}  // extern "C"
