// phox_npy.h : minimal .npy reader/writer with the semantics the reference's NP.hh relies on
// (sysrap/NP.hh): little-endian, C-order, version 1.0 header, dtypes <f4 <f8 <i4 <u4 <i8 <u8.
// Enough to load a persisted CSGFoundry directory (CSG/CSGFoundry.cc:2768-2802) and the SSim tables.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace phoxnpy {

struct Array {
    std::string dtype;                 // e.g. "<f4"
    std::vector<int64_t> shape;
    std::vector<char> data;
    int64_t count() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
    int itemsize() const { return dtype.size() >= 3 ? dtype[2] - '0' : 0; }
    template <typename T> const T* as() const { return reinterpret_cast<const T*>(data.data()); }
    bool empty() const { return data.empty(); }
};

inline Array load(const std::string& path, bool required = true) {
    Array a;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { if (required) throw std::runtime_error("cannot open " + path); return a; }
    char magic[10];
    if (std::fread(magic, 1, 10, f) != 10 || std::memcmp(magic, "\x93NUMPY", 6) != 0) { std::fclose(f); throw std::runtime_error("not an npy file: " + path); }
    int major = magic[6];
    size_t hlen = (unsigned char)magic[8] | ((unsigned char)magic[9] << 8);
    if (major >= 2) { unsigned char more[2]; if (std::fread(more, 1, 2, f) != 2) { std::fclose(f); throw std::runtime_error("short header"); } hlen |= (size_t)more[0] << 16 | (size_t)more[1] << 24; }
    std::string hdr(hlen, ' ');
    if (std::fread(&hdr[0], 1, hlen, f) != hlen) { std::fclose(f); throw std::runtime_error("short header: " + path); }
    size_t p = hdr.find("'descr':");
    size_t q0 = hdr.find('\'', p + 8), q1 = hdr.find('\'', q0 + 1);
    a.dtype = hdr.substr(q0 + 1, q1 - q0 - 1);
    if (hdr.find("'fortran_order': True") != std::string::npos) { std::fclose(f); throw std::runtime_error("fortran order not supported: " + path); }
    size_t s0 = hdr.find('(', hdr.find("'shape':")), s1 = hdr.find(')', s0);
    std::string sh = hdr.substr(s0 + 1, s1 - s0 - 1);
    size_t pos = 0;
    while (pos < sh.size()) {
        while (pos < sh.size() && (sh[pos] == ' ' || sh[pos] == ',')) pos++;
        if (pos >= sh.size()) break;
        a.shape.push_back(std::strtoll(sh.c_str() + pos, nullptr, 10));
        while (pos < sh.size() && sh[pos] != ',') pos++;
    }
    size_t bytes = (size_t)a.count() * a.itemsize();
    a.data.resize(bytes);
    if (bytes && std::fread(a.data.data(), 1, bytes, f) != bytes) { std::fclose(f); throw std::runtime_error("short data: " + path); }
    std::fclose(f);
    return a;
}

inline void save(const std::string& path, const char* dtype, const std::vector<int64_t>& shape, const void* data, size_t bytes) {
    std::string sh = "(";
    for (size_t i = 0; i < shape.size(); i++) sh += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
    sh += ")";
    std::string hdr = "{'descr': '" + std::string(dtype) + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ') + "\n";
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hdr.size() & 0xff), (unsigned char)(hdr.size() >> 8)};
    std::fwrite(pre, 1, 10, f);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    if (bytes) std::fwrite(data, 1, bytes, f);
    std::fclose(f);
}

}  // namespace phoxnpy
