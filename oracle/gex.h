// gex.h — tiny named-array container used by the test oracle tooling (ref_driver.cpp, ge_oracle.cpp).
// TEST INFRASTRUCTURE ONLY: nothing under oracle/ is part of the product path.
//
// File layout: magic "GEX1", then records
//   u32 name_len, name bytes, u32 dtype_len, dtype bytes (numpy style: u1,i4,u4,u8,i8,f8),
//   u32 ndim, u64 dims[ndim], u64 nbytes, raw little-endian data.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct GexWriter {
    FILE *f = nullptr;
    bool open(const std::string &path) {
        f = std::fopen(path.c_str(), "wb");
        if (!f) return false;
        std::fwrite("GEX1", 1, 4, f);
        return true;
    }
    void close() { if (f) std::fclose(f); f = nullptr; }
    void put_raw(const std::string &name, const char *dtype, const std::vector<uint64_t> &shape,
                 const void *data, uint64_t nbytes) {
        uint32_t nl = (uint32_t)name.size();
        std::fwrite(&nl, 4, 1, f); std::fwrite(name.data(), 1, nl, f);
        uint32_t dl = (uint32_t)std::strlen(dtype);
        std::fwrite(&dl, 4, 1, f); std::fwrite(dtype, 1, dl, f);
        uint32_t nd = (uint32_t)shape.size();
        std::fwrite(&nd, 4, 1, f);
        for (uint64_t d : shape) std::fwrite(&d, 8, 1, f);
        std::fwrite(&nbytes, 8, 1, f);
        if (nbytes) std::fwrite(data, 1, nbytes, f);
    }
    void put(const std::string &n, const std::vector<double> &v) { put_raw(n, "f8", {v.size()}, v.data(), v.size() * 8); }
    void put(const std::string &n, const std::vector<uint64_t> &v) { put_raw(n, "u8", {v.size()}, v.data(), v.size() * 8); }
    void put(const std::string &n, const std::vector<int64_t> &v) { put_raw(n, "i8", {v.size()}, v.data(), v.size() * 8); }
    void put(const std::string &n, const std::vector<int32_t> &v) { put_raw(n, "i4", {v.size()}, v.data(), v.size() * 4); }
    void put(const std::string &n, const std::vector<uint8_t> &v) { put_raw(n, "u1", {v.size()}, v.data(), v.size()); }
    void put2(const std::string &n, const std::vector<double> &v, uint64_t r, uint64_t c) { put_raw(n, "f8", {r, c}, v.data(), v.size() * 8); }
    void put2(const std::string &n, const std::vector<uint64_t> &v, uint64_t r, uint64_t c) { put_raw(n, "u8", {r, c}, v.data(), v.size() * 8); }
    void put2(const std::string &n, const std::vector<uint8_t> &v, uint64_t r, uint64_t c) { put_raw(n, "u1", {r, c}, v.data(), v.size()); }
    void scalar(const std::string &n, double x) { put_raw(n, "f8", {}, &x, 8); }
    void scalar_i(const std::string &n, int64_t x) { put_raw(n, "i8", {}, &x, 8); }
};
