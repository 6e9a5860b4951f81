// Host-side OBJ serialisation: replaces the per-line Python loop of
// `save_obj_mesh_with_color` (`mesh_util.py:189-198`) with byte-identical text:
//   "v %.4f %.4f %.4f %.4f %.4f %.4f\n" per vertex (position, colour), then
//   "f %d %d %d\n" per face with 1-based indices written as (f0, f2, f1).
// `%.4f` is reproduced exactly (glibc prints the correctly rounded decimal of the binary value,
// ties to even): p = |x| * 1e4 rounded, e = fma(|x|, 1e4, -p) its exact residual, and the
// rounding decision looks at (frac(p) - 0.5, e) lexicographically.  Chunks of the mesh are
// formatted by a few host threads into private buffers and written in order.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pifu_b200.h"
#include "common.cuh"

namespace {

inline char* put_uint(char* p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// exactly the bytes of printf("%.4f", x)
inline char* put_fixed4(char* p, double x) {
    const double a = std::fabs(x);
    // the fast path needs |x| * 1e4 < 2^52 (a fraction bit left in the product); beyond 1e9, inf, nan: libc
    if (!(a < 1e9)) return p + snprintf(p, 400, "%.4f", x);
    if (std::signbit(x)) *p++ = '-';
    const double s = a * 1e4;
    const double e = std::fma(a, 1e4, -s);                               // s + e == a * 1e4 exactly
    double r = std::floor(s);
    const double d = (s - r) - 0.5;                                      // exact
    if (d > 0.0 || (d == 0.0 && e > 0.0)) r += 1.0;
    else if (d == 0.0 && e == 0.0 && std::fmod(r, 2.0) != 0.0) r += 1.0;  // tie: to even
    const unsigned long long q = static_cast<unsigned long long>(r);
    p = put_uint(p, q / 10000ULL);
    const unsigned f = static_cast<unsigned>(q % 10000ULL);
    *p++ = '.';
    *p++ = static_cast<char>('0' + f / 1000);
    *p++ = static_cast<char>('0' + (f / 100) % 10);
    *p++ = static_cast<char>('0' + (f / 10) % 10);
    *p++ = static_cast<char>('0' + f % 10);
    return p;
}

inline char* put_int(char* p, long long v) {
    if (v < 0) { *p++ = '-'; return put_uint(p, static_cast<unsigned long long>(-v)); }
    return put_uint(p, static_cast<unsigned long long>(v));
}

}  // namespace

extern "C" int pifu_write_obj(const char* path, const double* verts, const double* colors, long long nverts,
                              const int* faces, long long nfaces) {
    if (!path || (nverts > 0 && (!verts || !colors)) || (nfaces > 0 && !faces) || nverts < 0 || nfaces < 0) {
        pifu::set_error("bad arguments to pifu_write_obj");
        return -1;
    }
    FILE* f = fopen(path, "wb");
    if (!f) { pifu::set_error("cannot open %s for writing", path); return -1; }
    const long long chunk = 1 << 16;
    const long long vchunks = (nverts + chunk - 1) / chunk, fchunks = (nfaces + chunk - 1) / chunk;
    const long long total = vchunks + fchunks;
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = static_cast<int>(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    bool ok = true;
    // rounds of `nthreads` chunks: format in parallel, write in order
    for (long long c0 = 0; c0 < total && ok; c0 += nthreads) {
        const int nc = static_cast<int>(total - c0 < nthreads ? total - c0 : nthreads);
        std::vector<std::string> bufs(nc);
        std::vector<std::thread> pool;
        for (int t = 0; t < nc; ++t) {
            pool.emplace_back([&, t]() {
                const long long c = c0 + t;
                std::string& out = bufs[t];
                if (c < vchunks) {
                    const long long b = c * chunk, e = b + chunk < nverts ? b + chunk : nverts;
                    out.resize(static_cast<size_t>(e - b) * 6 * 14 + 4096);
                    char* p = &out[0];
                    for (long long i = b; i < e; ++i) {
                        if (static_cast<size_t>(&out[0] + out.size() - p) < 2048) {      // a line is < 6 * 320 bytes even for 1e308
                            const size_t used = static_cast<size_t>(p - &out[0]);
                            out.resize(out.size() * 2);
                            p = &out[0] + used;
                        }
                        *p++ = 'v';
                        for (int k = 0; k < 3; ++k) { *p++ = ' '; p = put_fixed4(p, verts[3 * i + k]); }
                        for (int k = 0; k < 3; ++k) { *p++ = ' '; p = put_fixed4(p, colors[3 * i + k]); }
                        *p++ = '\n';
                    }
                    out.resize(static_cast<size_t>(p - &out[0]));
                } else {
                    const long long b = (c - vchunks) * chunk, e = b + chunk < nfaces ? b + chunk : nfaces;
                    out.resize(static_cast<size_t>(e - b) * 40 + 16);
                    char* p = &out[0];
                    for (long long i = b; i < e; ++i) {
                        *p++ = 'f'; *p++ = ' ';
                        p = put_int(p, static_cast<long long>(faces[3 * i]) + 1); *p++ = ' ';
                        p = put_int(p, static_cast<long long>(faces[3 * i + 2]) + 1); *p++ = ' ';
                        p = put_int(p, static_cast<long long>(faces[3 * i + 1]) + 1); *p++ = '\n';
                    }
                    out.resize(static_cast<size_t>(p - &out[0]));
                }
            });
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < nc && ok; ++t)
            ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) { pifu::set_error("write to %s failed", path); return -1; }
    return 0;
}
