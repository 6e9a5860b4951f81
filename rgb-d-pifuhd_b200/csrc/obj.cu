// Host-side OBJ serialisation: replaces the per-line Python loop of
// `save_obj_mesh_with_color` (`mesh_util.py:189-198`) with byte-identical text:
//   "v %.4f %.4f %.4f %.4f %.4f %.4f\n" per vertex (position, colour), then
//   "f %d %d %d\n" per face with 1-based indices written as (f0, f2, f1).
// `%.4f` is reproduced exactly (glibc prints the correctly rounded decimal of the binary value,
// ties to even): p = |x| * 1e4 rounded, e = fma(|x|, 1e4, -p) its exact residual, and the
// rounding decision looks at (frac(p) - 0.5, e) lexicographically.  Chunks of the mesh are
// formatted by a pool of host threads into private buffers and written in parallel at their final offsets.
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/pifu_b200.h"
#include "common.cuh"

namespace {

// "00" .. "99"
struct Digits2 {
    char t[200];
    constexpr Digits2() : t() {
        for (int i = 0; i < 100; ++i) { t[2 * i] = static_cast<char>('0' + i / 10); t[2 * i + 1] = static_cast<char>('0' + i % 10); }
    }
};
constexpr Digits2 DIGITS2;

inline char* put_uint(char* p, unsigned long long v) {
    if (v < 10) { *p++ = static_cast<char>('0' + v); return p; }        // mesh coordinates: the integer part is one digit
    int n;                                                              // digits (face indices: mostly 5-7)
    if (v < 10000ULL) n = v < 100ULL ? 2 : (v < 1000ULL ? 3 : 4);
    else if (v < 100000000ULL) n = v < 1000000ULL ? (v < 100000ULL ? 5 : 6) : (v < 10000000ULL ? 7 : 8);
    else { n = 9; for (unsigned long long w = v / 1000000000ULL; w; w /= 10) ++n; }
    char* e = p + n;
    char* q = e;
    while (v >= 100) {
        const unsigned r = static_cast<unsigned>(v % 100);
        v /= 100;
        q -= 2;
        q[0] = DIGITS2.t[2 * r]; q[1] = DIGITS2.t[2 * r + 1];
    }
    if (v >= 10) { q -= 2; q[0] = DIGITS2.t[2 * v]; q[1] = DIGITS2.t[2 * v + 1]; }
    else *--q = static_cast<char>('0' + v);
    return e;
}

// exactly the bytes of printf("%.4f", x)
inline char* put_fixed4(char* p, double x) {
    const double a = std::fabs(x);
    // the fast path needs |x| * 1e4 < 2^52 (a fraction bit left in the product); beyond 1e9, inf, nan: libc
    if (!(a < 1e9)) return p + snprintf(p, 400, "%.4f", x);
    if (std::signbit(x)) *p++ = '-';
    const double s = a * 1e4;
    // e = a * 1e4 - s exactly, without fma (a libm call on a generic x86-64 host build: it was most of this function's
    // time).  Dekker's product with Veltkamp's split of a; 1e4 = 625 * 2^4 has 10 significant bits, so it needs no
    // split and both partial products are exact.
    const double t = a * 134217729.0;                                    // 2^27 + 1
    const double ah = t - (t - a), al = a - ah;
    const double e = (ah * 1e4 - s) + al * 1e4;
    long long q = static_cast<long long>(s);                             // floor: 0 <= s < 1e13
    const double d = (s - static_cast<double>(q)) - 0.5;                 // exact
    if (d > 0.0 || (d == 0.0 && e > 0.0)) q += 1;
    else if (d == 0.0 && e == 0.0 && (q & 1LL)) q += 1;                   // tie: to even
    const unsigned long long ip = static_cast<unsigned long long>(q) / 10000ULL;
    p = put_uint(p, ip);
    const unsigned f = static_cast<unsigned>(static_cast<unsigned long long>(q) - ip * 10000ULL);
    *p++ = '.';
    const unsigned hi = f / 100, lo = f - hi * 100;
    p[0] = DIGITS2.t[2 * hi]; p[1] = DIGITS2.t[2 * hi + 1];
    p[2] = DIGITS2.t[2 * lo]; p[3] = DIGITS2.t[2 * lo + 1];
    return p + 4;
}

inline char* put_int(char* p, long long v) {
    if (v < 0) { *p++ = '-'; return put_uint(p, static_cast<unsigned long long>(-v)); }
    return put_uint(p, static_cast<unsigned long long>(v));
}

}  // namespace

namespace {

// one chunk of lines in a private, growable buffer (not zero-filled: a std::string's resize would memset it first)
struct Chunk {
    char* data = nullptr;
    size_t len = 0, cap = 0;
    ~Chunk() { free(data); }
    bool reserve(size_t need) {                        // room for `need` more bytes
        if (len + need <= cap) return true;
        size_t ncap = cap ? cap * 2 : need;
        while (ncap < len + need) ncap *= 2;
        char* n = static_cast<char*>(realloc(data, ncap));
        if (!n) return false;
        data = n; cap = ncap;
        return true;
    }
};

constexpr long long OBJ_CHUNK = 1 << 14;               // lines per chunk: ~1.4 MB of vertex text, ~0.35 MB of face text

bool format_chunk(Chunk& out, long long c, long long vchunks, const double* verts, const double* colors, long long nverts,
                  const int* faces, long long nfaces) {
    if (c < vchunks) {
        const long long b = c * OBJ_CHUNK, e = b + OBJ_CHUNK < nverts ? b + OBJ_CHUNK : nverts;
        if (!out.reserve(static_cast<size_t>(e - b) * 6 * 14 + 4096)) return false;
        for (long long i = b; i < e; ++i) {
            if (!out.reserve(2048)) return false;      // a line is < 6 * 320 bytes even for 1e308
            char* p = out.data + out.len;
            *p++ = 'v';
            for (int k = 0; k < 3; ++k) { *p++ = ' '; p = put_fixed4(p, verts[3 * i + k]); }
            for (int k = 0; k < 3; ++k) { *p++ = ' '; p = put_fixed4(p, colors[3 * i + k]); }
            *p++ = '\n';
            out.len = static_cast<size_t>(p - out.data);
        }
    } else {
        const long long b = (c - vchunks) * OBJ_CHUNK, e = b + OBJ_CHUNK < nfaces ? b + OBJ_CHUNK : nfaces;
        if (!out.reserve(static_cast<size_t>(e - b) * 40 + 16)) return false;
        char* p = out.data;
        for (long long i = b; i < e; ++i) {
            *p++ = 'f'; *p++ = ' ';
            p = put_int(p, static_cast<long long>(faces[3 * i]) + 1); *p++ = ' ';
            p = put_int(p, static_cast<long long>(faces[3 * i + 2]) + 1); *p++ = ' ';
            p = put_int(p, static_cast<long long>(faces[3 * i + 1]) + 1); *p++ = '\n';
        }
        out.len = static_cast<size_t>(p - out.data);
    }
    return true;
}

bool pwrite_all(int fd, const char* p, size_t n, off_t off) {
    while (n) {
        const ssize_t w = pwrite(fd, p, n, off);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= static_cast<size_t>(w); off += w;
    }
    return true;
}

}  // namespace

// A pool of host threads pulls chunk numbers from a counter and formats each chunk into its own buffer; the calling
// thread is the writer: it waits for chunk 0, 1, 2, ... in order and appends each to the file while the later chunks are
// still being formatted (the page-cache copy of 30 MB costs about as much as formatting it on 8-16 threads, and
// parallel pwrites of one file did not scale).
extern "C" int pifu_write_obj(const char* path, const double* verts, const double* colors, long long nverts,
                              const int* faces, long long nfaces) {
    if (!path || (nverts > 0 && (!verts || !colors)) || (nfaces > 0 && !faces) || nverts < 0 || nfaces < 0) {
        pifu::set_error("bad arguments to pifu_write_obj");
        return -1;
    }
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { pifu::set_error("cannot open %s for writing", path); return -1; }
    const long long vchunks = (nverts + OBJ_CHUNK - 1) / OBJ_CHUNK, fchunks = (nfaces + OBJ_CHUNK - 1) / OBJ_CHUNK;
    const long long total = vchunks + fchunks;
    std::vector<Chunk> chunks(static_cast<size_t>(total));
    std::vector<std::atomic<int>> state(static_cast<size_t>(total));     // 0 pending, 1 formatted, -1 failed
    for (auto& st : state) st.store(0, std::memory_order_relaxed);
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = static_cast<int>(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    if (const char* env = getenv("PIFU_OBJ_THREADS")) { const int n = atoi(env); if (n >= 1 && n <= 64) nthreads = n; }
    int nworkers = nthreads > 1 ? nthreads - 1 : 1;
    if (nworkers > total) nworkers = static_cast<int>(total > 0 ? total : 1);
    const bool timing = getenv("PIFU_OBJ_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    std::atomic<long long> next{0};
    std::atomic<bool> stop{false};
    auto work = [&]() {
        for (long long c = next.fetch_add(1); c < total && !stop.load(std::memory_order_relaxed); c = next.fetch_add(1)) {
            const bool fine = format_chunk(chunks[static_cast<size_t>(c)], c, vchunks, verts, colors, nverts, faces, nfaces);
            state[static_cast<size_t>(c)].store(fine ? 1 : -1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nworkers; ++t) pool.emplace_back(work);
    bool ok = true, oom = false;
    off_t off = 0;
    double wait_ms = 0.0;
    for (long long c = 0; c < total && ok; ++c) {
        int st;
        const auto w0 = std::chrono::steady_clock::now();
        while ((st = state[static_cast<size_t>(c)].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
        wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        if (st < 0) { ok = false; oom = true; break; }
        Chunk& k = chunks[static_cast<size_t>(c)];
        ok = pwrite_all(fd, k.data, k.len, off);
        off += static_cast<off_t>(k.len);
        free(k.data);                                                    // the text of a written chunk is not needed again
        k.data = nullptr; k.len = k.cap = 0;
    }
    if (!ok) stop.store(true);
    for (auto& th : pool) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (close(fd) != 0) ok = false;
    if (timing)
        fprintf(stderr, "pifu_write_obj: %.2f ms (writer waited %.2f ms for chunks), close %.2f ms, %d formatting threads\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(), wait_ms,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), nworkers);
    if (!ok) {
        if (oom) pifu::set_error("out of memory formatting %s", path); else pifu::set_error("write to %s failed", path);
        return -1;
    }
    return 0;
}
