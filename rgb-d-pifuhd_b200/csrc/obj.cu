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
#include <sys/mman.h>
#include <sys/stat.h>
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

// ---------------------------------------------------------------------------------------------------------------------
// OBJ reader: the inverse of the writer, for `meshcleaning(obj_path)` (`reconstruction.py:325-344` loads the file the
// pipeline wrote a moment ago; a per-line Python loop needs 2.5 s for the 1 M lines of a 512^3 mesh).  The file is
// mapped, cut into per-thread segments at line starts, counted, then parsed in place:
//   "v x y z [r g b]"  -> doubles (decimal fast path: up to 15 significant digits and 10^k <= 10^22 is one correctly
//                         rounded division of two exact doubles; anything else goes through strtod)
//   "f a[/..] b[/..] c[/..]" -> 0-based int32 (a, c, b): the order the faces had when they were GIVEN to the writer
// Every other line is skipped, like the Python loop this replaces (`line.startswith("v ")` / `"f "`).
namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap(const_cast<char*>(p), n);
        if (fd >= 0) close(fd);
    }
    bool open_file(const char* path) {
        fd = open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = static_cast<size_t>(st.st_size);
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { n = 0; return false; }
        p = static_cast<const char*>(m);
        return true;
    }
};

inline const char* line_end(const char* q, const char* e) {
    const void* nl = memchr(q, '\n', static_cast<size_t>(e - q));
    return nl ? static_cast<const char*>(nl) : e;
}
inline bool blank(char ch) { return ch == ' ' || ch == '\t' || ch == '\r'; }

// one number of a vertex line starting at q (< le); advances q past it.  false: no number there
inline bool parse_double(const char*& q, const char* le, double* out) {
    static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                   1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* s = q;
    bool neg = false;
    if (s < le && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
    unsigned long long m = 0;
    int digits = 0, frac = 0;
    const char* d0 = s;
    while (s < le && *s >= '0' && *s <= '9') { m = m * 10 + static_cast<unsigned>(*s - '0'); ++digits; ++s; }
    if (s < le && *s == '.') {
        ++s;
        while (s < le && *s >= '0' && *s <= '9') { m = m * 10 + static_cast<unsigned>(*s - '0'); ++digits; ++frac; ++s; }
    }
    const bool plain = s > d0 && digits > 0 && (s == le || blank(*s));
    if (plain && digits <= 15 && frac <= 22) {                   // (leading zeros only make `digits` pessimistic)
        const double v = static_cast<double>(m) / P10[frac];
        *out = neg ? -v : v;
        q = s;
        return true;
    }
    // exponents, inf / nan, long mantissas: libc, on a terminated copy of the token
    const char* t = q;
    while (t < le && !blank(*t)) ++t;
    const size_t len = static_cast<size_t>(t - q);
    if (len == 0 || len > 400) return false;
    char buf[408];
    memcpy(buf, q, len);
    buf[len] = 0;
    char* endp = nullptr;
    const double v = strtod(buf, &endp);
    if (endp != buf + len) return false;
    *out = v;
    q = t;
    return true;
}

struct ObjCounts { long long nv = 0, nf = 0; int cols = 0; bool bad = false; };

// phase 1 (count) or phase 2 (parse into the buffers); [b, e) starts at a line start
void scan_segment(const char* b, const char* e, ObjCounts& cnt, bool parse, double* verts, double* colors, int* faces,
                  long long v0, long long f0) {
    long long iv = v0, jf = f0;
    for (const char* q = b; q < e;) {
        const char* le = line_end(q, e);
        if (le - q >= 2 && q[0] == 'v' && q[1] == ' ') {
            if (!parse) {
                if (cnt.cols == 0) {                               // values on the first vertex line of the segment
                    int k = 0;
                    const char* s = q + 2;
                    double tmp;
                    for (;;) {
                        while (s < le && blank(*s)) ++s;
                        if (s >= le || !parse_double(s, le, &tmp)) break;
                        ++k;
                    }
                    cnt.cols = k;
                }
                ++cnt.nv;
            } else {
                const char* s = q + 2;
                double val[6] = {0, 0, 0, 0, 0, 0};
                int k = 0;
                for (; k < 6; ++k) {
                    while (s < le && blank(*s)) ++s;
                    if (s >= le || !parse_double(s, le, &val[k])) break;
                }
                if (k < 3 || (colors != nullptr && k < 6)) cnt.bad = true;
                verts[3 * iv] = val[0]; verts[3 * iv + 1] = val[1]; verts[3 * iv + 2] = val[2];
                if (colors) { colors[3 * iv] = val[3]; colors[3 * iv + 1] = val[4]; colors[3 * iv + 2] = val[5]; }
                ++iv;
            }
        } else if (le - q >= 2 && q[0] == 'f' && q[1] == ' ') {
            if (!parse) {
                ++cnt.nf;
            } else {
                const char* s = q + 2;
                long long idx[3] = {0, 0, 0};
                int k = 0;
                for (; k < 3; ++k) {
                    while (s < le && blank(*s)) ++s;
                    if (s >= le) break;
                    bool neg = false;
                    if (*s == '-') { neg = true; ++s; }
                    const char* d0 = s;
                    long long a = 0;
                    while (s < le && *s >= '0' && *s <= '9') { a = a * 10 + (*s - '0'); ++s; }
                    if (s == d0) break;
                    idx[k] = neg ? -a : a;
                    while (s < le && !blank(*s)) ++s;              // "/vt/vn" of the reference: ignored
                }
                if (k < 3) cnt.bad = true;
                faces[3 * jf] = static_cast<int>(idx[0] - 1);
                faces[3 * jf + 1] = static_cast<int>(idx[2] - 1);
                faces[3 * jf + 2] = static_cast<int>(idx[1] - 1);
                ++jf;
            }
        }
        q = le < e ? le + 1 : e;
    }
}

// segment s of T: starts at the first line start at or after s * n / T
std::vector<const char*> segment_starts(const Mapped& m, int T) {
    std::vector<const char*> st(static_cast<size_t>(T) + 1, m.p + m.n);
    st[0] = m.p;
    for (int s = 1; s < T; ++s) {
        const char* q = m.p + m.n / static_cast<size_t>(T) * static_cast<size_t>(s);
        if (q > m.p && q[-1] != '\n') { const char* le = line_end(q, m.p + m.n); q = le < m.p + m.n ? le + 1 : m.p + m.n; }
        st[static_cast<size_t>(s)] = q;
    }
    for (int s = 1; s <= T; ++s)
        if (st[static_cast<size_t>(s)] < st[static_cast<size_t>(s) - 1]) st[static_cast<size_t>(s)] = st[static_cast<size_t>(s) - 1];
    return st;
}

int obj_threads(size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    int n = static_cast<int>(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    if (const char* env = getenv("PIFU_OBJ_THREADS")) { const int k = atoi(env); if (k >= 1 && k <= 64) n = k; }
    const size_t by_size = bytes / (256 * 1024) + 1;               // small files: not worth the threads
    return static_cast<int>(by_size < static_cast<size_t>(n) ? by_size : static_cast<size_t>(n));
}

template <typename F>
void run_segments(int T, F&& body) {
    std::vector<std::thread> pool;
    for (int s = 1; s < T; ++s) pool.emplace_back([&, s]() { body(s); });
    body(0);
    for (auto& th : pool) th.join();
}

int obj_scan(const char* path, long long* counts, bool parse, double* verts, double* colors, int* faces, long long nverts,
             long long nfaces) {
    Mapped m;
    if (!path || !m.open_file(path)) { pifu::set_error("cannot read %s", path ? path : "(null)"); return -1; }
    const int T = obj_threads(m.n);
    const std::vector<const char*> st = segment_starts(m, T);
    std::vector<ObjCounts> cnt(static_cast<size_t>(T));
    run_segments(T, [&](int s) { scan_segment(st[static_cast<size_t>(s)], st[static_cast<size_t>(s) + 1], cnt[static_cast<size_t>(s)], false, nullptr, nullptr, nullptr, 0, 0); });
    long long nv = 0, nf = 0;
    int cols = 0;
    std::vector<long long> v0(static_cast<size_t>(T)), f0(static_cast<size_t>(T));
    for (int s = 0; s < T; ++s) {
        v0[static_cast<size_t>(s)] = nv; f0[static_cast<size_t>(s)] = nf;
        nv += cnt[static_cast<size_t>(s)].nv; nf += cnt[static_cast<size_t>(s)].nf;
        if (cols == 0) cols = cnt[static_cast<size_t>(s)].cols;
    }
    if (counts) { counts[0] = nv; counts[1] = nf; counts[2] = cols; }
    if (!parse) return 0;
    if (nv != nverts || nf != nfaces || (nv > 0 && !verts) || (nf > 0 && !faces)) {
        pifu::set_error("%s holds %lld vertices and %lld faces, the buffers %lld and %lld", path, nv, nf, nverts, nfaces);
        return -1;
    }
    if (colors && cols < 6) { pifu::set_error("%s has no vertex colours (%d values per vertex line)", path, cols); return -1; }
    run_segments(T, [&](int s) {
        scan_segment(st[static_cast<size_t>(s)], st[static_cast<size_t>(s) + 1], cnt[static_cast<size_t>(s)], true, verts, colors, faces,
                     v0[static_cast<size_t>(s)], f0[static_cast<size_t>(s)]);
    });
    for (int s = 0; s < T; ++s)
        if (cnt[static_cast<size_t>(s)].bad) { pifu::set_error("%s: malformed vertex or face line", path); return -1; }
    return 0;
}

}  // namespace

extern "C" int pifu_obj_counts(const char* path, long long* counts) {
    if (!counts) { pifu::set_error("bad arguments to pifu_obj_counts"); return -1; }
    return obj_scan(path, counts, false, nullptr, nullptr, nullptr, 0, 0);
}

extern "C" int pifu_read_obj(const char* path, double* verts, double* colors, int* faces, long long nverts, long long nfaces) {
    if (nverts < 0 || nfaces < 0) { pifu::set_error("bad arguments to pifu_read_obj"); return -1; }
    return obj_scan(path, nullptr, true, verts, colors, faces, nverts, nfaces);
}
