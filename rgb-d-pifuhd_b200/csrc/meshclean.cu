// Largest connected component of a triangle mesh on the device, replacing `meshcleaning` (`reconstruction.py:325-344`:
// trimesh.load -> mesh.split() -> keep the component with the greatest extent along x -> export).
//
// trimesh is not available offline; what it does for this call is restated here (and, independently, in
// oracle/mesh_oracle.py with scipy): faces are adjacent when they share an edge that exactly two faces use
// (`face_adjacency`), components are the connected components of that graph, `split()` keeps only WATERTIGHT
// components (every edge of the component used by exactly two of its faces) of at least 4 faces, in the order of
// their first face; `meshcleaning` keeps the first component of maximal extent along axis 0.
//
// Edges go into an open-addressing hash table (64-bit key = the two vertex numbers, atomicCAS claim; per slot the use
// count and the first two faces), components come from a lock-free union-find over faces (larger root hooks under the
// smaller, so a component's root is its first face), statistics from atomics keyed by root, and the kept faces /
// vertices are compacted with the library's device-wide scan.
#include "common.cuh"
#include "internal.h"
#include "scan.cuh"

namespace pifu {

namespace {

constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ unsigned long long hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void edge_insert_kernel(const int* __restrict__ faces, long long nf, unsigned long long* __restrict__ keys,
                                   uint32_t* __restrict__ count, int* __restrict__ first, int* __restrict__ second,
                                   unsigned long long mask) {
    const long long t = blockIdx.x * 256LL + threadIdx.x;
    if (t >= 3 * nf) return;
    const long long f = t / 3;
    const int e = static_cast<int>(t - 3 * f);
    const int a = faces[3 * f + e], b = faces[3 * f + (e + 1) % 3];
    const unsigned long long lo = static_cast<unsigned long long>(a < b ? a : b), hi = static_cast<unsigned long long>(a < b ? b : a);
    const unsigned long long key = (lo << 32) | hi;
    unsigned long long slot = hash64(key) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(keys + slot, EMPTY_KEY, key);
        if (old == EMPTY_KEY || old == key) break;
        slot = (slot + 1) & mask;
    }
    const uint32_t c = atomicAdd(count + slot, 1u);
    if (c == 0u) first[slot] = static_cast<int>(f);
    else if (c == 1u) second[slot] = static_cast<int>(f);
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
    for (;;) {
        const int p = parent[x];
        if (p == x) return x;
        const int g = parent[p];
        if (g != p) parent[x] = g;                 // path halving (benign race: parents only ever move towards the root)
        x = p;
    }
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // a: the larger root, hooks under b
        if (atomicCAS(parent + a, a, b) == a) return;
    }
}

__global__ void init_parent_kernel(int* __restrict__ parent, long long nf) {
    const long long f = blockIdx.x * 256LL + threadIdx.x;
    if (f < nf) parent[f] = static_cast<int>(f);
}

// faces that share an edge used by exactly two faces are adjacent
__global__ void edge_union_kernel(const uint32_t* __restrict__ count, const int* __restrict__ first, const int* __restrict__ second,
                                  unsigned long long slots, int* __restrict__ parent) {
    const unsigned long long s = blockIdx.x * 256ull + threadIdx.x;
    if (s < slots && count[s] == 2u) uf_union(parent, first[s], second[s]);
}

__device__ __forceinline__ long long order_key(double v) {      // monotone map double -> int64
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : b ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double order_val(long long k) {
    return __longlong_as_double(k >= 0 ? k : k ^ 0x7fffffffffffffffLL);
}

struct CompStats {
    int* root;              // [nf] component (its first face) of every face
    uint32_t* nfaces;       // [nf] by root
    uint32_t* open_edges;   // [nf] by root: edges not used by exactly two faces
    long long* lo;          // [nf] by root: extent along the axis
    long long* hi;
};

__global__ void face_stats_kernel(const double* __restrict__ verts, const int* __restrict__ faces, long long nf, int axis,
                                  int* __restrict__ parent, CompStats st) {
    const long long f = blockIdx.x * 256LL + threadIdx.x;
    if (f >= nf) return;
    const int r = uf_find(parent, static_cast<int>(f));
    st.root[f] = r;
    atomicAdd(st.nfaces + r, 1u);
    double lo = verts[3 * static_cast<long long>(faces[3 * f]) + axis], hi = lo;
    for (int q = 1; q < 3; ++q) {
        const double v = verts[3 * static_cast<long long>(faces[3 * f + q]) + axis];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    atomicMin(st.lo + r, order_key(lo));
    atomicMax(st.hi + r, order_key(hi));
}

__global__ void edge_stats_kernel(const uint32_t* __restrict__ count, const int* __restrict__ first, unsigned long long slots,
                                  int* __restrict__ parent, uint32_t* __restrict__ open_edges) {
    const unsigned long long s = blockIdx.x * 256ull + threadIdx.x;
    if (s < slots && count[s] != 0u && count[s] != 2u) atomicAdd(open_edges + uf_find(parent, first[s]), 1u);
}

// one block: the first component (lowest first face) of maximal extent among the eligible ones -> *winner (-1: none)
__global__ void __launch_bounds__(1024) pick_kernel(CompStats st, long long nf, int only_watertight, int min_faces, int* __restrict__ winner) {
    __shared__ double s_h[1024];
    __shared__ int s_r[1024];
    double best = -1.0;
    int br = -1;
    for (long long f = threadIdx.x; f < nf; f += 1024) {
        if (st.root[f] != f) continue;
        if (st.nfaces[f] < static_cast<uint32_t>(min_faces)) continue;
        if (only_watertight && st.open_edges[f] != 0u) continue;
        const double h = order_val(st.hi[f]) - order_val(st.lo[f]);
        if (h > best) { best = h; br = static_cast<int>(f); }          // ascending f per thread: ties keep the lower face
    }
    s_h[threadIdx.x] = best; s_r[threadIdx.x] = br;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double h2 = s_h[threadIdx.x + o];
            const int r2 = s_r[threadIdx.x + o];
            const bool take = r2 >= 0 && (s_r[threadIdx.x] < 0 || h2 > s_h[threadIdx.x] || (h2 == s_h[threadIdx.x] && r2 < s_r[threadIdx.x]));
            if (take) { s_h[threadIdx.x] = h2; s_r[threadIdx.x] = r2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *winner = s_r[0];
}

__global__ void mark_kernel(const int* __restrict__ faces, const int* __restrict__ root, long long nf, const int* __restrict__ winner,
                            uint32_t* __restrict__ fkeep, uint32_t* __restrict__ vkeep) {
    const long long f = blockIdx.x * 256LL + threadIdx.x;
    if (f >= nf) return;
    const bool k = *winner >= 0 && root[f] == *winner;
    fkeep[f] = k ? 1u : 0u;
    if (k) for (int q = 0; q < 3; ++q) vkeep[faces[3 * f + q]] = 1u;
}

// per-element exclusive offsets from the scan of per-tile sums: tile = 256 consecutive elements
__global__ void tile_sums_kernel(const uint32_t* __restrict__ flags, long long n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t red[SCAN_BLOCK / 32];
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    const uint32_t t = block_sum(i < n ? flags[i] : 0u, red);
    if (threadIdx.x == 0) sums[blockIdx.x] = t;
}

__global__ void compact_verts_kernel(const uint32_t* __restrict__ vkeep, const uint32_t* __restrict__ offs, long long nv,
                                     const double* __restrict__ verts, const double* __restrict__ colors,
                                     double* __restrict__ out_verts, double* __restrict__ out_colors, int* __restrict__ newid) {
    const long long v = blockIdx.x * 256LL + threadIdx.x;
    const uint32_t flag = v < nv ? vkeep[v] : 0u;
    uint32_t bt;
    const uint32_t rel = block_exclusive_scan(flag, &bt);
    if (v >= nv) return;
    if (flag) {
        const long long o = static_cast<long long>(offs[blockIdx.x]) + rel;
        newid[v] = static_cast<int>(o);
        for (int q = 0; q < 3; ++q) {
            out_verts[3 * o + q] = verts[3 * v + q];
            if (colors) out_colors[3 * o + q] = colors[3 * v + q];
        }
    }
}

__global__ void compact_faces_kernel(const uint32_t* __restrict__ fkeep, const uint32_t* __restrict__ offs, long long nf,
                                     const int* __restrict__ faces, const int* __restrict__ newid, int* __restrict__ out_faces) {
    const long long f = blockIdx.x * 256LL + threadIdx.x;
    const uint32_t flag = f < nf ? fkeep[f] : 0u;
    uint32_t bt;
    const uint32_t rel = block_exclusive_scan(flag, &bt);
    if (f >= nf || !flag) return;
    const long long o = static_cast<long long>(offs[blockIdx.x]) + rel;
    for (int q = 0; q < 3; ++q) out_faces[3 * o + q] = newid[faces[3 * f + q]];
}

struct Scratch {
    void* p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
};

}  // namespace

int mesh_clean(const double* verts, const double* colors, const int* faces, long long nv, long long nf, int only_watertight,
               int axis, double* out_verts, double* out_colors, int* out_faces, long long* counts_host, int num_sms, cudaStream_t s) {
    (void)num_sms;
    counts_host[0] = counts_host[1] = 0;
    if (nv <= 0 || nf <= 0) { set_error("mesh_clean: empty mesh"); return -1; }
    if (nf > 0x7ffffff0LL / 3 || nv > 0x7ffffff0LL) { set_error("mesh_clean: mesh too large"); return -1; }
    unsigned long long slots = 1;
    while (slots < static_cast<unsigned long long>(6 * nf)) slots <<= 1;        // load factor <= 0.5 (3 nf / 2 distinct edges when closed)
    const long long vtiles = (nv + 255) / 256, ftiles = (nf + 255) / 256;
    // one scratch allocation, carved up
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~static_cast<size_t>(255); return o; };
    const size_t o_keys = take(slots * 8), o_count = take(slots * 4), o_first = take(slots * 4), o_second = take(slots * 4);
    const size_t o_parent = take(nf * 4), o_root = take(nf * 4), o_nfaces = take(nf * 4), o_open = take(nf * 4);
    const size_t o_lo = take(nf * 8), o_hi = take(nf * 8), o_fkeep = take(nf * 4), o_vkeep = take(nv * 4), o_newid = take(nv * 4);
    const size_t o_vs = take((vtiles + 1) * 4), o_fs = take((ftiles + 1) * 4);
    const size_t o_part = take(static_cast<size_t>(scan_partials_needed(vtiles > ftiles ? vtiles : ftiles)) * 4);
    const size_t o_tot = take(4 * 8), o_win = take(4);
    Scratch sc;
    PIFU_CUDA(cudaMalloc(&sc.p, off));
    uint8_t* b = static_cast<uint8_t*>(sc.p);
    auto* keys = reinterpret_cast<unsigned long long*>(b + o_keys);
    auto* count = reinterpret_cast<uint32_t*>(b + o_count);
    auto* first = reinterpret_cast<int*>(b + o_first);
    auto* second = reinterpret_cast<int*>(b + o_second);
    auto* parent = reinterpret_cast<int*>(b + o_parent);
    CompStats st;
    st.root = reinterpret_cast<int*>(b + o_root);
    st.nfaces = reinterpret_cast<uint32_t*>(b + o_nfaces);
    st.open_edges = reinterpret_cast<uint32_t*>(b + o_open);
    st.lo = reinterpret_cast<long long*>(b + o_lo);
    st.hi = reinterpret_cast<long long*>(b + o_hi);
    auto* fkeep = reinterpret_cast<uint32_t*>(b + o_fkeep);
    auto* vkeep = reinterpret_cast<uint32_t*>(b + o_vkeep);
    auto* newid = reinterpret_cast<int*>(b + o_newid);
    auto* vs = reinterpret_cast<uint32_t*>(b + o_vs);
    auto* fs = reinterpret_cast<uint32_t*>(b + o_fs);
    auto* part = reinterpret_cast<uint32_t*>(b + o_part);
    auto* tot = reinterpret_cast<unsigned long long*>(b + o_tot);
    auto* win = reinterpret_cast<int*>(b + o_win);
    PIFU_CUDA(cudaMemsetAsync(keys, 0xff, slots * 8, s));
    PIFU_CUDA(cudaMemsetAsync(count, 0, slots * 4, s));
    PIFU_CUDA(cudaMemsetAsync(st.nfaces, 0, nf * 4, s));
    PIFU_CUDA(cudaMemsetAsync(st.open_edges, 0, nf * 4, s));
    PIFU_CUDA(cudaMemsetAsync(st.lo, 0x7f, nf * 8, s));          // large positive keys
    PIFU_CUDA(cudaMemsetAsync(st.hi, 0x80, nf * 8, s));          // large negative keys
    PIFU_CUDA(cudaMemsetAsync(vkeep, 0, nv * 4, s));
    const unsigned fb = static_cast<unsigned>(ftiles), vb = static_cast<unsigned>(vtiles);
    const unsigned sb = static_cast<unsigned>((slots + 255) / 256);
    init_parent_kernel<<<fb, 256, 0, s>>>(parent, nf);
    edge_insert_kernel<<<static_cast<unsigned>((3 * nf + 255) / 256), 256, 0, s>>>(faces, nf, keys, count, first, second, slots - 1);
    edge_union_kernel<<<sb, 256, 0, s>>>(count, first, second, slots, parent);
    face_stats_kernel<<<fb, 256, 0, s>>>(verts, faces, nf, axis, parent, st);
    edge_stats_kernel<<<sb, 256, 0, s>>>(count, first, slots, parent, st.open_edges);
    pick_kernel<<<1, 1024, 0, s>>>(st, nf, only_watertight, only_watertight ? 4 : 1, win);
    mark_kernel<<<fb, 256, 0, s>>>(faces, st.root, nf, win, fkeep, vkeep);
    tile_sums_kernel<<<vb, 256, 0, s>>>(vkeep, nv, vs);
    device_exclusive_scan(vs, nullptr, vtiles, part, tot, s);
    compact_verts_kernel<<<vb, 256, 0, s>>>(vkeep, vs, nv, verts, colors, out_verts, out_colors, newid);
    tile_sums_kernel<<<fb, 256, 0, s>>>(fkeep, nf, fs);
    device_exclusive_scan(fs, nullptr, ftiles, part, tot + 2, s);
    compact_faces_kernel<<<fb, 256, 0, s>>>(fkeep, fs, nf, faces, newid, out_faces);
    PIFU_CUDA(cudaGetLastError());
    unsigned long long h[4] = {0, 0, 0, 0};
    int hw = -1;
    PIFU_CUDA(cudaMemcpyAsync(h, tot, sizeof(h), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaMemcpyAsync(&hw, win, sizeof(hw), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    if (hw < 0) { set_error("mesh_clean: no %scomponent of at least %d faces", only_watertight ? "watertight " : "", only_watertight ? 4 : 1); return -1; }
    counts_host[0] = static_cast<long long>(h[0]);
    counts_host[1] = static_cast<long long>(h[2]);
    return 0;
}

}  // namespace pifu
