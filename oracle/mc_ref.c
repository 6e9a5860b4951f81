/* CPU ORACLE (test infrastructure - NOT a product path): sequential marching cubes.
 *
 * Stands in for skimage.measure.marching_cubes_lewiner (call site mesh_util.py:84), which is
 * third-party code absent from this image ("parity unpinned": this file follows the behaviour
 * SURVEY.md §8(c) recalls, with the generated case tables of gen_mc_tables.py, and is what the
 * CUDA kernel is compared against bit for bit).
 *
 *   - volume [n0][n1][n2] float32, cells visited with axis 2 fastest, then axis 1, then axis 0
 *   - case bit i set iff corner i's value > level (strict)
 *   - a vertex is created the first time a cell's triangle list references its lattice edge
 *     (edge cache keyed by the edge's lower voxel + axis); numbering = creation order
 *   - position = weighted mean of the two corners, weights 1 / (FLT_EPSILON + |v - level|) in double
 *   - degenerate triangles are kept; faces in traversal order, table order within a cell
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MC_TABLE_QUALIFIER static const
#include "../rgb-d-pifuhd_b200/csrc/mc_tables.h"

static double* g_verts = 0;
static float* g_normals = 0;
static float* g_values = 0;
static int* g_faces = 0;
static long long g_nv = 0, g_nf = 0, g_cap_v = 0, g_cap_f = 0;
static unsigned char* g_cases = 0;
static long long g_ncells = 0;

static void push_vertex(const double* p, const float* n, float val) {
    if (g_nv == g_cap_v) {
        g_cap_v = g_cap_v ? g_cap_v * 2 : 4096;
        g_verts = (double*)realloc(g_verts, sizeof(double) * 3 * g_cap_v);
        g_normals = (float*)realloc(g_normals, sizeof(float) * 3 * g_cap_v);
        g_values = (float*)realloc(g_values, sizeof(float) * g_cap_v);
    }
    memcpy(g_verts + 3 * g_nv, p, 3 * sizeof(double));
    memcpy(g_normals + 3 * g_nv, n, 3 * sizeof(float));
    g_values[g_nv] = val;
    ++g_nv;
}

static void push_face(int a, int b, int c) {
    if (g_nf == g_cap_f) {
        g_cap_f = g_cap_f ? g_cap_f * 2 : 4096;
        g_faces = (int*)realloc(g_faces, sizeof(int) * 3 * g_cap_f);
    }
    g_faces[3 * g_nf] = a; g_faces[3 * g_nf + 1] = b; g_faces[3 * g_nf + 2] = c;
    ++g_nf;
}

/* n[] = global extents, off = global index of local plane 0 (slab form) */
static double grad(const float* f, const int* n, const long long* st, const int* p, int a, int off) {
    const long long i = p[0] * st[0] + p[1] * st[1] + p[2] * st[2];
    const int g = p[a] + (a == 0 ? off : 0);
    const int lo = g > 0 ? -1 : 0, hi = g < n[a] - 1 ? 1 : 0;
    return ((double)f[i + hi * st[a]] - (double)f[i + lo * st[a]]) / (double)(hi - lo);
}

static long long g_ghost_verts = 0;

/* Slab form (mirrors pifu_mc_count_slab, include/pifu_b200.h): f holds planes [i0, i0 + n0) of a
 * volume with g0 planes; cell layers [0, layers) are traversed; with ghost != 0 the first layer
 * belongs to the previous slab: it creates (numbers) only the vertices a traversal of the whole
 * volume would create in that layer - not those on its low face, which an earlier layer owns -
 * and emits no faces.  returns 0 on success; results are fetched with mc_ref_sizes / mc_ref_copy */
int mc_ref_run_slab(const float* f, int n0, int n1, int n2, double level, int i0, int g0, int layers, int ghost) {
    const int n[3] = {g0, n1, n2};
    const long long st[3] = {(long long)n1 * n2, n2, 1};
    const long long nvox = (long long)n0 * n1 * n2;
    int* cache = (int*)malloc(sizeof(int) * 3 * nvox);      /* edge (voxel, axis) -> vertex id */
    long long q;
    int i, j, k, c, t, a;
    if (!cache) return -1;
    for (q = 0; q < 3 * nvox; ++q) cache[q] = -1;
    g_nv = g_nf = 0;
    g_ghost_verts = 0;
    g_ncells = (long long)layers * (n1 - 1) * (n2 - 1);
    g_cases = (unsigned char*)realloc(g_cases, g_ncells > 0 ? g_ncells : 1);
    q = 0;
    for (i = 0; i < layers; ++i) {
        if (ghost && i == 1) g_ghost_verts = g_nv;
        for (j = 0; j < n1 - 1; ++j)
            for (k = 0; k < n2 - 1; ++k, ++q) {
                int cs = 0;
                for (c = 0; c < 8; ++c) {
                    const long long v = (i + MC_CORNER[c][0]) * st[0] + (j + MC_CORNER[c][1]) * st[1] + (k + MC_CORNER[c][2]);
                    if ((double)f[v] > level) cs |= 1 << c;
                }
                g_cases[q] = (unsigned char)cs;
                for (t = 0; t < MC_NTRI[cs]; ++t) {
                    int vid[3];
                    for (c = 0; c < 3; ++c) {
                        const int e = MC_TRIS[cs][3 * t + c];
                        const int ca = MC_EDGE_CORNERS[e][0], cb = MC_EDGE_CORNERS[e][1];
                        int pa[3], pb[3], lo[3];
                        long long key;
                        pa[0] = i + MC_CORNER[ca][0]; pa[1] = j + MC_CORNER[ca][1]; pa[2] = k + MC_CORNER[ca][2];
                        pb[0] = i + MC_CORNER[cb][0]; pb[1] = j + MC_CORNER[cb][1]; pb[2] = k + MC_CORNER[cb][2];
                        for (a = 0; a < 3; ++a) lo[a] = pa[a] < pb[a] ? pa[a] : pb[a];
                        key = 3 * (lo[0] * st[0] + lo[1] * st[1] + lo[2]) + MC_EDGE_AXIS[e];
                        if (ghost && i == 0 && pa[0] == 0 && pb[0] == 0) { vid[c] = -1; continue; }   /* owned by the slab before */
                        if (cache[key] < 0) {
                            const double va = (double)f[pa[0] * st[0] + pa[1] * st[1] + pa[2]];
                            const double vb = (double)f[pb[0] * st[0] + pb[1] * st[1] + pb[2]];
                            const double fa = 1.0 / ((double)FLT_EPSILON + fabs(va - level));
                            const double fb = 1.0 / ((double)FLT_EPSILON + fabs(vb - level));
                            const double fs = fa + fb;
                            double p[3], g[3], len;
                            float nr[3];
                            for (a = 0; a < 3; ++a) {
                                const int o = a == 0 ? i0 : 0;
                                const double pa_w = (double)(pa[a] + o) * fa, pb_w = (double)(pb[a] + o) * fb;
                                const double ga_w = grad(f, n, st, pa, a, i0) * fa, gb_w = grad(f, n, st, pb, a, i0) * fb;
                                p[a] = (pa_w + pb_w) / fs;
                                g[a] = (ga_w + gb_w) / fs;
                            }
                            {
                                const double g0 = g[0] * g[0], g1 = g[1] * g[1], g2 = g[2] * g[2];
                                len = sqrt((g0 + g1) + g2);
                            }
                            for (a = 0; a < 3; ++a) nr[a] = len > 0.0 ? (float)(-g[a] / len) : 0.f;
                            cache[key] = (int)g_nv;
                            push_vertex(p, nr, (float)(va > vb ? va : vb));
                        }
                        vid[c] = cache[key];
                    }
                    if (!(ghost && i == 0)) push_face(vid[0], vid[1], vid[2]);
                }
            }
    }
    free(cache);
    (void)n0;
    return 0;
}

int mc_ref_run(const float* f, int n0, int n1, int n2, double level) {
    return mc_ref_run_slab(f, n0, n1, n2, level, 0, n0, n0 - 1, 0);
}

long long mc_ref_ghost_verts(void) { return g_ghost_verts; }

void mc_ref_sizes(long long* nv, long long* nf, long long* ncells) { *nv = g_nv; *nf = g_nf; *ncells = g_ncells; }

void mc_ref_copy(double* verts, int* faces, float* normals, float* values, unsigned char* cases) {
    if (verts) memcpy(verts, g_verts, sizeof(double) * 3 * g_nv);
    if (faces) memcpy(faces, g_faces, sizeof(int) * 3 * g_nf);
    if (normals) memcpy(normals, g_normals, sizeof(float) * 3 * g_nv);
    if (values) memcpy(values, g_values, sizeof(float) * g_nv);
    if (cases) memcpy(cases, g_cases, g_ncells);
}
