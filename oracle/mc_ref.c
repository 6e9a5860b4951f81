/* CPU ORACLE (test infrastructure - NOT a product path): sequential marching cubes.
 *
 * Stands in for skimage.measure.marching_cubes_lewiner (call site mesh_util.py:84), which is third-party code
 * absent from this image ("parity unpinned").  It is INDEPENDENT of the product's lookup tables: it includes no
 * table and derives every cell's triangles at run time, from the eight corner values, by the rule stated in
 * rgb-d-pifuhd_b200/tools/gen_mc_tables.py ("THE RULE"): face segments with Lewiner's face test (asymptotic
 * decider) on ambiguous faces, loops from the lowest edge id oriented towards decreasing values, fan
 * triangulation unless a diagonal lies inside a cube face.  The CUDA kernels read the generated tables; a wrong
 * table entry, a wrong sub-case index or a wrong decider fails the bit-exact GPU-vs-oracle tests.
 *
 *   - volume [n0][n1][n2] float32, cells visited with axis 2 fastest, then axis 1, then axis 0
 *   - case bit i set iff corner i's value > level (strict)
 *   - a vertex is created the first time a cell's triangle list references its lattice edge
 *     (edge cache keyed by the edge's lower voxel + axis); numbering = creation order
 *   - position = weighted mean of the two corners, weights 1 / (FLT_EPSILON + |v - level|) in double
 *   - degenerate triangles are kept; faces in traversal order, rule order within a cell
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- cube conventions (the only constants shared with the generator, restated here) ---- */
/* corner offsets along (axis0, axis1, axis2): corner c of the rule sits at (d2, d1, d0) = ... reversed */
static const unsigned char CORNER[8][3] = {{0,0,0},{0,0,1},{0,1,1},{0,1,0},{1,0,0},{1,0,1},{1,1,1},{1,1,0}};
static const unsigned char EDGE_CORNERS[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};
static const unsigned char FACE[6][4] = {{0,1,2,3},{4,5,6,7},{0,1,5,4},{3,2,6,7},{0,3,7,4},{1,2,6,5}};

static int edge_between(int a, int b) {
    int e;
    for (e = 0; e < 12; ++e)
        if ((EDGE_CORNERS[e][0] == a && EDGE_CORNERS[e][1] == b) || (EDGE_CORNERS[e][0] == b && EDGE_CORNERS[e][1] == a)) return e;
    return -1;
}

static int edge_axis(int e) {
    const unsigned char* a = CORNER[EDGE_CORNERS[e][0]];
    const unsigned char* b = CORNER[EDGE_CORNERS[e][1]];
    return a[0] != b[0] ? 0 : (a[1] != b[1] ? 1 : 2);
}

static int edges_share_face(int e1, int e2) {
    int f, k;
    for (f = 0; f < 6; ++f) {
        int h1 = 0, h2 = 0;
        for (k = 0; k < 4; ++k) {
            const int e = edge_between(FACE[f][k], FACE[f][(k + 1) & 3]);
            if (e == e1) h1 = 1;
            if (e == e2) h2 = 1;
        }
        if (h1 && h2) return f + 1;
    }
    return 0;
}

/* ---- triangulations of an index polygon lo..hi, in the rule's enumeration order ---- */
typedef struct { unsigned char n; unsigned char t[10][3]; } TriList;
typedef struct { int count; TriList* lists; } TriSet;
static TriSet g_memo[13][13];
static int g_memo_done[13][13];

static const TriSet* tri_set(int lo, int hi) {
    TriSet* s = &g_memo[lo][hi];
    int m, i, j;
    if (g_memo_done[lo][hi]) return s;
    g_memo_done[lo][hi] = 1;
    if (hi - lo + 1 < 3) {
        s->count = 1;
        s->lists = (TriList*)calloc(1, sizeof(TriList));
        return s;
    }
    if (hi - lo + 1 == 3) {
        s->count = 1;
        s->lists = (TriList*)calloc(1, sizeof(TriList));
        s->lists[0].n = 1;
        s->lists[0].t[0][0] = (unsigned char)lo; s->lists[0].t[0][1] = (unsigned char)(lo + 1); s->lists[0].t[0][2] = (unsigned char)hi;
        return s;
    }
    {
        int total = 0, at = 0;
        for (m = lo + 1; m < hi; ++m) total += tri_set(lo, m)->count * tri_set(m, hi)->count;
        s->lists = (TriList*)calloc((size_t)total, sizeof(TriList));
        for (m = lo + 1; m < hi; ++m) {
            const TriSet* L = tri_set(lo, m);
            const TriSet* R = tri_set(m, hi);
            for (i = 0; i < L->count; ++i)
                for (j = 0; j < R->count; ++j) {
                    TriList* o = &s->lists[at++];
                    int q;
                    o->n = 0;
                    for (q = 0; q < L->lists[i].n; ++q) memcpy(o->t[o->n++], L->lists[i].t[q], 3);
                    o->t[o->n][0] = (unsigned char)lo; o->t[o->n][1] = (unsigned char)m; o->t[o->n][2] = (unsigned char)hi; ++o->n;
                    for (q = 0; q < R->lists[j].n; ++q) memcpy(o->t[o->n++], R->lists[j].t[q], 3);
                }
        }
        s->count = total;
    }
    return s;
}

/* diagonals (triangle sides that are not loop sides) lying inside a cube face, counted per triangle side */
static int bad_diagonals(const int* loop, int n, const unsigned char (*t)[3], int nt) {
    int q, s, bad = 0;
    for (q = 0; q < nt; ++q)
        for (s = 0; s < 3; ++s) {
            const int ia = t[q][s], ib = t[q][(s + 1) % 3];
            const int d = ia > ib ? ia - ib : ib - ia;
            if (d == 1 || d == n - 1) continue;                  /* a side of the loop */
            if (edges_share_face(loop[ia], loop[ib])) ++bad;
        }
    return bad;
}

/* the rule, for one cell: corner values v[8] (as double), inside bits cs -> triangle edge ids out[3 * ntri] */
static int cell_triangles(const double* v, int cs, double level, int* out) {
    int adj[12][2], deg[12];
    int seen[12];
    int f, k, e, ntri = 0;
    for (e = 0; e < 12; ++e) { deg[e] = 0; seen[e] = 0; }
    for (f = 0; f < 6; ++f) {
        int in[4], cut[4], ncut = 0;
        for (k = 0; k < 4; ++k) in[k] = (cs >> FACE[f][k]) & 1;
        for (k = 0; k < 4; ++k) {
            cut[k] = in[k] != in[(k + 1) & 3] ? edge_between(FACE[f][k], FACE[f][(k + 1) & 3]) : -1;
            if (cut[k] >= 0) ++ncut;
        }
        if (ncut == 2) {
            int a = -1, b = -1;
            for (k = 0; k < 4; ++k) if (cut[k] >= 0) { if (a < 0) a = cut[k]; else b = cut[k]; }
            adj[a][deg[a]++] = b;
            adj[b][deg[b]++] = a;
        } else if (ncut == 4) {
            /* Lewiner's face test: inside corners joined across the face iff the product of the inside corners'
             * (value - level) exceeds that of the outside corners' */
            const int i0 = in[0] ? 0 : 1;
            const double pin = (v[FACE[f][i0]] - level) * (v[FACE[f][i0 + 2]] - level);
            const double pout = (v[FACE[f][i0 ^ 1]] - level) * (v[FACE[f][(i0 ^ 1) + 2]] - level);
            const int joined = pin > pout;
            for (k = 0; k < 4; ++k)
                if ((in[k] != 0) != (joined != 0)) {             /* corner k is cut off: segment (edge k-1, edge k) */
                    const int a = cut[(k + 3) & 3], b = cut[k];
                    adj[a][deg[a]++] = b;
                    adj[b][deg[b]++] = a;
                }
        }
    }
    for (e = 0; e < 12; ++e) {
        int loop[12], n = 0, prev = -1, cur = e;
        if (deg[e] != 2 || seen[e]) continue;
        for (;;) {
            int nx;
            loop[n++] = cur;
            seen[cur] = 1;
            nx = (prev < 0 || adj[cur][0] != prev) ? adj[cur][0] : adj[cur][1];
            if (nx == e && n > 1) break;
            prev = cur;
            cur = nx;
        }
        {
            /* orientation on the first segment: see gen_mc_tables.py orient() */
            const int p = loop[0], q = loop[1];
            const int f1 = edges_share_face(p, q) - 1;
            double cen[3] = {0, 0, 0}, nrm[3], mp[3], mq[3], d[3], w[3], cr[3], s;
            int a = EDGE_CORNERS[p][0], b = EDGE_CORNERS[p][1], x;
            for (k = 0; k < 4; ++k) for (x = 0; x < 3; ++x) cen[x] += CORNER[FACE[f1][k]][x] / 4.0;
            for (x = 0; x < 3; ++x) nrm[x] = -2.0 * (cen[x] - 0.5);
            for (x = 0; x < 3; ++x) {
                mp[x] = (CORNER[EDGE_CORNERS[p][0]][x] + CORNER[EDGE_CORNERS[p][1]][x]) / 2.0;
                mq[x] = (CORNER[EDGE_CORNERS[q][0]][x] + CORNER[EDGE_CORNERS[q][1]][x]) / 2.0;
                d[x] = mq[x] - mp[x];
            }
            if ((cs >> a) & 1) { const int t = a; a = b; b = t; }         /* a outside, b inside */
            for (x = 0; x < 3; ++x) w[x] = (double)CORNER[a][x] - (double)CORNER[b][x];
            cr[0] = d[1] * nrm[2] - d[2] * nrm[1];
            cr[1] = d[2] * nrm[0] - d[0] * nrm[2];
            cr[2] = d[0] * nrm[1] - d[1] * nrm[0];
            s = cr[0] * w[0] + cr[1] * w[1] + cr[2] * w[2];
            if (s < 0) {
                int lo = 1, hi = n - 1;
                while (lo < hi) { const int t = loop[lo]; loop[lo] = loop[hi]; loop[hi] = t; ++lo; --hi; }
            }
        }
        {
            unsigned char fan[10][3];
            const unsigned char (*pick)[3] = fan;
            int nt = n - 2, q;
            for (q = 0; q < nt; ++q) { fan[q][0] = 0; fan[q][1] = (unsigned char)(q + 1); fan[q][2] = (unsigned char)(q + 2); }
            if (bad_diagonals(loop, n, fan, nt) != 0) {
                const TriSet* all = tri_set(0, n - 1);
                int best = -1, best_bad = 1 << 30, i;
                for (i = 0; i < all->count; ++i) {
                    const int bd = bad_diagonals(loop, n, all->lists[i].t, all->lists[i].n);
                    if (bd < best_bad) { best_bad = bd; best = i; if (bd == 0) break; }
                }
                pick = all->lists[best].t;
                nt = all->lists[best].n;
            }
            for (q = 0; q < nt; ++q) {
                out[3 * ntri] = loop[pick[q][0]]; out[3 * ntri + 1] = loop[pick[q][1]]; out[3 * ntri + 2] = loop[pick[q][2]];
                ++ntri;
            }
        }
    }
    return ntri;
}

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
    /* p[a] > 0: the slab's first plane has nothing below it in memory; only vertices of the ghost layer (dropped by the
     * caller) can ask for it */
    const int lo = (g > 0 && p[a] > 0) ? -1 : 0, hi = g < n[a] - 1 ? 1 : 0;
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
                int cs = 0, ntri, tris[30];
                double cv[8];
                for (c = 0; c < 8; ++c) {
                    const long long v = (i + CORNER[c][0]) * st[0] + (j + CORNER[c][1]) * st[1] + (k + CORNER[c][2]);
                    cv[c] = (double)f[v];
                    if (cv[c] > level) cs |= 1 << c;
                }
                g_cases[q] = (unsigned char)cs;
                if (cs == 0 || cs == 255) continue;
                ntri = cell_triangles(cv, cs, level, tris);
                for (t = 0; t < ntri; ++t) {
                    int vid[3];
                    for (c = 0; c < 3; ++c) {
                        const int e = tris[3 * t + c];
                        const int ca = EDGE_CORNERS[e][0], cb = EDGE_CORNERS[e][1];
                        int pa[3], pb[3], lo[3];
                        long long key;
                        pa[0] = i + CORNER[ca][0]; pa[1] = j + CORNER[ca][1]; pa[2] = k + CORNER[ca][2];
                        pb[0] = i + CORNER[cb][0]; pb[1] = j + CORNER[cb][1]; pb[2] = k + CORNER[cb][2];
                        for (a = 0; a < 3; ++a) lo[a] = pa[a] < pb[a] ? pa[a] : pb[a];
                        key = 3 * (lo[0] * st[0] + lo[1] * st[1] + lo[2]) + edge_axis(e);
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

/* test hook: the rule for one cell, from its eight corner values -> triangle edge ids; returns the triangle count */
int mc_ref_cell(const double* corner_values, double level, int* tris_out) {
    int c, cs = 0;
    for (c = 0; c < 8; ++c) if (corner_values[c] > level) cs |= 1 << c;
    if (cs == 0 || cs == 255) return 0;
    return cell_triangles(corner_values, cs, level, tris_out);
}
