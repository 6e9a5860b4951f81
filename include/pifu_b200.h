/* pifu_b200.h - C ABI of libpifu_b200.so: the B200-native reconstruction hot path of
 * RGB-D-PIFuHD (occupancy query on explicit points / dense lattice / octree, marching cubes).
 *
 * The reference has no FFI or operator registry (it is pure Python); its boundary for this
 * path is the Python object API.  Each entry point below names the reference interface it
 * replaces (file:line under the reference tree).  The Python layer in
 * `rgb-d-pifuhd_b200/` binds these with ctypes and keeps the reference's signatures
 * (see INTEGRATION.md).
 *
 * Conventions: every function returns 0 on success and -1 on failure (never throws);
 * `pifu_last_error()` gives the message.  "device pointer" = CUDA device memory on the
 * context's device; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * A context owns one workspace: calls on the same context must be issued from one thread
 * and are ordered on the stream they are given.
 */
#ifndef PIFU_B200_H
#define PIFU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pifu_ctx pifu_ctx;

#define PIFU_LEVEL_COARSE 0 /* PIFuNetwNML  (PIFuNetwNML.py:17-71)  */
#define PIFU_LEVEL_FINE 1   /* PIFuMRNet    (PIFuMRNet.py:14-57)    */

#define PIFU_GEMM_TCGEN05 0 /* tensor-core layer kernel, CTA pairs / cta_group::2 (default, product path) */
#define PIFU_GEMM_SIMT 1    /* CUDA-core cross-check of the same operands (tests only) */
#define PIFU_GEMM_TCGEN05_1CTA 2 /* tensor-core layer kernel without CTA pairs (A/B measurements) */

const char* pifu_last_error(void);
int pifu_abi_version(void);

/* One context per device. */
int pifu_create(int device, pifu_ctx** out);
void pifu_destroy(pifu_ctx* ctx);

/* Snapshot one MLP (replaces reading `net.mlp` - MLP.py:13-40 - at query time).
 * filter_channels[n_channels], res_layers[n_res], merge_layer as given to MLP.__init__
 * (<= 0 means len(filter_channels)//2, MLP.py:25).  weights[i] / biases[i] are device
 * pointers to Conv1d i's fp32 weight [Cout][Cin] and bias [Cout].  The stack is un-normalised
 * (mlp_norm == 'none', MLP.py:66-67) until pifu_set_mlp_norm is called; hidden widths must be
 * multiples of 128 and the last hidden width 128 or 256. */
int pifu_set_mlp(pifu_ctx* ctx, int level, int n_channels, const int* filter_channels, int n_res,
                 const int* res_layers, int merge_layer, const float* const* weights,
                 const float* const* biases, void* stream);

/* Normalisation between each hidden Conv1d and its leaky_relu (MLP.py:36-41,66-69), to be set
 * after pifu_set_mlp (which resets it to none).  groups = 32: GroupNorm(32, C); groups = 0: one
 * group per channel = BatchNorm1d with batch statistics (a module left in train mode, as
 * reconstruction.py:288-289 leaves the fine net).  gammas[i] / betas[i]: device fp32 [Cout_i] for the
 * hidden layers i = 0 .. n_channels - 3.  The statistics run over all points of one call
 * ([1, C, N] in the reference), so with a normalised MLP every pifu_query / pifu_eval_grid /
 * pifu_eval_lattice_ids call is one statistics domain (at most 2^22 points, never chunked), the
 * lattice chain kernel is not used, and pifu_eval_grid_octree (one-call form) is refused: the
 * caller splits the work the way the reference's num_samples does.  BatchNorm1d in eval mode is a
 * per-channel affine map and is folded into the weights by the Python layer instead. */
int pifu_set_mlp_norm(pifu_ctx* ctx, int level, int groups, double eps, const float* const* gammas,
                      const float* const* betas, void* stream);

/* Snapshot a feature map (replaces reading `net.im_feat_list[-1]`, PIFuNetwNML.py:94-97 /
 * PIFuMRNet.py:114-117): device fp32 NCHW [1][C][H][W]. */
int pifu_set_features(pifu_ctx* ctx, int level, const float* nchw, int C, int H, int W, void* stream);

/* projection: 0 orthogonal / 1 perspective (BasePIFuNet.py:79); DepthNormalizer constants
 * z * z_mul / z_div (DepthNormalizer.py:23: loadSize // 2 and z_size). */
int pifu_set_options(pifu_ctx* ctx, int perspective, float z_mul, float z_div);

/* Arithmetic of the per-point MLP (MLP.py:55-73; the reference computes in fp32).
 *   PIFU_PREC_FAST   (default) every tensor-core operand is one fp16 image, fp32 accumulate: logit error ~6e-4 of the
 *                    logit spread;
 *   PIFU_PREC_SPLIT  every operand (features, activations, weights) is carried as fp16 + fp16 residual and a layer
 *                    computes x_hi W_hi + x_lo W_hi + x_hi W_lo (three tensor-core passes, fp32 accumulate): fp32-level
 *                    error, per-layer kernels only; `terms` selects the residual products (bit 0: activations and
 *                    features, bit 1: weights; 3 = both) so each rounding point's share of the error can be measured;
 *   PIFU_PREC_HYBRID fast everywhere (chain kernels included), then the points whose occupancy falls inside
 *                    (band_lo, band_hi) are evaluated again in split precision: outside the band the sigmoid's slope
 *                    <= band_lo (1 - band_lo) scales the fast path's logit error below the parity gate and no sign at
 *                    0.5 can differ.  One host synchronisation per 4 Mi outputs sizes the second pass. */
#define PIFU_PREC_FAST 0
#define PIFU_PREC_SPLIT 1
#define PIFU_PREC_HYBRID 2
int pifu_set_precision(pifu_ctx* ctx, int mode, int terms, float band_lo, float band_hi);
/* points re-evaluated by the hybrid mode since the context was created */
long long pifu_refined_points(pifu_ctx* ctx);
int pifu_set_gemm_impl(pifu_ctx* ctx, int impl);
int pifu_set_chunk_tiles(pifu_ctx* ctx, int tiles_of_128_points);

/* Occupancy of explicit points.  Replaces PIFuNetwNML.query (PIFuNetwNML.py:99-141; levels = 1)
 * and PIFuMRNet.query (PIFuMRNet.py:119-186; levels = 2) for one view/crop.
 * points: device fp32, reference layout [3][n] with row stride `pstride` elements.
 * calib_local / calib_global: host, 16 floats row-major (4x4).  For levels == 1 both must be
 * the same matrix.  Outputs are device pointers, any may be NULL:
 *   out_pred     [n]      preds (mask * sigmoid), BasePIFuNet.get_preds (BasePIFuNet.py:136-142)
 *   out_pred_low [n]      coarse prediction (netG.intermediate_preds_list[-1]); levels == 2 only
 *   out_phi      [C][n]   coarse merge-layer feature `netG.phi` (row stride n) */
#define PIFU_QUERY_NO_MASK 1 /* raw sigmoid, no in-bounds mask: calc_normal (PIFuMRNet.py:232-237) */
#define PIFU_QUERY_PRECISE 2 /* this call in split precision whatever pifu_set_precision says (finite differences) */
int pifu_query(pifu_ctx* ctx, int levels, int flags, const float* points, long long pstride, long long n,
               const float* calib_local, const float* calib_global, float* out_pred,
               float* out_pred_low, float* out_phi, void* stream);

/* Occupancy on a run of lattice points, never materialising coordinates.  Replaces
 * create_grid + the calib pre-transform + eval_grid/batch_eval (mesh_util.py:12-38, 59-65,
 * 98-120).  Lattice point id = (i*R1 + j)*R2 + k (C order of the reference's sdf volume);
 * ids [id_begin, id_end) are evaluated into out[0 .. id_end-id_begin) (device fp32).
 * calib: host 16 floats; calib_inv: host 16 doubles = numpy.linalg.inv(calib) as the
 * reference computes it on the host (mesh_util.py:62). */
int pifu_eval_grid(pifu_ctx* ctx, int levels, int R0, int R1, int R2, long long id_begin,
                   long long id_end, const float* calib, const double* calib_inv, float* out,
                   void* stream);

/* Same, for an explicit list of lattice ids (device int64 [n]); out[p] = occupancy of ids[p].  Sorted
 * lists (octree frontiers) are the fast case: see the chain kernel's run-list form below. */
int pifu_eval_lattice_ids(pifu_ctx* ctx, int levels, int R0, int R1, int R2, const long long* ids,
                          long long n, const float* calib, const double* calib_inv, float* out,
                          void* stream);

/* Octree (coarse-to-fine) lattice evaluation on the device.  Replaces eval_grid_octree
 * (mesh_util.py:124-187): float64 field, last plane of each axis never evaluated (:135),
 * stride R0 // init_resolution halving to 1 (:138-185), skip test (max - min) < threshold on
 * the 8 cell corners and inclusive fill (:154-184).
 * One-call form: sdf64 (device double [R0*R1*R2]) and/or sdf32 (device float, the cast
 * scikit-image applies on entry) may be NULL; evaluated_per_level (host, max_levels entries,
 * -1 padded) receives the number of lattice points evaluated at each level. */
int pifu_eval_grid_octree(pifu_ctx* ctx, int levels, int R0, int R1, int R2, int init_resolution,
                          double threshold, const float* calib, const double* calib_inv, double* sdf64,
                          float* sdf32, long long* evaluated_per_level, int max_levels, void* stream);

/* Stepwise form, so a multi-GPU driver can split each level's frontier across ranks:
 *   begin -> { frontier (compacted lattice ids of this level, C order) -> [evaluate ids,
 *   e.g. pifu_eval_lattice_ids on a share + all-gather] -> commit(values of the whole
 *   frontier) } until frontier reports step == 0 -> export. */
int pifu_octree_begin(pifu_ctx* ctx, int R0, int R1, int R2, int init_resolution, double threshold,
                      void* stream);
int pifu_octree_frontier(pifu_ctx* ctx, long long* n, const long long** ids_device, int* step, void* stream);
int pifu_octree_commit(pifu_ctx* ctx, const float* values_device, void* stream);
/* Same with float64 values: the generic eval_grid_octree(coords, eval_func) form stores whatever
 * the caller's eval_func returns into the float64 field (mesh_util.py:148-149). */
int pifu_octree_commit64(pifu_ctx* ctx, const double* values_device, void* stream);
int pifu_octree_export(pifu_ctx* ctx, double* sdf64, float* sdf32, void* stream);

/* Slab form of the stepwise octree for a volume sharded along axis 0 (nothing to cite in the
 * single-device reference beyond mesh_util.py:124-187).  The rank keeps the bookkeeping of planes
 * [plane_begin, plane_end) of the R0-plane volume - its own planes [own_begin, own_end) plus a margin of twice the
 * initial stride (R0 // init_resolution) on either side, clipped to the volume; all four are multiples of the initial
 * stride (or R0).  pifu_octree_frontier then compacts the frontier of the OWN planes only (global lattice ids), and
 * each level is closed with pifu_octree_commit_pairs: the (lattice id, value) pairs of every rank's frontier (any
 * order; ids < 0 are padding); pairs outside [plane_begin, plane_end) are ignored.  No boundary plane is exchanged -
 * what the cells beyond the margin would contribute cannot reach the own planes, nor the two planes either side of
 * them that slab marching cubes reads (octree.cu).  pifu_octree_field32 gives the float32 field of the local planes
 * (library-owned device memory, valid until the next begin): *plane_begin = global index of its first plane. */
int pifu_octree_begin_slab(pifu_ctx* ctx, int R0, int R1, int R2, int init_resolution, double threshold, int plane_begin,
                           int plane_end, int own_begin, int own_end, void* stream);
int pifu_octree_commit_pairs(pifu_ctx* ctx, const long long* ids_device, const float* values_device, long long n, void* stream);
/* Planes (global indices, inside the local planes, starting on the current stride) whose frontier the next
 * pifu_octree_frontier compacts.  A rank may compact - and evaluate itself - the frontier of ALL its local planes at the
 * coarse levels, whose frontiers are small: the values are the ones its neighbours compute (a point's value does not
 * depend on the call it is evaluated in), so those levels need no communication at all; pifu_octree_commit then takes
 * the values in frontier order as in the single-device form. */
int pifu_octree_set_frontier_planes(pifu_ctx* ctx, int plane_begin, int plane_end);
int pifu_octree_field32(pifu_ctx* ctx, const float** field_device, int* plane_begin, int* planes);

/* Marching cubes on a device float32 volume [n0][n1][n2] at `level` (strict v > level is
 * inside).  Replaces measure.marching_cubes_lewiner (call site mesh_util.py:84; third-party,
 * see DESIGN.md "parity unpinned").  Ambiguous faces are resolved by Lewiner's face test (the asymptotic decider);
 * interior (tunnel) tests and the centre vertex of Lewiner's 33 cases are not implemented - INTEGRATION.md lists
 * the deviations.  Two calls because the output size is data dependent:
 * count (synchronises, returns sizes) then emit into caller-allocated device buffers:
 * verts double [nverts][3] in volume-index coordinates (axis order of the volume), faces
 * int32 [nfaces][3], optional normals float [nverts][3] and values float [nverts]. */
int pifu_mc_count(pifu_ctx* ctx, const float* field, int n0, int n1, int n2, double level,
                  long long* nverts, long long* nfaces, void* stream);
int pifu_mc_emit(pifu_ctx* ctx, double* verts, int* faces, float* normals, float* values, void* stream);

/* Slab form of pifu_mc_count for a volume sharded along axis 0 (the reference is
 * single-device, there is nothing to cite beyond mesh_util.py:84).  `field` holds planes
 * [i_global0, i_global0 + n0) of a volume with global_n0 planes; the cell layers [0, cell_layers)
 * of it are processed (extra planes above them only feed the normals).  Vertex positions, border
 * handling and vertex ownership use global plane indices, so the slab's vertices and faces are
 * exactly those a traversal of the whole volume creates in these layers.  With ghost_layers == 1
 * the first cell layer belongs to the previous slab: it is classified and numbers its vertices
 * (the next layer's triangles refer to them) but emits no faces; *ghost_verts returns how many
 * leading vertices it numbered, so that for a slab whose first own vertex has global number G
 *   global vertex id = local id - *ghost_verts + G
 * and the caller drops the first *ghost_verts vertices (they are numbered but NOT written: the slab before computes
 * them, and their normals would need a plane this slab does not hold).  pifu_mc_emit is unchanged. */
int pifu_mc_count_slab(pifu_ctx* ctx, const float* field, int n0, int n1, int n2, double level,
                       int i_global0, int global_n0, int cell_layers, int ghost_layers,
                       long long* nverts, long long* nfaces, long long* ghost_verts, void* stream);

/* One-call, fully asynchronous form of count + emit (single volume: i_global0 = 0, global_n0 = n0, cell_layers = n0 - 1,
 * ghost_layers = 0; or a slab as above): classify, scan and emission are queued on `stream` without a host
 * synchronisation in between.  The outputs are caller-allocated with capacities cap_verts / cap_faces (rows);
 * counts_device (device, 3 x uint64) receives the true numbers of vertices, faces and ghost-layer vertices.  When a
 * count exceeds its capacity nothing is written to that array: the caller reads the counts (its only synchronisation,
 * typically together with the transfer of the mesh) and repeats the call with larger buffers. */
int pifu_mc_extract(pifu_ctx* ctx, const float* field, int n0, int n1, int n2, double level, int i_global0, int global_n0,
                    int cell_layers, int ghost_layers, double* verts, int* faces, float* normals, float* values,
                    long long cap_verts, long long cap_faces, unsigned long long* counts_device, void* stream);

/* Host-side OBJ writer (no GPU involved).  Replaces save_obj_mesh_with_color (mesh_util.py:189-198):
 * "v %.4f %.4f %.4f %.4f %.4f %.4f" per vertex (host double verts [nverts][3], colors [nverts][3]) then
 * "f %d %d %d" per face (host int32 faces [nfaces][3], written 1-based as f0, f2, f1); the text is
 * byte-identical to the reference's. */
int pifu_write_obj(const char* path, const double* verts, const double* colors, long long nverts,
                   const int* faces, long long nfaces);

/* Host-side OBJ reader, the inverse of pifu_write_obj: what meshcleaning (reconstruction.py:325-344, `trimesh.load`) needs
 * of the file the pipeline wrote.  pifu_obj_counts: counts (HOST, 3 entries) = "v " lines, "f " lines, values on the first
 * vertex line (3, or 6 with colours).  pifu_read_obj: verts [nverts][3], colors [nverts][3] or NULL, faces [nfaces][3]
 * 0-based in the order they were GIVEN to the writer (the file stores f0, f2, f1); "a/b/c" references keep the vertex
 * index; every other line is skipped.  -1 when the counts differ from the file's, a line is malformed, or colours are
 * asked of a file without them. */
int pifu_obj_counts(const char* path, long long* counts);
int pifu_read_obj(const char* path, double* verts, double* colors, int* faces, long long nverts, long long nfaces);

/* Element-wise helper of the PyTorch image encoders (no context): y = relu?((x - mean) / sqrt(var + eps) * weight
 * + bias) per channel, the eval-mode BatchNorm2d + ReLU pairs of Filter.py:23-69 in one pass.  x, y: device fp32
 * [N][C][HW] contiguous (y may alias x); statistics / affine parameters: device fp32 [C] (weight / bias may be null).
 * Launches on the CURRENT device. */
int pifu_bn_relu_f32(const float* x, const float* running_mean, const float* running_var, const float* weight,
                     const float* bias, double eps, int relu, float* y, long long N, int C, long long HW, void* stream);

/* out[n] = concatenate(a[n], b[n], c[n]) + s[n]: the tail of a hourglass block (Filter.py:65-67) in one pass.  Device
 * fp32, contiguous, 16-byte aligned; la / lb / lc = floats of ONE image in a / b / c (multiples of 4); s and out hold
 * la + lb + lc floats per image.  Launches on the CURRENT device. */
int pifu_cat3_add_f32(const float* a, const float* b, const float* c, const float* s, float* out, long long N,
                      long long la, long long lb, long long lc, void* stream);

/* Vertex colours from an image (no context; launches on the CURRENT device).  Replaces, for one view,
 * `xyz = net.projection(verts, calib); color = index(image, xyz[:, :2])` of gen_mesh_imgColor (reconstruction.py:110-116;
 * BasePIFuNet.py:11-65): points device fp32 [3][n] (row stride pstride) are projected with calib (host, 16 floats) -
 * orthogonal, or perspective (xy / z) - and the device fp32 image [C][H][W] is sampled bilinearly (align_corners=True,
 * zeros outside, the arithmetic of the query kernels).  out: device fp32 [C][n]. */
int pifu_sample_image(const float* image_nchw, int C, int H, int W, const float* points, long long pstride, long long n,
                      const float* calib, int perspective, float* out, void* stream);

/* Largest connected component of a triangle mesh (no context; CURRENT device; synchronous).  Replaces meshcleaning
 * (reconstruction.py:325-344: trimesh.load(path).split(), keep the component with the greatest extent along axis 0).
 * trimesh is third-party and absent here; its behaviour for this call is restated: faces sharing an edge that exactly two
 * faces use are adjacent; with only_watertight (what split() defaults to) a component counts only if every one of its
 * edges is used by exactly two faces and it has at least 4 faces; components are ordered by their first face; the first
 * one of maximal extent wins.  All pointers are device memory: verts double [nverts][3], colors double [nverts][3] or
 * NULL, faces int32 [nfaces][3]; outputs have the capacity of the inputs and receive the kept vertices (original order)
 * and the renumbered faces; counts (HOST, 2 entries) receives the kept vertex and face counts.  -1 with an error
 * message when no component qualifies (trimesh: `cc[0]` raises). */
int pifu_mesh_clean(const double* verts, const double* colors, const int* faces, long long nverts, long long nfaces,
                    int only_watertight, double* out_verts, double* out_colors, int* out_faces, long long* counts, void* stream);

/* Number of kernels launched by this context since creation (bench accounting). */
long long pifu_launch_count(pifu_ctx* ctx);

/* Per-launch CUDA-event timing of the tensor-core layer kernel (bench.py's roofline leg):
 * enable, run the workload, read (synchronises): launches, summed device time and summed
 * algorithmic FLOPs (2 * points * Cin * Cout per launch, un-padded Cin). */
int pifu_profile_enable(pifu_ctx* ctx, int on);
int pifu_profile_read(pifu_ctx* ctx, long long* launches, double* total_ms, double* total_flops);

/* Same, restricted to one kernel kind (0 = per-layer tcgen05 kernel, 1 = chain kernel, lattice form, 2 = chain kernel, run-list form);
 * does not clear the records. */
int pifu_profile_read_kind(pifu_ctx* ctx, int kind, long long* launches, double* total_ms, double* total_flops);

/* Chain kernel (whole MLP stack of a 128-point tile in one kernel, activations resident in
 * shared/tensor memory).  Two forms, both taken automatically when both MLPs have the reference
 * configuration (options.py:86-87,92-93), the projection is orthogonal and the calibration does not
 * mix z into x/y:
 *   lattice form   pifu_eval_grid, R2 a multiple of 128: a tile is 128 consecutive points of one
 *                  lattice column;
 *   run-list form  pifu_eval_lattice_ids / pifu_eval_grid_octree: a tile is 128 consecutive entries of
 *                  the id list; consecutive ids of one lattice column (an octree frontier is compacted
 *                  in C order) share one set of per-column constants.  Every list takes this form
 *                  whatever its run lengths, so a point's value does not depend on how a caller cuts
 *                  a list into calls.
 * Anything else takes the per-layer kernels.  pifu_set_chain(ctx, 0) forces the per-layer kernels
 * everywhere, 2 keeps only the lattice form (A/B measurements, parity tests between the paths), 1
 * (default) enables both; pifu_chain_ready reports 1 when the chain operands are built and enabled. */
int pifu_set_chain(pifu_ctx* ctx, int enabled);
int pifu_chain_ready(pifu_ctx* ctx);

/* Test hook: one layer, Y = act(W X^T + b), through the layer kernel (synchronous).  All
 * pointers are device fp32: X [M][K] (points x channels), W [N][K], b [N]; Y is channel-major
 * [N][M] like the reference's [C, N] activations. */
int pifu_debug_gemm(pifu_ctx* ctx, const float* X, const float* W, const float* b, int M, int K, int N,
                    int leaky, float* Y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIFU_B200_H */
