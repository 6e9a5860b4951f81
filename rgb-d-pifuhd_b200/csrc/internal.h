// Declarations shared between the translation units of libpifu_b200.so (not part of the ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct pifu_ctx;

namespace pifu {

int launch_pack_weights(const float* W, int cin, const int* colmap, int num_kb, int N, int BN,
                        uint8_t* out, cudaStream_t s);
int launch_pack_weights_split(const float* W, int cin, const int* colmap, int num_kb, int w_nkb, int kb0, int N, int BN,
                              int lo, uint8_t* out, cudaStream_t s);
int launch_pack_rows(const float* X, int M, int K, int num_kb, uint8_t* out, cudaStream_t s);
int launch_nchw_to_nhwc(const float* in, float* out, int C, int HW, cudaStream_t s);
int launch_unblock(const uint8_t* buf, const uint8_t* buf_lo, int kb_stride, int kb_off, int C, int n, float* dst, long long ld,
                   cudaStream_t s);

// gather.cu: bilinear samples of an NCHW image at projected points (vertex colours); calib12 = host, rows 0..2 of the 4x4
int launch_sample_image(const float* img, int C, int H, int W, const float* pts, long long pstride, long long n,
                        const float* calib12, int perspective, float* out, cudaStream_t s);

// meshclean.cu
int mesh_clean(const double* verts, const double* colors, const int* faces, long long nv, long long nf, int only_watertight,
               int axis, double* out_verts, double* out_colors, int* out_faces, long long* counts_host, int num_sms, cudaStream_t s);

// norm.cu: normalised stacks (GroupNorm / batch statistics) and the un-fused last layer
struct ASeg;
int launch_group_norm(const float* x, uint8_t* buf, uint8_t* buf_lo, int nkb, int channels, int groups, int m_tiles, int n_valid,
                      const float* gamma, const float* beta, double eps, double* stats, cudaStream_t s);
int launch_head(const ASeg* segs, int nseg, const float* w, float b, const uint8_t* mask, int mask_bit, float* out,
                int m_tiles, int n_valid, cudaStream_t s);

// refine.cu: hybrid precision, compaction of the near-surface points of a call and scatter of their re-evaluation
int launch_select_band(const float* out, long long n, float lo, float hi, const long long* ids, long long key0,
                       long long pos0, long long* sel_key, long long* sel_pos, unsigned long long* count, int num_sms,
                       cudaStream_t s);
int launch_scatter(const float* vals, const long long* pos, int m, float* out, cudaStream_t s);

// runs.cu: segmentation of a sorted lattice-id list into runs of rows that share a lattice column
constexpr int RUN_BLOCK_ROWS = 1024;
int launch_run_heads(const long long* ids, long long n, int R2, uint32_t* block_heads, cudaStream_t s);
int launch_run_assign(const long long* ids, long long row0, int m, int R2, const uint32_t* block_heads,
                      int* rowseg, long long* seg_ids, cudaStream_t s);

// api.cu services used by the octree driver
int eval_ids(pifu_ctx* c, int levels, int R0, int R1, int R2, const long long* ids, long long n,
             const float* calib, const double* calib_inv, float* out, cudaStream_t s);
int ctx_check_ready(pifu_ctx* c, int levels);      // MLPs + feature maps set, workspace allocated
bool ctx_normalised(pifu_ctx* c, int levels);      // a normalised MLP: every entry call is one statistics domain
int ctx_num_sms(pifu_ctx* c);
void ctx_count_launch(pifu_ctx* c, int n);

int octree_begin(pifu_ctx* c, int R0, int R1, int R2, int init_res, double threshold, cudaStream_t s, float* sdf32_target,
                 int lb, int le, int fb, int fe);      // le < 0: the whole volume
int octree_frontier(pifu_ctx* c, long long* n, cudaStream_t s);
const long long* octree_ids(pifu_ctx* c);
int octree_commit(pifu_ctx* c, const float* vals, const double* vals64, cudaStream_t s, const long long* pair_ids, long long n_pairs);
int octree_export(pifu_ctx* c, double* sdf64, float* sdf32, cudaStream_t s);
int octree_vals(pifu_ctx* c, long long n, float** out);

struct OctreeState;
struct McState;
OctreeState*& ctx_octree(pifu_ctx* c);
McState*& ctx_mc(pifu_ctx* c);
void octree_free(OctreeState* s);
void mc_free(McState* s);

}  // namespace pifu
