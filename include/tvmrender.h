/*
 * tvmrender.h -- C ABI of libtvmrender.so: the TensoRF-VM per-ray rendering hot path of
 * tensorf-myc (FREDZEL2020/jittor-MYC-NeRFs), rebuilt as hand-written sm_100a CUDA kernels.
 *
 * Reference interface this library replaces (paths relative to the reference's tensorf-myc/):
 *   renderer.py:12-27                 OctreeRender_trilinear_fast   -> tvm_forward over all rays
 *   models/tensorBase.py:476-536      TensorBase.execute            -> tvm_forward / tvm_backward
 *   models/tensorBase.py:340-360      sample_ray                    -> march stage of tvm_forward
 *   models/tensorBase.py:39-59        AlphaGridMask.sample_alpha    -> tvm_pack_alpha + march stage
 *   models/tensoRF.py:209-244         compute_density/appfeature    -> gather stages
 *   models/tensorBase.py:62-86        MLPRender_Fea                 -> appearance stage
 *   models/tensorBase.py:17-24        raw2alpha                     -> composite stage
 *   models/tensorBase.py:451-473      compute_alpha                 -> tvm_density_alpha
 *   train.py:228,260                  loss.backward (Jittor autograd over the ops above) -> tvm_backward
 * The reference has no C ABI of its own; its only native-plugin idiom is Jittor's
 * jt.code(cuda_header=..., cuda_src=...) with raw inK_p/outK_p device pointers
 * (jnerf-myc/python/jnerf/models/samplers/density_grid_sampler/calc_rgb.py:35-75), which is what
 * binds to these symbols.
 *
 * Conventions: every pointer is a DEVICE pointer unless the name ends in _host; all floating point
 * is IEEE fp32; the caller owns every buffer; every entry point takes the CUDA stream (as void*) it
 * enqueues on and never synchronises; return value 0 = ok, negative = error (tvm_last_error() explains).
 * State kept by the library: a thread-local error string; a per-device cache of the SM count; and, only while
 * tvm_profile_enable(1) is in force, a mutex-guarded list of cudaEvents (measurement hook, see the end of
 * this file).  Nothing else survives a call: compute entry points are re-entrant and may run concurrently on
 * distinct streams with distinct workspaces.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TVMRENDER_H_
#define TVMRENDER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVM_ABI_VERSION 29

/* flags for tvm_forward / tvm_backward */
#define TVM_WHITE_BG      0x1u  /* rgb_map += 1 - acc_map          (tensorBase.py:523-524) */
#define TVM_NO_ERT        0x2u  /* disable early ray termination (march every in-box sample) */
#define TVM_MLP_FP32      0x0u  /* appearance head in fp32 FMA (parity 1e-4)                */
#define TVM_MLP_BF16      0x10u /* appearance head on tcgen05 tensor cores, bf16 x bf16 -> fp32 (parity 1e-2) */
#define TVM_MLP_FP16      0x20u /* tcgen05, fp16 x fp16 -> fp32: 11-bit operands; measured 2e-5 from the oracle, i.e. inside the
                                  fp32 tolerance 1e-4.  Its backward is the bf16 tensor-core one when TvmModel.tc_weights_bwd
                                  is given, else the fp32 one.  fp16 saturates at 65504: a model whose activations or
                                  features exceed that needs TVM_MLP_BF16 (same kernels, 8-bit mantissa, fp32 range)          */
#define TVM_MLP_MASK      0x30u
#define TVM_EVAL_ONLY     0x4u  /* tvm_forward only: the caller needs the two maps and nothing else -- no tvm_backward will follow on
                                  this workspace and nobody reads its stash.  The appearance head then composites as it goes:
                                  every entry adds w * rgb to its ray in 32-bit FIXED POINT (units of 2^-31: colours are sigmoids
                                  and sum(w) <= 1; a term is rounded to 4.7e-10, a ray to < 2.5e-7; integer addition is
                                  associative, so the pixels do not depend on the order of the atomics, on chunking or on the
                                  launch), a per-ray pass converts, adds the background and clamps.  k_composite, the per-block
                                  tables (blk_mask / blk_base, their memset) and the per-entry colours are not written: those
                                  members of TvmWorkspaceLayout are unspecified afterwards.  Ignored (the stash is written as
                                  usual) with TvmAux outputs, TVM_VARIANT_REF, TVM_SAMPLING_NPP and n_samples <= 64.           */

/* model variant */
#define TVM_VARIANT_VM    0     /* TensorVMSplit + MLPRender_Fea          (tensoRF.py:141, tensorBase.py:62) */
#define TVM_VARIANT_REF   1     /* REFTensoRF + MLPRender_Fea_Ref         (REFTensoRF.py:64, :5)             */
#define TVM_REF_HEAD_LD   48    /* basis_t row length for TVM_VARIANT_REF: [basis app_dim | normal 3 | diffuse 3 | specular 1 | rho 1 | 0...] */

/* sample placement along the ray */
#define TVM_SAMPLING_UNIFORM 0  /* TensorBase.sample_ray            (tensorBase.py:340-360)   */
#define TVM_SAMPLING_NPP     1  /* NerfPlusPlus.sample_ray          (nerfplusplus.py:239-269): depths linear from near to
                                   the exit of the radii-sphere, ALWAYS stratified-jittered; `jitter` is then [n][S]       */
#define TVM_NPP_BG_SAMPLES 512  /* nerfplusplus.py:285 */

/* activation (tensorBase.py:444-448) */
#define TVM_ACT_SOFTPLUS  0
#define TVM_ACT_RELU      1

/* Host-side descriptor of one model; device pointers refer to the PACKED layouts written by
 * tvm_pack_grid / tvm_pack_alpha / tvm_pack_linear (channels-last grids, bit-packed alpha volume,
 * [in][out] linear weights).  Plain-old-data, passed by pointer, copied by value into the launch. */
typedef struct TvmModel {
  /* scalars derived exactly as TensorBase.update_stepSize does (tensorBase.py:197-209) */
  float aabb[6];            /* xmin ymin zmin xmax ymax zmax                     */
  float inv_aabb_size[3];   /* 2.0 / aabbSize                (tensorBase.py:201) */
  int32_t grid[3];          /* gridSize Gx Gy Gz                                  */
  float step_size;          /* mean(units) * step_ratio      (tensorBase.py:204) */
  float near_, far_;        /* near_far                                           */
  float density_shift;      /* tensorBase.py:446                                  */
  float distance_scale;     /* tensorBase.py:511                                  */
  float weight_thres;       /* rayMarch_weight_thres (always 1e-4 in the reference) */
  int32_t act;              /* TVM_ACT_*                                          */
  int32_t n_density;        /* channels per density component (16)                */
  int32_t n_app;            /* channels per appearance component (48)             */
  int32_t app_dim;          /* basis_mat outputs (27)                             */
  int32_t view_pe, fea_pe;  /* positional-encoding frequencies (2, 2)             */
  int32_t feature_c;        /* MLP width (128)                                    */
  /* factor grids, channels-last: plane k is [G[m1]][G[m0]][C], line k is [G[v]][C]
   * with matMode [[0,1],[0,2],[1,2]], vecMode [2,1,0] (tensorBase.py:168-169)    */
  const float* density_plane[3];
  const float* density_line[3];
  const float* app_plane[3];
  const float* app_line[3];
  /* appearance head, [in][out_padded] layouts from tvm_pack_linear              */
  int32_t variant;          /* TVM_VARIANT_*                                              */
  const float* basis_t;     /* [3*n_app][32] basis_mat.weight^T, zero padded; TVM_VARIANT_REF: [3*n_app][48] =
                               (basis_mat | normal_linear | diffuse_linear | specular_linear | rho_linear)^T  */
  const float* head_bias;   /* TVM_VARIANT_REF: [48] biases of the stacked heads (0 for basis_mat); else NULL */
  const float* w1_t;        /* [in_mlp_c][feature_c] renderModule.mlp.0.weight^T          */
  const float* b1;          /* [feature_c]                                                */
  const float* w2_t;        /* [feature_c][feature_c]                                     */
  const float* b2;          /* [feature_c]                                                */
  const float* w3;          /* [3][feature_c]        renderModule.mlp.4.weight (as is)    */
  const float* b3;          /* [3]                                                        */
  /* alpha mask (NULL = no mask): bit (z*H + y)*W + x of a little-endian uint32 stream    */
  const uint32_t* alpha_bits;
  int32_t alpha_grid[3];    /* W(x) H(y) D(z)                                             */
  float alpha_aabb_min[3];  /* AlphaGridMask.aabb[0]          (tensorBase.py:44)          */
  float alpha_inv_size[3];  /* 1.0 / aabbSize * 2             (tensorBase.py:46)          */
  /* optional empty-space index from tvm_pack_alpha_bricks (NULL = none): one bit per 8x8x8-voxel
   * brick, bit (bz*BH + by)*BW + bx with B* = ceil(dim/8); set iff any voxel of the brick is set  */
  const uint32_t* alpha_bricks;
  /* optional 2x2x2-dilated copy of alpha_bits from tvm_pack_alpha_dilated (NULL = none): bit (z,y,x) = OR of the eight
   * corner bits (x..x+1, y..y+1, z..z+1); decides interior samples with one lookup, bit-identically to the 8-tap test    */
  const uint32_t* alpha_dilated;
  /* optional tensor-core operand images written by tvm_pack_mlp_tc (NULL = not packed)   */
  const void* tc_weights;
  int32_t sampling;         /* TVM_SAMPLING_*                                                            */
  float radii;              /* TVM_SAMPLING_NPP: radius of the bounding sphere (configs/Scarf.txt:14)    */
  /* optional 16-bit PAIR RECORDS of the appearance grids written by tvm_pack_pair16 in the format of the mode they are used
   * with (bf16 for TVM_MLP_BF16, fp16 for TVM_MLP_FP16; all NULL = none).  Record x of row r holds texels (r, x) and
   * (r, min(x+1, W-1)) interleaved by groups of 4 channels: [t0 c..c+3 | t1 c..c+3] = 16 bytes per group, 4*C bytes per
   * record; a line is one row.  When present the tensor-core appearance head gathers from them: one 16-byte load per lane
   * brings both taps of an axis pair (a 64-byte segment per sample and instruction; half the gather bytes and a third of
   * the load instructions of the fp32 grids; the plane x line products are rounded to that format as the GEMM operand
   * anyway).  TVM_MLP_FP32 and every backward kernel ignore them.                                                        */
  const void* app_plane_pair[3];
  const void* app_line_pair[3];
  /* optional second operand image for tvm_backward, packed by tvm_pack_mlp_tc with TVM_MLP_BF16 (NULL = none): lets a
   * TVM_MLP_FP16 step (fp16 forward, inside the fp32 tolerance) take the tensor-core backward, whose operands are bf16 */
  const void* tc_weights_bwd;
  /* optional neighbourhood words of the brick index from tvm_pack_alpha_bricks3 (NULL = none): one uint32 per 8^3 brick,
   * bit (dz+1)*9 + (dy+1)*3 + (dx+1) = brick (x+dx, y+dy, z+dz) holds a set voxel.  The empty-space test of a 32-sample
   * block is then one load and one AND instead of a loop over up to 27 bricks.  Never changes a mask decision.              */
  const uint32_t* alpha_bricks3;
} TvmModel;

/* NerfPlusPlus background network (nerfplusplus.py:66-140 with bg_D=3, W=128, skips=[1], bg_freq=2,
 * bg_view_freq=2: position embedding 20, view embedding 15), fp32, [in][out] layouts.  The 128->256
 * base_remap layer has no activation, so it is folded into the first rgb layer by tvm_bg_fold
 * (a weights-only product): rgb_layers.0(cat(remap(x), v)) = (W_a W_r) x + W_v v + (W_a b_r + b).     */
typedef struct TvmBgNet {
  const float* w0_t;        /* [20][128]   base_layers.0.0.weight^T                                    */
  const float* b0;          /* [128]                                                                    */
  const float* w1_t;        /* [128][128]  base_layers.1.0                                              */
  const float* b1;
  const float* w2_t;        /* [148][128]  base_layers.2.0: rows 0..19 input_pts, 20..147 base (skip)   */
  const float* b2;
  const float* w_sigma;     /* [128]       sigma_layers.0.weight                                        */
  const float* b_sigma;     /* [1]                                                                      */
  const float* wf_t;        /* [128][64]   (rgb_layers.0.weight[:, :256] @ base_remap_layers.0.weight)^T */
  const float* bf;          /* [64]        rgb_layers.0.weight[:, :256] @ base_remap.bias + rgb_layers.0.bias */
  const float* wv_t;        /* [15][64]    rgb_layers.0.weight[:, 256:271]^T                            */
  const float* w_rgb;       /* [3][64]     rgb_layers.2.weight                                          */
  const float* b_rgb;       /* [3]                                                                      */
  const void* tc_weights;   /* optional bf16 tensor-core operand image written by tvm_pack_bg_tc (NULL = not packed;
                               required when tvm_forward_npp runs with TVM_MLP_BF16)                    */
} TvmBgNet;

/* Optional per-sample outputs for parity tests (any member may be NULL).  Requesting aux
 * disables early ray termination so that every mask bit is produced.                         */
typedef struct TvmAux {
  uint32_t* bbox_bits;      /* [n][ceil(S/32)]  in-bbox mask of sample_ray (tensorBase.py:358)      */
  uint32_t* valid_bits;     /* [n][ceil(S/32)]  ray_valid after the alpha mask (tensorBase.py:491-496) */
  uint32_t* app_bits;       /* [n][ceil(S/32)]  app_mask = weight > thres (tensorBase.py:513)       */
  float* sigma;             /* [n][S]                                                                */
  float* weight;            /* [n][S]                                                                */
  float* rgb;               /* [n][S][3]        per-sample colour (0 where !app_mask)                */
  float* acc_map;           /* [n]                                                                   */
  float* bg_lambda;         /* [n]    tvm_forward_npp: gated foreground transmittance (nerfplusplus.py:277-278,313)  */
  float* bg_rgb_map;        /* [n][3] tvm_forward_npp: background colour before the bg_lambda factor                 */
  float* penalty;           /* [1] TVM_VARIANT_REF: += sum w[app] * relu(-d.n)^2 (REFTensoRF.py:236-238); does not
                               by itself select the parity (no-ERT) instantiation                                   */
} TvmAux;

/* Gradient targets of tvm_backward: same PACKED layouts as TvmModel; accumulated with
 * atomics (red.global.add.f32); the caller zeroes them.                                       */
typedef struct TvmGrads {
  float* density_plane[3];
  float* density_line[3];
  float* app_plane[3];
  float* app_line[3];
  float* basis_t;           /* [3*n_app][32]; TVM_VARIANT_REF: [3*n_app][48] (basis_mat | normal | diffuse | specular | rho) */
  float* head_bias;         /* TVM_VARIANT_REF: [48] biases of the stacked heads; else NULL                                  */
  float* w1_t;              /* [in_mlp_c][feature_c] */
  float* b1;
  float* w2_t;
  float* b2;
  float* w3;                /* [3][feature_c] */
  float* b3;
} TvmGrads;

/* Work counters filled by tvm_forward (device memory, 8 x uint64): see TVM_CNT_*              */
#define TVM_CNT_M_IN   0    /* samples inside the bbox that were marched                        */
#define TVM_CNT_M_V    1    /* samples whose density was gathered (passed the alpha mask)       */
#define TVM_CNT_M_A    2    /* samples with weight > thres (appearance evaluated)               */
#define TVM_CNT_RAYS   3    /* rays with at least one gathered sample                           */
#define TVM_CNT_BG_RAYS 4   /* tvm_forward_npp: rays whose background was evaluated (bg_lambda > 0.1) */
#define TVM_CNT_BG_SAMPLES 5 /* tvm_forward_npp: background samples pushed through the bg MLP       */
#define TVM_CNT_WORDS  8

const char* tvm_last_error(void);
int tvm_abi_version(void);
/* number of CUDA devices visible, or negative on error (no CPU fallback exists) */
int tvm_device_count(void);

/* ---- parameter upload (replaces nothing in the reference: Jittor owns NCHW Vars) ------------ */
/* NCHW [1,C,H,W] -> channels-last [H][W][C]; lines are the W==1 case.                          */
int tvm_pack_grid(const float* nchw, int C, int H, int W, float* out_hwc, void* stream);
/* inverse of tvm_pack_grid (used on gradients so the host optimiser sees NCHW)                */
int tvm_unpack_grid(const float* hwc, int C, int H, int W, float* out_nchw, void* stream);
/* Several 2-D transposes dst[c][r] = src[r][c] in one launch: tvm_pack_grid is the transpose of [C][H*W], tvm_unpack_grid of
 * [H*W][C], tvm_pack_linear / tvm_unpack_linear a transpose with a padded leading dimension, a bias a 1-row job: a training
 * step moves all parameters (and all gradients) with one call each.  jobs_host is a HOST array.                          */
#define TVM_TRANSPOSE_MAX 32
typedef struct TvmTransposeJob {
  const float* src;         /* [rows][src_ld], columns 0..cols-1 are read */
  float* dst;               /* [cols][dst_ld], columns 0..rows-1 are written (tvm_pack_linear's padding stays as it is) */
  int32_t rows, cols;
  int32_t src_ld, dst_ld;   /* leading dimensions in floats; 0 = dense (cols / rows).  rows = 1 is a plain copy. */
} TvmTransposeJob;
int tvm_transpose_batch(const TvmTransposeJob* jobs_host, int n_jobs, void* stream);
/* Linear weight [out][in] -> [in][out_pad] (zero padded columns)                              */
int tvm_pack_linear(const float* w_out_in, int out_c, int in_c, int out_pad, float* out_t, void* stream);
int tvm_unpack_linear(const float* w_t, int out_c, int in_c, int out_pad, float* out_w, void* stream);
/* channels-last fp32 grid [rows][W][C] (a plane; a line is rows = 1, W = L) -> [rows][W] pair records of 2*C 16-bit values
 * (bf16 for flags = TVM_MLP_BF16, fp16 for TVM_MLP_FP16, round to nearest even): TvmModel.app_plane_pair / app_line_pair.
 * C must be a multiple of 4; dst holds rows*W*2*C values and must be 16-byte aligned.                                  */
int tvm_pack_pair16(const float* src, int rows, int W, int C, void* dst, uint32_t flags, void* stream);
/* {0,1} fp32 volume [D][H][W] -> bit stream (bit set iff value > 0); n_words = ceil(D*H*W/32)  */
int tvm_pack_alpha(const float* volume, int D, int H, int W, uint32_t* bits, void* stream);
/* brick index of a packed alpha volume; n_words = ceil(ceil(D/8)*ceil(H/8)*ceil(W/8) / 32)        */
int tvm_pack_alpha_bricks(const uint32_t* bits, int D, int H, int W, uint32_t* bricks, void* stream);
/* Neighbourhood words of a brick index: bricks3 = uint32 [ceil(D/8) * ceil(H/8) * ceil(W/8)], see TvmModel.alpha_bricks3 */
int tvm_pack_alpha_bricks3(const uint32_t* bricks, int D, int H, int W, uint32_t* bricks3, void* stream);
/* 2x2x2-dilated copy of a packed alpha volume (same size and bit order as `bits`)                    */
int tvm_pack_alpha_dilated(const uint32_t* bits, int D, int H, int W, uint32_t* dilated, void* stream);
/* bytes of the tensor-core operand image for (in_mlp_c, feature_c, app_dim, n_app)            */
size_t tvm_tc_weights_bytes(const TvmModel* m_host);
/* builds the K-major UMMA operand images of basis/W1/W2/W3 (bf16 or fp16 per `flags`) from the packed fp32 weights; the
 * backward kernel (TVM_MLP_BF16 only) reads the same buffer */
int tvm_pack_mlp_tc(const TvmModel* m_host, void* tc_weights_out, uint32_t flags /* TVM_MLP_BF16 | TVM_MLP_FP16 */, void* stream);

/* NeRF++ background network on the tensor cores: bytes of / builder for TvmBgNet.tc_weights (16-byte aligned) */
size_t tvm_bg_tc_bytes(void);
int tvm_pack_bg_tc(const TvmBgNet* bg_host, void* tc_weights_out, void* stream);

/* ---- the hot path --------------------------------------------------------------------------- */
/* bytes of scratch for n rays x n_samples in the worst case (every sample weighted): never overflows; what tvm_backward* need */
int tvm_workspace_bytes(int n_rays, int n_samples, size_t* out_bytes);

/* Bounded workspaces (evaluation renders).  The worst case above is 44 B x n_rays x n_samples -- 29 GB for an 800 x 800 frame at
 * 1036 samples, of which a trained scene uses 1-2 %.  tvm_forward / tvm_forward_npp therefore accept ANY workspace that holds the
 * per-ray tables and at least one entry per ray: the entry list is then bounded by what fits
 * (tvm_workspace_capacity(n_rays, n_samples, ws_bytes); tvm_workspace_bytes_bounded is the inverse).  Samples beyond the
 * capacity are COUNTED but not stored, nothing is written outside the workspace, and the rays' outputs are then incomplete:
 * the caller reads tvm_forward_entries (one 4-byte copy, synchronises `stream`) after the launch and, when it returns more than
 * the capacity, renders the rays again in smaller pieces or a larger workspace -- the count tells how large.  The reference has
 * no counterpart (renderer.py:19-27 slices rays into fixed chunks and lets Jittor allocate per op); this is the
 * overflow-and-relaunch contract of the host mirror (tensorf.py::_render_bounded).  tvm_backward* need the worst-case size. */
int tvm_workspace_bytes_bounded(int n_rays, int n_samples, uint32_t max_entries, size_t* out_bytes);
int tvm_workspace_capacity(int n_rays, int n_samples, size_t ws_bytes, uint32_t* out_entries);
int tvm_forward_entries(const void* ws, void* stream, uint32_t* out_entries);

/* Where tvm_forward leaves the per-sample results inside the caller's workspace (byte offsets from `ws`), valid until the next
 * call on that workspace.  This is how the PRODUCTION march (no TvmAux, empty-space skipping and early ray termination on)
 * is held to the reference's masks: blk_mask [n][n_blocks] are the app_mask bits (weight > thres, tensorBase.py:513) of
 * each 32-sample block, ent [*n_entries] the (ray, sample) pairs in ray-major order -- the compaction order of the
 * reference's boolean indexing (tensorBase.py:516-518) within a ray -- ent_w their weights, ent_rgb their colours, acc the
 * acc_map (tensorBase.py:520).  A Jittor binding can expose them as extra outputs.  After a TVM_EVAL_ONLY launch n_entries,
 * ent, ent_w and acc are as described (same march), blk_mask / blk_base / ent_rgb / rgb_sum are unspecified.              */
typedef struct TvmWorkspaceLayout {
  size_t n_entries;         /* uint32: number of entries (samples with weight > thres)                  */
  size_t blk_mask;          /* uint32 [n][n_blocks]: app_mask bits of block b of ray r (0 = never visited) */
  size_t blk_base;          /* uint32 [n][n_blocks]: index of the block's first entry (valid where blk_mask != 0) */
  size_t ent;               /* uint32 [capacity][2]: (ray, sample index)                                 */
  size_t ent_w;             /* float  [capacity]: weight                                                 */
  size_t ent_rgb;           /* float  [capacity][3]: colour of the sample (tensorBase.py:518)            */
  size_t acc;               /* float  [n]: acc_map                                                       */
  size_t rgb_sum;           /* float  [n][3]: sum_k w_k rgb_k before white background / clamp            */
  uint32_t capacity;        /* entries the workspace can hold (n * n_samples)                            */
  int32_t n_blocks;         /* ceil(n_samples / 32)                                                      */
  size_t bytes;             /* = tvm_workspace_bytes                                                     */
} TvmWorkspaceLayout;
int tvm_workspace_layout(int n_rays, int n_samples, TvmWorkspaceLayout* out);

/* TensorBase.execute for n rays (tensorBase.py:476-536, ndc_ray=False):
 *   rays [n][6] (origin, unit direction), jitter [n] or NULL (is_train: rng += U[0,1) per ray,
 *   tensorBase.py:351-353), rgb_map [n][3], depth_map [n], counters [TVM_CNT_WORDS] uint64 or NULL
 *   (accumulated, caller zeroes).  The workspace keeps what tvm_backward needs until the next call (not with TVM_EVAL_ONLY). */
int tvm_forward(const TvmModel* m_host, const float* rays, int n_rays, int n_samples,
                const float* jitter, uint32_t flags, float* rgb_map, float* depth_map,
                const TvmAux* aux_host, uint64_t* counters, void* ws, size_t ws_bytes, void* stream);

/* NerfPlusPlus.execute (nerfplusplus.py:272-318): foreground = tvm_forward on black with TVM_SAMPLING_NPP
 * (fg_rand [n][S]); bg_lambda = prod(1 - alpha + 1e-6), zeroed when <= 0.1; background = 512 inverse-depth
 * samples per ray (bg_rand [n][512]) through the bg MLP, composited front to back;
 * rgb_map += bg_lambda * bg_rgb_map.  Rays with bg_lambda == 0 skip the background entirely.
 * TVM_MLP_FP32: background MLP in fp32 FMA; TVM_MLP_BF16: on tcgen05 tensor cores (bg_host->tc_weights).  */
int tvm_forward_npp(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays, int n_samples,
                    const float* fg_rand, const float* bg_rand, uint32_t flags, float* rgb_map, float* depth_map,
                    const TvmAux* aux_host, uint64_t* counters, void* ws, size_t ws_bytes, void* stream);
/* weights-only fold of base_remap into rgb_layers.0 (see TvmBgNet)                                       */
int tvm_bg_fold(const float* remap_w /*[256][128]*/, const float* remap_b /*[256]*/, const float* rgb0_w /*[64][271]*/,
                const float* rgb0_b /*[64]*/, float* wf_t /*[128][64]*/, float* bf /*[64]*/, float* wv_t /*[15][64]*/,
                void* stream);

/* Backward of tvm_forward w.r.t. every parameter, given d_rgb_map [n][3] = dL/d rgb_map
 * (row a12 of SURVEY §8a; coordinates are detached, depth carries no gradient).  Must follow a
 * tvm_forward with the same arguments on the same workspace.  TVM_VARIANT_REF: d_penalty is a DEVICE scalar
 * dL/d penalty (train.py:253-255: normal_vector_penalty_weight) or NULL.                                  */
int tvm_backward(const TvmModel* m_host, const float* rays, int n_rays, int n_samples,
                 const float* jitter, uint32_t flags, const float* rgb_map, const float* d_rgb_map,
                 const float* d_penalty, const TvmGrads* grads_host, void* ws, size_t ws_bytes, void* stream);

/* Gradient targets of the NeRF++ background network: same packed layouts as TvmBgNet (tc_weights has no gradient);
 * accumulated with atomics, the caller zeroes them.  wf_t / bf / wv_t are gradients of the FOLDED first colour layer;
 * tvm_bg_fold_bwd carries them back to base_remap_layers.0 and rgb_layers.0.                               */
typedef struct TvmBgGrads {
  float* w0_t; float* b0; float* w1_t; float* b1; float* w2_t; float* b2;
  float* w_sigma; float* b_sigma; float* wf_t; float* bf; float* wv_t; float* w_rgb; float* b_rgb;
} TvmBgGrads;

/* Backward of tvm_forward_npp (NerfPlusPlus.execute, nerfplusplus.py:272-318, under Jittor autograd; train.py:228,260):
 * rgb_map = fg + bg_lambda * bg_rgb_map, so besides tvm_backward's foreground terms (black background) the density
 * grids receive dL/d bg_lambda through prod(1 - alpha + 1e-6) (zero where the gate bg_lambda <= 0.1 closed) and the
 * background network receives bg_lambda * d_rgb_map through the 512-sample compositing.  Must follow a tvm_forward_npp
 * with the same arguments on the same workspace.  The background backward runs in fp32 in every TVM_MLP_* mode.     */
int tvm_backward_npp(const TvmModel* m_host, const TvmBgNet* bg_host, const float* rays, int n_rays, int n_samples,
                     const float* fg_rand, const float* bg_rand, uint32_t flags, const float* rgb_map,
                     const float* d_rgb_map, const TvmGrads* grads_host, const TvmBgGrads* bg_grads_host, void* ws,
                     size_t ws_bytes, void* stream);
/* transpose of tvm_bg_fold: gradients of the folded layer -> base_remap_layers.0.{weight,bias}, rgb_layers.0.{weight,bias}
 * (outputs are overwritten, not accumulated)                                                                         */
int tvm_bg_fold_bwd(const float* remap_w /*[256][128]*/, const float* remap_b /*[256]*/, const float* rgb0_w /*[64][271]*/,
                    const float* d_wf_t /*[128][64]*/, const float* d_bf /*[64]*/, const float* d_wv_t /*[15][64]*/,
                    float* d_remap_w /*[256][128]*/, float* d_remap_b /*[256]*/, float* d_rgb0_w /*[64][271]*/,
                    float* d_rgb0_b /*[64]*/, void* stream);

/* TensorBase.compute_alpha (tensorBase.py:451-473): alpha = 1 - exp(-sigma(xyz) * length)       */
int tvm_density_alpha(const TvmModel* m_host, const float* xyz, int n_pts, float length,
                      float* alpha_out, void* stream);

/* MSE loss head of train.py:228: loss = mean((rgb_map - target)^2), d_rgb_map = 2 (rgb_map - target) / (3 n) * scale */
int tvm_mse_loss(const float* rgb_map, const float* target, int n_rays, float grad_scale,
                 float* loss_out, float* d_rgb_map, void* stream);

/* ---- callers either side of the ray path ------------------------------------------------- */
/* TensorBase.getDenseAlpha (tensorBase.py:366-384) on a lattice of grid_host = {Gx, Gy, Gz} nodes spanning the model aabb:
 * alpha = compute_alpha(node, length), clamped to [0,1] and written transposed as updateAlphaMask wants it, [Gz][Gy][Gx].   */
int tvm_dense_alpha(const TvmModel* m_host, const int32_t* grid_host, float length, float* alpha_zyx, void* stream);
/* TensorBase.updateAlphaMask (tensorBase.py:386-409) after getDenseAlpha: max_pool3d(k=3, pad=1), >= thres -> {0,1};
 * bits_out = bit-packed volume (layout of tvm_pack_alpha), volume_out = the same as fp32 [Gz][Gy][Gx] or NULL,
 * bbox_idx[6] = {min x,y,z, max x,y,z} voxel indices of the set voxels ({INT_MAX.., -1..} when none), n_set = their count. */
int tvm_alpha_mask_from_dense(const float* alpha_zyx, const int32_t* grid_host, float thres, float* volume_out,
                              uint32_t* bits_out, int32_t* bbox_idx, uint64_t* n_set, void* stream);
/* TensorBase.filtering_rays (tensorBase.py:411-441): mask_out[i] = 1 iff ray i is kept.  bbox_only: t_max > t_min of the slab
 * test (:424-429); else any of the n_samples points of self.sample_ray(..., is_train=False) has sample_alpha > 0 (:432-433;
 * no bbox gate, as the reference).  TVM_SAMPLING_NPP: that sampler is NerfPlusPlus.sample_ray, always jittered --
 * fg_rand [n][n_samples] carries its draws (NULL otherwise).                                                              */
int tvm_filter_rays(const TvmModel* m_host, const float* rays, int n_rays, int n_samples, int bbox_only,
                    const float* fg_rand, uint8_t* mask_out, void* stream);
/* get_ray_directions / get_ray_directions_blender + get_rays (dataLoader/ray_utils.py:81-153), optional unit
 * normalisation of the camera-space direction (dataLoader/blender.py:75); c2w_host is a HOST [3][4] matrix; rays_out [H*W][6]. */
int tvm_generate_rays(const float* c2w_host, int H, int W, float fx, float fy, float cx, float cy, int blender,
                      int normalize, float* rays_out, void* stream);
/* up_sampling_VM (tensoRF.py:248-262): F.interpolate(bilinear, align_corners=True) of one NCHW grid [C][H][W] -> [C][H2][W2]   */
int tvm_upsample_grid(const float* src_nchw, int C, int H, int W, float* dst_nchw, int H2, int W2, void* stream);
/* The same for up to TVM_UPSAMPLE_MAX_GRIDS grids in ONE launch (the three planes and three lines of up_sampling_VM, or all
 * twelve of upsample_volume_grid, tensoRF.py:264-269): HOST arrays src_nchw[n], dst_nchw[n] of device pointers,
 * src_chw[n][3] = {C, H, W}, dst_hw[n][2] = {H2, W2}.  Same arithmetic per output as tvm_upsample_grid (bit-identical).      */
#define TVM_UPSAMPLE_MAX_GRIDS 12
int tvm_upsample_grids(int n_grids, const float* const* src_nchw, const int32_t* src_chw, float* const* dst_nchw,
                       const int32_t* dst_hw, void* stream);

/* Regularisers of train.py:233-251, value and gradient in one pass: *loss_accum += weight * f(x) (device scalar, may be NULL),
 * grad += weight * df/dx (same shape as x, may be NULL).  weight_dev (nullable): DEVICE scalar multiplied into `weight`
 * (the per-step weight of a replayed CUDA graph).  TVLoss (utils.py:123-142) of one plane [1][C][H][W];
 * density_L1's mean |x| (tensoRF.py:191-195); vectorDiffs (tensoRF.py:177-186) of one line [1][C][L][1].                        */
int tvm_tv_loss(const float* plane_nchw, int C, int H, int W, float weight, const float* weight_dev, float* loss_accum,
                float* grad_nchw, void* stream);
/* tvm_tv_loss for up to TVM_TV_MAX planes in one launch (the six factor planes of a step); jobs_host is a HOST array */
#define TVM_TV_MAX 8
typedef struct TvmTvJob {
  const float* plane_nchw;  /* [1][C][H][W] */
  float* grad_nchw;         /* += weight * d TV / d plane, or NULL */
  int32_t C, H, W;
  float weight;
  const float* weight_dev;  /* nullable DEVICE scalar multiplied into weight */
  int32_t overwrite;        /* 0: grad += ...; 1: grad = ... (a fresh gradient buffer needs neither a memset nor a read) */
} TvmTvJob;
int tvm_tv_loss_batch(const TvmTvJob* jobs_host, int n_jobs, float* loss_accum, void* stream);
int tvm_l1_loss(const float* x, size_t n, float weight, const float* weight_dev, float* loss_accum, float* grad, void* stream);
int tvm_vector_diffs(const float* line_cl, int C, int L, float weight, const float* weight_dev, float* loss_accum, float* grad,
                     void* stream);
/* jt.optim.Adam.step (train.py:187,261) over n_tensors parameter tensors in one launch per TVM_ADAM_MAX_TENSORS:
 * m = b0 m + (1-b0) g; v = b1 v + (1-b1) g^2; p -= m * (lr sqrt(1-b1^step)/(1-b0^step)) / (sqrt(v) + eps); step is 1-based.
 * hyper_dev (nullable): DEVICE array {sqrt(1-b1^step)/(1-b0^step), lr[0], lr[1], ...}; when given it overrides `step` and the
 * tensors' `lr` with hyper_dev[0] and hyper_dev[1 + lr_index], so that a captured CUDA graph can be replayed with new values. */
#define TVM_ADAM_MAX_TENSORS 32
typedef struct TvmAdamTensor {
  float* p;          /* parameter (updated in place)  */
  const float* g;    /* gradient                      */
  float* m;          /* first moment                  */
  float* v;          /* second moment                 */
  size_t n;          /* elements                      */
  float lr;          /* learning rate of its group    */
  int32_t lr_index;  /* index of its group in hyper_dev */
} TvmAdamTensor;
int tvm_adam_step(const TvmAdamTensor* tensors_host, int n_tensors, float beta0, float beta1, float eps, int step,
                  const float* hyper_dev, void* stream);

/* ---- data-parallel training: gradient exchange over NVLink peer memory --------------------------------- */
/* In-place SUM all-reduce of floats [offset, offset + n) of a buffer every rank of ONE NVSwitch node holds at the same
 * (symmetric) allocation, written directly over peer memory (no NCCL): two-shot in one kernel -- rank r reduces slice r
 * (NVLS: multimem.ld_reduce adds inside the switch and multimem.st broadcasts; without a multicast address: peer loads and
 * peer stores), bracketed by per-CTA flag barriers with the same CTA of every peer.  Plain kernel nodes: capturable into a
 * CUDA graph; the result is bit-identical on every rank.  Every rank must enqueue the same sequence of calls.
 * The reference (train.py:260) is single-process: this replaces the all-reduce a data-parallel Jittor run would issue.     */
#define TVM_AR_MAX_WORLD 16
#define TVM_AR_MAX_CTAS 128
typedef struct TvmPeerComm {
  void* bufs[TVM_AR_MAX_WORLD];         /* peer-mapped DEVICE pointers to the symmetric buffer, index = rank (own included)   */
  uint32_t* signals[TVM_AR_MAX_WORLD];  /* peer-mapped pointers to each rank's signal pad: tvm_allreduce_signal_words uint32,
                                           zeroed once before the first call (then never reset: flags carry an epoch)          */
  void* multicast;                      /* NVLS multicast address of the buffer, or NULL (peer-to-peer path)                 */
  uint32_t* epoch_dev;                  /* this rank's call counter: TWO device words, zeroed once                           */
  int32_t rank, world;
} TvmPeerComm;
int tvm_allreduce_signal_words(int world, size_t* out_words);
/* offset_floats and n_floats must be multiples of 4; n_ctas (1..TVM_AR_MAX_CTAS, the same on every rank) CTAs of 512 threads */
int tvm_allreduce_sum(const TvmPeerComm* comm_host, size_t offset_floats, size_t n_floats, int n_ctas, void* stream);

/* tvm_backward with the gradient exchange of a data-parallel step fused into its schedule: the flat gradient buffer
 * [0, total) the TvmGrads point into is the communicator's symmetric buffer, laid out density grids first ([0, split)), then
 * appearance grids / basis / MLP ([split, total)).  As soon as the appearance backward has finished, [split, total) is
 * all-reduced on `side_stream` while the density scatter (k_march_bwd) runs on `stream`; [0, split) follows on `stream`.
 * On return (stream order) the whole buffer holds the sum over the ranks.                                              */
typedef struct TvmGradExchange {
  const TvmPeerComm* comm;
  size_t split_floats, total_floats;     /* multiples of 4 */
  int32_t n_ctas;                         /* grid of the exposed all-reduce ([0, split))                                      */
  int32_t n_ctas_overlapped;              /* grid of the all-reduce that shares the SMs with k_march_bwd (keep it small)      */
  void* side_stream;                      /* a second stream of the caller, != stream                                         */
  int32_t phase;                          /* 0: the whole backward, both exchanges, joined into `stream` on return.
                                             1 / 2: the two halves as separate calls that leave BOTH exchanges on side_stream and do
                                             not join -- 1 = appearance backward + exchange of [split, total); 2 = density scatter +
                                             exchange of [0, split).  The caller records an event on side_stream after each call and
                                             waits for it before touching that half, so the appearance tail of the step (gradient
                                             unpack, regularisers, Adam) can run while the density half is still on the wire.      */
} TvmGradExchange;
int tvm_backward_dp(const TvmModel* m_host, const float* rays, int n_rays, int n_samples, const float* jitter,
                    uint32_t flags, const float* rgb_map, const float* d_rgb_map, const float* d_penalty,
                    const TvmGrads* grads_host, void* ws, size_t ws_bytes, const TvmGradExchange* xchg, void* stream);

/* Known-answer self-test of the tcgen05 shared-memory descriptor conventions (K-major and MN-major reads of one image);
 * P [128][128], Q [128][160], W [128][128] fp32 -> D1 = P^T Q [128][160], D2 = P W [128][128], D3 = P W^T [128][128].
 * Test hook: no reference counterpart.                                                                              */
int tvm_selftest_umma(const float* P, const float* Q, const float* W, float* D1, float* D2, float* D3, void* stream);

/* Measurement: pure 64-byte gathers (4 lanes x 16 B per tap, 18 taps in flight, the access shape of the density gather) at
 * pseudo-random 64-byte texels of buf[n_floats]; n_groups lane groups x iters x 18 taps x 64 B are read.  With a buffer
 * that fits L2 this is the "L2 gather peak" bench.py quotes beside the HBM roofline.  No reference counterpart. */
int tvm_bench_gather(const float* buf, size_t n_floats, int n_groups, int iters, float* sink, void* stream);

/* ---- measurement hooks (bench.py's roofline leg; off by default) ------------------------------ */
/* When enabled, every kernel tvm_forward / tvm_backward launches is bracketed by cudaEvents on the
 * caller's stream.  Stages: see TVM_STAGE_*.  Process-global (one mutex-guarded record list shared by all streams and
 * threads; events are created on the device current at the launch): a measurement aid, off by default. */
#define TVM_STAGE_MARCH      0
#define TVM_STAGE_APP        1
#define TVM_STAGE_COMPOSITE  2
#define TVM_STAGE_BWD_APP    3
#define TVM_STAGE_BWD_MARCH  4
#define TVM_STAGE_BG         5
#define TVM_STAGE_BWD_BG     6
#define TVM_STAGE_COUNT      8
int tvm_profile_enable(int on);
/* waits for the recorded events, adds per-stage milliseconds / launch counts into the two
 * [TVM_STAGE_COUNT] host arrays and clears the records */
int tvm_profile_collect(float* ms_by_stage_host, int* launches_by_stage_host);

#ifdef __cplusplus
}
#endif
#endif /* TVMRENDER_H_ */
