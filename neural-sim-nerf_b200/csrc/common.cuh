// Shared declarations of libnsr_b200: packed-network layout, launch helpers, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nsr_b200.h"

namespace nsr {

// ----------------------------------------------------------------------------- packed network layout
// One network = NUM_CHUNKS operand chunks (fp16, [128 out-rows x 64 K] each, UMMA K-major
// "no swizzle" canonical layout: 8x8 core matrices of 128 contiguous bytes, K-adjacent core
// matrices 128 B apart, 8-row groups 1024 B apart) in the exact order the MMA issuer consumes
// them, followed by an fp32 tail (biases + the two tiny heads evaluated on CUDA cores).
//
// GEMM steps ("layers") per 128-point tile; K-chunk sources: E = xyz encoding (63 ch + pad),
// A0..A3 = the four 64-wide K slices of the 256 activations, V = view-dir encoding (27 ch + pad).
//   step 0      pts_linears.0           N=256  K-chunks: E
//   step 1..4   pts_linears.1-4         N=256  A0 A1 A2 A3
//   step 5      pts_linears.5 (skip)    N=256  E A0 A1 A2 A3          (RH:105-106: cat[input_pts, h])
//   step 6..7   pts_linears.6-7         N=256  A0 A1 A2 A3            (+ alpha head in the step-7 epilogue)
//   step 8      feature_linear          N=256  A0 A1 A2 A3            (no activation, RH:110)
//   step 9      views_linears.0         N=128  A0 A1 A2 A3 V          (RH:111: cat[feature, input_views]; + rgb head)
// Chunk order inside a step: for each 128-wide half of N, all its K-chunks.
constexpr int CHUNK_ROWS = 128;
constexpr int CHUNK_K = 64;
constexpr int CHUNK_BYTES = CHUNK_ROWS * CHUNK_K * 2;  // 16384
constexpr int NUM_STEPS = 10;
constexpr int NUM_CHUNKS = 2 * 1 + 4 * 8 + 10 + 2 * 8 + 8 + 5;  // 73
constexpr int CHUNK_PAIR_BYTES = 2 * CHUNK_BYTES;              // [hi chunk | lo chunk], lo = fp16(W - fp16(W))
// Backward (data-gradient) chunks follow the forward ones: same [128 x 64] operand tiles of the TRANSPOSED
// weights (row = forward input feature, K = forward output feature), in the order the backward steps use them:
//   bstep 0  views_linears^T, view-dir rows   side N=32   K=128 (2 chunks)      -> dL/d(view-dir encoding)
//   bstep 1  views_linears^T, feature rows    N=256       K=128 (2 per half)    -> dL/dfeature
//   bstep 2  feature_linear^T                 N=256       K=256                 -> dL/dh7 (+ alpha head)
//   bstep 3,4  pts_linears.7^T, .6^T          N=256       K=256
//   bstep 5  pts_linears.5^T, xyz rows        side N=64   K=256 (4 chunks)      -> dL/d(xyz encoding), skip branch
//   bstep 6  pts_linears.5^T, h rows          N=256       K=256
//   bstep 7..10  pts_linears.4^T .. .1^T      N=256       K=256
//   bstep 11 pts_linears.0^T                  side N=64   K=256 (4 chunks)      -> dL/d(xyz encoding)
constexpr int NUM_BSTEPS = 12;
constexpr int NUM_BWD_CHUNKS = 2 + 4 + 3 * 8 + 4 + 5 * 8 + 4;  // 78
constexpr int BWD_CHUNK0 = NUM_CHUNKS;                         // index of the first backward chunk pair
// Mixed-precision forward chunks (NSR_FLAG_MIXED_F8) follow the backward ones: the 73 forward chunks again, every layer
// scaled by its own power of two 2^b (max|W| 2^b in [2^14, 2^15); the epilogue multiplies the accumulator by 2^-b).
//   steps < MIX_X3_STEPS and the chunks whose A operand is an encoding (step 5 kc 0, step 9 kc 4):
//       [fp16(W 2^b) | fp16 residual]                           -- the error-compensated split, as above
//   all other chunks:
//       [fp16(W 2^b) = Wh, 16 KB | e4m3(Wh 2^-11), 8 KB | e4m3(W 2^b - Wh), 8 KB]
//   the two 8-bit tiles are [128 x 64] K-major no-swizzle (core matrix = 8 rows x 16 bytes, 8-row groups 512 B apart): the
//   B operands of the kind::f8f6f4 correction products  e4m3(x_lo 2^11) . e4m3(Wh 2^-11)  +  e4m3(x_hi) . e4m3(W_lo).
constexpr int MIX_CHUNK0 = NUM_CHUNKS + NUM_BWD_CHUNKS;
constexpr int MIX_X3_STEPS = 3;
constexpr int MIX_XLO_SHIFT = 11;
constexpr int F8_TILE_BYTES = CHUNK_ROWS * CHUNK_K;            // 8192
constexpr int WEIGHT_BYTES = (2 * NUM_CHUNKS + NUM_BWD_CHUNKS) * CHUNK_PAIR_BYTES;  // 7,340,032

// fp32 tail (offsets in floats)
constexpr int TAIL_BIAS = 0;            // [10][256]  (step s bias at s*256; step 9 uses 128)
constexpr int TAIL_WALPHA = 2560;       // [256]
constexpr int TAIL_WRGB = 2816;         // [128][4]   (w_rgb[0][j], w_rgb[1][j], w_rgb[2][j], 0)
constexpr int TAIL_MISC = 3328;         // b_alpha, b_rgb[0..2]
constexpr int TAIL_MIXSCALE = 3336;     // [10] 2^-b of each step's mixed-precision chunks
constexpr int TAIL_FLOATS = 3360;
constexpr int TAIL_BYTES = TAIL_FLOATS * 4;  // 13440
// fp32 copy of the density branch (pts_linears.0-7 + alpha head) after the tail, for the coarse-pass refinement (refine.cu): layer l
// TRANSPOSED, [K_l][256] floats (thread j of a block reads W[k][j]: coalesced), then w_alpha [256].  K = 63, 256 x4, 319 (cat[xyz
// encoding, h], RH:105-106), 256, 256.
constexpr int REF_OFF = WEIGHT_BYTES + TAIL_BYTES;            // 7,353,472 (16-byte aligned)
__host__ __device__ constexpr int ref_layer_k(int l) { return l == 0 ? 63 : (l == 5 ? 319 : 256); }
__host__ __device__ constexpr int ref_layer_off(int l) {      // in floats; l = 8 -> w_alpha
  int o = 0;
  for (int i = 0; i < l; ++i) o += ref_layer_k(i) * 256;
  return o;
}
constexpr int REF_FLOATS = ref_layer_off(8) + 256;            // 491,264
constexpr int PACKED_BYTES = REF_OFF + REF_FLOATS * 4;

__host__ __device__ constexpr int step_n_halves(int s) { return s == 9 ? 1 : 2; }
__host__ __device__ constexpr int step_k_chunks(int s) { return s == 0 ? 1 : ((s == 5 || s == 9) ? 5 : 4); }
// mixed mode: does chunk (step, kc) carry fp8 correction tiles (else the fp16 hi/lo split)?
__host__ __device__ constexpr bool mix_chunk_is_f8(int s, int kc) { return s >= MIX_X3_STEPS && !(s == 5 && kc == 0) && !(s == 9 && kc == 4); }
// Issue order of a step's weight chunks.  Packed order is [half 0: kc 0..nk-1][half 1: kc 0..nk-1]; the kernels CONSUME a
// two-half step as  (h0, K early) (h1, K early) (h0, K late) (h1, K late), where "early" are the k0 chunks whose A operand is
// complete with the first half of the previous step's epilogue (an encoding, or activations 0..127) and "late" the ones that
// need its second half (activations 128..255).  That way the tensor pipe always has 2 x k0 chunks of work that do not depend
// on the epilogue still in flight, and the epilogue of accumulator 0 runs under (h1, K late).  Slot i -> (half, kc).
__host__ __device__ constexpr int step_k_early(int s) { return s == 0 ? 1 : (s == 5 ? 3 : 2); }
__host__ __device__ inline void issue_slot(int nk, int k0, int halves, int i, int& nh, int& kc) {
  if (halves == 1) {
    nh = 0;
    kc = i;
    return;
  }
  const int k1 = nk - k0;
  if (i < k0) {
    nh = 0;
    kc = i;
  } else if (i < 2 * k0) {
    nh = 1;
    kc = i - k0;
  } else if (i < 2 * k0 + k1) {
    nh = 0;
    kc = k0 + (i - 2 * k0);
  } else {
    nh = 1;
    kc = k0 + (i - 2 * k0 - k1);
  }
}
// slot after which accumulator half 0 of a two-half step is complete
__host__ __device__ inline int last_slot_half0(int nk, int k0) { return (nk - k0) > 0 ? 2 * k0 + (nk - k0) - 1 : k0 - 1; }
__host__ __device__ constexpr bool bstep_is_side(int b) { return b == 0 || b == 5 || b == 11; }
__host__ __device__ constexpr int bstep_n_halves(int b) { return bstep_is_side(b) ? 1 : 2; }
__host__ __device__ constexpr int bstep_k_chunks(int b) { return b <= 1 ? 2 : 4; }
__host__ __device__ constexpr int bstep_side_n(int b) { return b == 0 ? 32 : 64; }

// ----------------------------------------------------------------------------- saved ReLU sign bits (pose-gradient backward)
// With relu_mask != NULL the forward kernel stores one bit per ReLU output (is the fp16-rounded activation non-zero?) of the
// eight 256-wide layers and of the 128-wide views layer: 2176 bits = 272 B per point instead of ~5 KB of activations.  The
// backward kernel then skips its forward recompute (GEMM steps 0..9) altogether.  Per 128-point tile: 68 words x 128 rows of
// u32, word (l*8 + w) = columns [32w, 32w+32) of layer l (w < 8), word 64 + w = columns [32w, 32w+32) of the views layer
// (w < 4); inside a word bit j = column 2j, bit 16 + j = column 2j + 1 (the packed fp16 pair order of the operand words).
constexpr int MASK_WORDS = 8 * 8 + 4;                   // per point
constexpr int MASK_TILE_WORDS = MASK_WORDS * 128;       // 8704
constexpr int MASK_TILE_BYTES = MASK_TILE_WORDS * 4;    // 34,816
__host__ __device__ inline size_t relu_mask_bytes(int64_t n_points) { return size_t((n_points + 127) / 128) * MASK_TILE_BYTES; }

// ----------------------------------------------------------------------------- active set (two-tier evaluation, DESIGN.md "precision")
// A sample point whose density is certainly <= 0 has alpha == 0 exactly (RN:356: 1 - exp(-relu(sigma) dist)), so its weight is
// exactly 0 and neither its colour nor the value of its sigma reach any output of raw2outputs (RN:343-387) or any gradient.
// Tier 1 evaluates EVERY point with one fp16 MMA per product through pts_linears.0-7 and the alpha head only (sigma~); points with
// sigma~ <= -tau are certified empty (|sigma~ - sigma| is orders of magnitude below tau, and that is verified at run time on the
// points that do get both evaluations); the others -- the "active set", a compacted list of point indices -- are evaluated again
// with the default error-compensated arithmetic (tier 2), bit for bit what the dense fp16x3 pass computes for them.
//   ctrl block (16 x u32, zeroed by the host before the pass):
constexpr int AS_COUNT = 0;        // number of active points (atomic counter of the tier-1 kernel)
constexpr int AS_VMAX = 1;         // max |sigma~ - sigma| over the active points as float bits (atomicMax; NaN / inf sort above)
constexpr int AS_FORCE_DENSE = 2;  // set before the pass (from the coarse pass's active fraction): skip tier 1, evaluate everything
constexpr int AS_DENSE_FINAL = 3;  // set by the re-evaluation kernel: verification failed, every point was re-evaluated densely
constexpr int AS_CTRL_WORDS = 16;
constexpr int AS_CTRL_BYTES = 256; // one ctrl block, padded
//   active set buffer handed across the C ABI (nsr_active_set_bytes): [ctrl, 256 B][list: int32 per point]
__host__ __device__ inline size_t active_set_bytes(int64_t n_points) { return AS_CTRL_BYTES + size_t(n_points) * 4; }
struct TwoTierParams {
  float tau;            // certified empty: sigma~ <= -tau
  float verify_max;     // tier 2 must not move any active point's sigma by more than this, else everything is re-evaluated
  float force_frac;     // coarse active fraction above which the fine pass skips tier 1
};
const TwoTierParams& two_tier_params();
// role of an MLP launch with respect to the active set
enum : int { AS_ROLE_PLAIN = 0, AS_ROLE_TIER1 = 1, AS_ROLE_TIER2 = 2, AS_ROLE_REDO = 3 };

// ----------------------------------------------------------------------------- backward "dump" (operands of dL/dMLP, RN:691-707)
// With dump != NULL the backward kernel writes, for P = 128 * tiles points, every weight layer's input activations
// and pre-activation gradients as fp16, one array after the other:
//   EX [P,64] xyz encoding | EV [P,32] view-dir encoding | H0..H7 [P,256] post-ReLU | F [P,256] feature |
//   HV [P,128] views hidden | GV [P,128] dL/d(views pre-act) | GF [P,256] dL/dfeature | G0..G7 [P,256] dL/d(pts_linears.l pre-act)
// (gradients divided by one power-of-two `gscale`), and then the fp16 RESIDUALS of all of them (x - fp16(x)) in the same order,
// dump_lo(P) bytes further on: the weight-gradient GEMMs run the same error-compensated product as the forward pass
// (G_hi H_hi + G_lo H_hi + G_hi H_lo), which keeps dL/dW within 1e-3 of fp32 autograd (single fp16 operands: 1.4e-3 on the first
// layers of the fine network).  Each array is stored tile by tile (128 points) in the UMMA
// MN-major no-swizzle canonical layout, i.e. 8x8 blocks [8 points][8 features] of 128 contiguous bytes, feature blocks
// adjacent, point blocks (W/8)*128 bytes apart, so the weight-gradient kernel can feed slices of it to tcgen05.mma
// as BOTH operands of dW = G^T H with the points as the K dimension -- no transposition anywhere.
__host__ __device__ inline size_t dump_off_ex(size_t) { return 0; }
__host__ __device__ inline size_t dump_off_ev(size_t P) { return P * 128; }
__host__ __device__ inline size_t dump_off_h(size_t P, int l) { return P * 192 + size_t(l) * P * 512; }   // l = 8 -> F
__host__ __device__ inline size_t dump_off_hv(size_t P) { return dump_off_h(P, 9); }
__host__ __device__ inline size_t dump_off_gv(size_t P) { return dump_off_hv(P) + P * 256; }
__host__ __device__ inline size_t dump_off_gf(size_t P) { return dump_off_gv(P) + P * 256; }
__host__ __device__ inline size_t dump_off_g(size_t P, int l) { return dump_off_gf(P) + P * 512 + size_t(l) * P * 512; }
__host__ __device__ inline size_t dump_lo(size_t P) { return dump_off_g(P, 8); }        // offset of the residual copy of every array
__host__ __device__ inline size_t dump_total(size_t P) { return 2 * dump_lo(P); }
// byte offset inside a [P, W] array of the 16-byte group (tile, row, feature group fg = feature / 8)
__host__ __device__ inline size_t dump_blocked_off(int tile, int row, int W, int fg) {
  return size_t(tile) * (size_t(128) * W * 2) + size_t(row >> 3) * (W / 8) * 128 + size_t(fg) * 128 + size_t(row & 7) * 16;
}

// ----------------------------------------------------------------------------- host-side plumbing (api.cu)
void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch();
int current_device_sms(int* sms);
int ensure_dynamic_smem(const void* func, int bytes);

// ray_stage.cu
int launch_coarse_z(const float* rays, int64_t n, int S, uint32_t flags, const float* t_rand, float* z, cudaStream_t st);
int launch_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld, int64_t n, int S, uint32_t flags,
                       float* rgb, float* disp, float* acc, float* weights, float* depth, cudaStream_t st);
int launch_sample_pdf(const float* bins, const float* weights, int64_t n, int B, int N, const float* u, float* out,
                      cudaStream_t st);
int launch_resample_merge(const float* z, const float* w, int64_t n, int S, int Ni, const float* u, float* z_fine,
                          float* z_samples, float* z_std, cudaStream_t st, const uint32_t* ctrl_coarse = nullptr,
                          uint32_t* ctrl_fine = nullptr, uint32_t force_count = 0);
int launch_make_rays(int H, int W, const float* K9, const float* c2w12, float near_, float far_, float* rays,
                     cudaStream_t st);
int launch_pack_rays(const float* o, const float* d, int64_t n, float near_, float far_, float* rays, cudaStream_t st);
// refine.cu: fp32 re-evaluation of the coarse density where hierarchical sampling is ill-conditioned (DESIGN.md "coarse refinement")
size_t refine_workspace_bytes(int64_t n_rays);
void set_refine_tau_limit(float tau);
void set_refine_sigma_hi(float sigma_hi);
int launch_coarse_refine(const float* rays, const float* z, int64_t n, int S, const void* packed, float* raw, void* workspace, cudaStream_t st);
// image_stage.cu
int launch_to8b(const float* x, int64_t n, uint8_t* out, cudaStream_t st);
int launch_make_rays_dev(int H, int W, const float* K9, const float* c2w_dev, int ld_c2w, float near_, float far_, float* rays,
                         cudaStream_t st);
int launch_c2w_grad(int W, const float* K9, const float* rays, const float* d_rays, const int32_t* pixel_idx, int64_t n,
                    float* d_c2w, int accumulate, double* partials, cudaStream_t st);
size_t c2w_grad_workspace_bytes();
// train_stage.cu
struct AdamJob {
  float* p;
  const float* g;
  float* m;
  float* v;
  int n;
};
struct AdamJobs {
  AdamJob j[2 * 2 * NSR_NET_NUM_TENSORS];
  int count;
};
int launch_uniform(uint64_t seed, uint32_t stream, float* out, int64_t count, cudaStream_t st);
int launch_sigma_noise(uint64_t seed, uint32_t stream, float* raw, int64_t n_points, float std, cudaStream_t st);
int launch_mse_grad(const float* x, const float* y, int64_t count, float* d_x, float* loss, cudaStream_t st);
int launch_adam(const AdamJobs& jobs, float beta1, float beta2, float lr, float eps, int64_t step, cudaStream_t st);
// mlp_forward.cu
int launch_pack_net(const float* const* weights, const float* const* biases, void* packed, cudaStream_t st);
int launch_mlp_forward(const float* rays, const float* z_or_pts, int64_t n, int S, const void* packed, uint32_t flags,
                       float* raw, cudaStream_t st, uint32_t* relu_mask = nullptr, void* dump = nullptr, int role = AS_ROLE_PLAIN,
                       void* active_set = nullptr);
int set_tier1_pair(int enabled);   // tier 1 as CTA pairs (cta_group::2); returns the previous setting
// mlp_backward.cu: d_raw [n,S,4] -> d_pts [n,S,8] = (dL/dpoint[3], 0, dL/dviewdir[3], 0) per sample
int launch_mlp_backward(const float* rays, const float* z, int64_t n, int S, const void* packed, const float* d_raw,
                        float* d_pts, void* dump, const float* gscale, cudaStream_t st, const uint32_t* relu_mask = nullptr,
                        const void* active_set = nullptr);
// wgrad.cu: dL/dW, dL/db of one network pass from the dump; accumulates (atomicAdd) into dW[12], dB[12] (fp32, reference shapes)
int launch_weight_grads(const void* dump, const float* d_raw, int64_t n_points, const float* gscale, float* const* dW,
                        float* const* dB, cudaStream_t st);
size_t mlp_dump_bytes(int64_t n_points);
// ray_stage.cu
int launch_raw2outputs_backward(const float* raw, const float* z, const float* rays, int64_t n, int S, uint32_t flags,
                                const float* d_rgb, float* d_raw, float* d_dnorm, float* gmax, cudaStream_t st);
int launch_ray_grad_reduce(const float* rays, const float* z, const float* d_pts, const float* d_dnorm, int64_t n, int S,
                           float* d_rays, cudaStream_t st);

}  // namespace nsr
