// Positional encoding + 8x256 NeRF MLP (RH:18-66, RH:99-122, RN:26-40) as one tcgen05 kernel.
//
// A CTA owns one 128-point tile at a time (UMMA M=128: row r == TMEM lane r == thread r of the
// epilogue warps) and runs all ten GEMM steps of the network on it without touching HBM.
//
//   TMEM (512 columns)   ACC0 [0,128) ACC1 [128,256)  fp32 accumulators of the two 128-wide halves of N
//                        AHI [256,384) ALO [384,512)  the 256 activations as fp16 hi / lo words (A operand,
//                                                     tcgen05.mma "TS" form: A from TMEM, B from shared memory)
//   shared memory        xyz / view-dir encodings of the current and the next tile (A operand, "SS" form),
//                        a ring of weight chunks streamed by the TMA engine (cp.async.bulk + mbarrier),
//                        the fp32 tail (biases, alpha / rgb heads)
//
// Precision.  fp16 operands with fp32 accumulation are NOT enough for the 1e-3 parity bar on sharp scenes
// (DESIGN.md "precision"), so by default every product runs as an error-compensated split:
//     x = x_hi + x_lo,  W = W_hi + W_lo  (fp16 each),   x.W ~= x_hi.W_hi + x_lo.W_hi + x_hi.W_lo
// (three MMAs per k-step into the same fp32 accumulator; the dropped x_lo.W_lo term is ~2^-22).
// SPLIT=1 (NSR_FLAG_FAST_FP16) issues only the first term.
//
// Pipeline.  A two-half step consumes its weight chunks as (h0, K early) (h1, K early) (h0, K late) (h1, K late)
// (common.cuh issue_slot): "early" chunks need only the first half of the previous step's epilogue (an encoding, or
// activations 0..127), so the tensor pipe always holds work that does not depend on the epilogue still in flight;
// accumulator 0 completes three quarters into the step and its epilogue -- which may overwrite A[K 0..127] at once,
// every chunk that read it having been issued before accumulator 0's last ones -- runs under (h1, K late);
// accumulator 1's epilogue runs under the next step's early chunks.  Encoder warps prepare tile i+1's encodings
// during tile i.
//
// Warp roles (384 threads, mlp_common.cuh): 0-7 epilogue (warp w drains TMEM lane quadrant w & 3, columns
// [64 (w >> 2), +64) of each accumulator half), 8-9 encoders (two rows per thread), 10 MMA issuer (the whole warp runs
// the uniform control flow, one elected lane issues), 11 weight producer (one lane drives the TMA engine).
//
// Two-tier evaluation (common.cuh "active set"): the CLASSIFY instantiation is tier 1 -- one fp16 MMA per product through
// pts_linears.0-7 and the alpha head only, writes sigma~ and appends every point that is not certainly empty to the active
// list; the default instantiation run over that list is tier 2 (rows of a tile are then arbitrary points: every row of
// the tile is independent, so a point's result does not depend on which tile it lands in).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp8.h>

#include "mlp_common.cuh"

namespace nsr {

// ----------------------------------------------------------------------------- weight packing
// canonical no-swizzle K-major offset (bytes) of element (row, k) inside a [128 x 64] fp16 chunk
__host__ __device__ __forceinline__ int chunk_off(int row, int k) { return (row >> 3) * 1024 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

struct NetPtrs {
  const float* w[NSR_NET_NUM_TENSORS];
  const float* b[NSR_NET_NUM_TENSORS];
};

// chunk index -> (step, n_half, k_chunk)
__device__ __forceinline__ void chunk_decode(int c, int& step, int& nh, int& kc) {
  int s = 0;
  for (; s < NUM_STEPS; ++s) {
    const int cnt = step_n_halves(s) * step_k_chunks(s);
    if (c < cnt) break;
    c -= cnt;
  }
  step = s;
  nh = c / step_k_chunks(s);
  kc = c % step_k_chunks(s);
}

// backward chunk index (0-based within the backward list) -> (bstep, n_half, k_chunk)
__device__ __forceinline__ void bwd_chunk_decode(int c, int& bstep, int& nh, int& kc) {
  int b = 0;
  for (; b < NUM_BSTEPS; ++b) {
    const int cnt = bstep_n_halves(b) * bstep_k_chunks(b);
    if (c < cnt) break;
    c -= cnt;
  }
  bstep = b;
  nh = c / bstep_k_chunks(b);
  kc = c % bstep_k_chunks(b);
}

// forward weight of GEMM step `step` (common.cuh): output row n, K chunk kc, column kl of the chunk (0 where padded)
__device__ __forceinline__ float fwd_weight(const NetPtrs& p, int step, int n, int kc, int kl) {
  if (step == 0) return kl < 63 ? p.w[0][n * 63 + kl] : 0.f;
  if (step == 5) {  // cat[input_pts(63), h(256)]  RH:106
    if (kc == 0) return kl < 63 ? p.w[5][n * 319 + kl] : 0.f;
    return p.w[5][n * 319 + 63 + (kc - 1) * 64 + kl];
  }
  if (step <= 7) return p.w[step][n * 256 + kc * 64 + kl];
  if (step == 8) return p.w[9][n * 256 + kc * 64 + kl];  // feature_linear
  // views_linears.0 on cat[feature(256), dirs(27)]  RH:111
  if (kc < 4) return p.w[8][n * 283 + kc * 64 + kl];
  return kl < 27 ? p.w[8][n * 283 + 256 + kl] : 0.f;
}

// canonical no-swizzle K-major offset (bytes) of element (row, k) inside a [128 x 64] 8-bit tile
__host__ __device__ __forceinline__ int f8_off(int row, int k) { return (row >> 3) * 512 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15); }

// Power-of-two exponent b of a step's weight matrix for the mixed-precision chunks: max|W| 2^b in [2^14, 2^15).
// Every thread of the block returns the same value.  `red` = 8 floats of shared memory.
__device__ int layer_shift(const NetPtrs& p, int step, float* red) {
  const int tensor = step <= 7 ? step : (step == 8 ? 9 : 8);
  const int count = step == 0 ? 256 * 63 : (step == 5 ? 256 * 319 : (step == 9 ? 128 * 283 : 65536));
  const float* w = p.w[tensor];
  float m = 0.f;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {   // every count is a multiple of 4
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll 4
    for (int i = threadIdx.x; i < count / 4; i += blockDim.x) {
      const float4 v = w4[i];
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  if (!(m > 0.f) || !isfinite(m)) return 0;
  int e;
  frexpf(m, &e);  // m = f 2^e, f in [0.5, 1)
  const int b = 15 - e;
  return b < -60 ? -60 : (b > 60 ? 60 : b);
}

__global__ void pack_net_kernel(NetPtrs p, uint8_t* __restrict__ out) {
  const int c = blockIdx.x;
  if (c >= NUM_CHUNKS && c < NUM_CHUNKS + NUM_BWD_CHUNKS) {
    // transposed weights for the data-gradient pass: element (row n, k) = W_forward[k][n]
    int bstep, nh, kc;
    bwd_chunk_decode(c - NUM_CHUNKS, bstep, nh, kc);
    __half* hi = reinterpret_cast<__half*>(out + size_t(c) * CHUNK_PAIR_BYTES);
    __half* lo = reinterpret_cast<__half*>(out + size_t(c) * CHUNK_PAIR_BYTES + CHUNK_BYTES);
    for (int e = threadIdx.x; e < CHUNK_ROWS * CHUNK_K; e += blockDim.x) {
      const int kl = e / CHUNK_ROWS, nl = e % CHUNK_ROWS;   // consecutive threads: consecutive n = contiguous in the source row
      const int n = nh * 128 + nl, k = kc * 64 + kl;
      float v = 0.f;
      switch (bstep) {
        case 0: if (nl < 27) v = p.w[8][k * 283 + 256 + nl]; break;   // views_linears [128 x 283], dirs columns
        case 1: v = p.w[8][k * 283 + n]; break;                        // views_linears, feature columns
        case 2: v = p.w[9][k * 256 + n]; break;                        // feature_linear
        case 3: v = p.w[7][k * 256 + n]; break;
        case 4: v = p.w[6][k * 256 + n]; break;
        case 5: if (nl < 63) v = p.w[5][k * 319 + nl]; break;          // pts_linears.5 [256 x 319], xyz columns
        case 6: v = p.w[5][k * 319 + 63 + n]; break;
        case 7: v = p.w[4][k * 256 + n]; break;
        case 8: v = p.w[3][k * 256 + n]; break;
        case 9: v = p.w[2][k * 256 + n]; break;
        case 10: v = p.w[1][k * 256 + n]; break;
        default: if (nl < 63) v = p.w[0][k * 63 + nl]; break;          // pts_linears.0 [256 x 63]
      }
      const __half h = __float2half_rn(v);
      hi[chunk_off(nl, kl) >> 1] = h;
      lo[chunk_off(nl, kl) >> 1] = __float2half_rn(v - __half2float(h));
    }
  } else if (c < NUM_CHUNKS) {
    int step, nh, kc;
    chunk_decode(c, step, nh, kc);
    __half* hi = reinterpret_cast<__half*>(out + size_t(c) * CHUNK_PAIR_BYTES);
    __half* lo = reinterpret_cast<__half*>(out + size_t(c) * CHUNK_PAIR_BYTES + CHUNK_BYTES);
    for (int e = threadIdx.x; e < CHUNK_ROWS * CHUNK_K; e += blockDim.x) {
      const int nl = e / CHUNK_K, kl = e % CHUNK_K;
      const float v = fwd_weight(p, step, nh * 128 + nl, kc, kl);
      const __half h = __float2half_rn(v);
      hi[chunk_off(nl, kl) >> 1] = h;
      lo[chunk_off(nl, kl) >> 1] = __float2half_rn(v - __half2float(h));
    }
  } else if (c >= MIX_CHUNK0 && c < MIX_CHUNK0 + NUM_CHUNKS) {
    // mixed-precision forward chunks: layer scaled by 2^b, fp8 correction tiles for the late layers (common.cuh)
    int step, nh, kc;
    chunk_decode(c - MIX_CHUNK0, step, nh, kc);
    __shared__ float s_red[8];
    const float scale = exp2f(float(layer_shift(p, step, s_red)));
    uint8_t* base = out + size_t(c) * CHUNK_PAIR_BYTES;
    __half* hi = reinterpret_cast<__half*>(base);
    __half* lo = reinterpret_cast<__half*>(base + CHUNK_BYTES);
    uint8_t* h8 = base + CHUNK_BYTES;
    uint8_t* l8 = base + CHUNK_BYTES + F8_TILE_BYTES;
    const bool f8 = mix_chunk_is_f8(step, kc);
    for (int e = threadIdx.x; e < CHUNK_ROWS * CHUNK_K; e += blockDim.x) {
      const int nl = e / CHUNK_K, kl = e % CHUNK_K;
      const float v = fwd_weight(p, step, nh * 128 + nl, kc, kl) * scale;
      const __half h = __float2half_rn(v);
      hi[chunk_off(nl, kl) >> 1] = h;
      if (f8) {
        h8[f8_off(nl, kl)] = __nv_cvt_float_to_fp8(__half2float(h) * (1.f / float(1 << MIX_XLO_SHIFT)), __NV_SATFINITE, __NV_E4M3);
        l8[f8_off(nl, kl)] = __nv_cvt_float_to_fp8(v - __half2float(h), __NV_SATFINITE, __NV_E4M3);
      } else {
        lo[chunk_off(nl, kl) >> 1] = __float2half_rn(v - __half2float(h));
      }
    }
  } else if (c >= MIX_CHUNK0 + NUM_CHUNKS + 1 + NUM_STEPS) {  // fp32 transposed copy of the density branch (common.cuh REF_*)
    const int l = c - (MIX_CHUNK0 + NUM_CHUNKS + 1 + NUM_STEPS);
    float* dst = reinterpret_cast<float*>(out + REF_OFF) + ref_layer_off(l);
    if (l == 8) {
      for (int j = threadIdx.x; j < 256; j += blockDim.x) dst[j] = p.w[10][j];
    } else {
      const int K = ref_layer_k(l);
      for (int e = threadIdx.x; e < K * 256; e += blockDim.x) {
        const int k = e / 256, j = e % 256;
        dst[e] = p.w[l][j * K + k];
      }
    }
  } else if (c == MIX_CHUNK0 + NUM_CHUNKS) {  // fp32 tail
    float* t = reinterpret_cast<float*>(out + WEIGHT_BYTES);
    for (int i = threadIdx.x; i < TAIL_FLOATS; i += blockDim.x) {
      float v = 0.f;
      if (i < TAIL_WALPHA) {
        const int s = i / 256, j = i % 256;
        if (s <= 7) v = p.b[s][j];
        else if (s == 8) v = p.b[9][j];
        else if (j < 128) v = p.b[8][j];
      } else if (i < TAIL_WRGB) {
        v = p.w[10][i - TAIL_WALPHA];
      } else if (i < TAIL_MISC) {
        const int j = (i - TAIL_WRGB) / 4, ch = (i - TAIL_WRGB) % 4;
        if (ch < 3) v = p.w[11][ch * 128 + j];
      } else {
        const int j = i - TAIL_MISC;
        if (j == 0) v = p.b[10][0];
        else if (j < 4) v = p.b[11][j - 1];
      }
      if (i < TAIL_MIXSCALE || i >= TAIL_MIXSCALE + NUM_STEPS) t[i] = v;   // the scales are written by the blocks below
    }
  } else if (c < MIX_CHUNK0 + NUM_CHUNKS + 1 + NUM_STEPS) {  // one block per GEMM step: 2^-b of its mixed-precision chunks
    const int step = c - (MIX_CHUNK0 + NUM_CHUNKS + 1);
    __shared__ float s_red[8];
    const int b = layer_shift(p, step, s_red);
    if (threadIdx.x == 0) reinterpret_cast<float*>(out + WEIGHT_BYTES)[TAIL_MIXSCALE + step] = exp2f(float(-b));
  }
}

int launch_pack_net(const float* const* weights, const float* const* biases, void* packed, cudaStream_t st) {
  NetPtrs p;
  for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) {
    p.w[i] = weights[i];
    p.b[i] = biases[i];
  }
  pack_net_kernel<<<2 * NUM_CHUNKS + NUM_BWD_CHUNKS + 1 + NUM_STEPS + 9, 256, 0, st>>>(p, static_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_net_kernel");
}

template <int SPLIT>
struct Cfg {
  static constexpr bool kSplit = SPLIT >= 2;   // operands carried as fp16 hi + a residual (3: fp16 everywhere, 2: mixed)
  static constexpr bool kMixed = SPLIT == 2;   // residual products of the late layers in e4m3 (common.cuh, MIX_*)
  static constexpr int STAGE_BYTES = kSplit ? CHUNK_PAIR_BYTES : CHUNK_BYTES;
  static constexpr int ENC_BYTES = 128 * 64 * 2;   // [128 x 64] fp16, 8-row groups 1024 B apart
  static constexpr int DIR_BYTES = 128 * 32 * 2;   // [128 x 32] fp16, 8-row groups 512 B apart
  static constexpr int INBUF_BYTES = (ENC_BYTES + DIR_BYTES) * (kSplit ? 2 : 1);
  static constexpr int OFF_ENC_HI = 0;
  static constexpr int OFF_DIR_HI = ENC_BYTES;
  static constexpr int OFF_ENC_LO = ENC_BYTES + DIR_BYTES;
  static constexpr int OFF_DIR_LO = 2 * ENC_BYTES + DIR_BYTES;
  // One buffer each: the xyz encoding of tile i is last read by step 5, the view-dir encoding by step 9,
  // so the encoder warps refill them for tile i+1 during steps 6..9 / 0..8 (enc_free barriers).
  static constexpr int SM_INBUF = 0;
  static constexpr int SM_RING = INBUF_BYTES;
  static constexpr int SM_BUDGET = 227 * 1024;
  static constexpr int SM_XCH_BYTES = 128 * 16;    // per-row partial (r, g, b, sigma) handed between the two column halves
  static constexpr int SM_BAR_BYTES = 512;         // up to 2 x 20 ring barriers (CTA-pair tier 1) + 9 others + the TMEM slot
  static constexpr int STAGES_FIT = (SM_BUDGET - SM_RING - TAIL_BYTES - SM_XCH_BYTES - SM_BAR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 10 ? 10 : STAGES_FIT;
  static constexpr int SM_TAIL = SM_RING + STAGES * STAGE_BYTES;
  static constexpr int SM_XCH = SM_TAIL + TAIL_BYTES;
  static constexpr int SM_BAR = SM_XCH + SM_XCH_BYTES;
  static constexpr int SM_TOTAL = SM_BAR + SM_BAR_BYTES;
  static_assert(STAGES >= 3, "weight ring too shallow");
};

struct MlpArgs {
  const float* rays;
  const float* z_or_pts;
  const uint8_t* packed;
  float* raw;
  int64_t n_points;
  int S;
  uint32_t flags;
  int num_tiles;
  uint32_t* relu_mask;        // optional: sign bits of every ReLU output for the backward pass (common.cuh MASK_*)
  int experiment;             // debug (NSR_EXPERIMENT): 1 = the producer re-arms the ring without copying after its first lap
                              // (wrong results; isolates the cost of streaming the weights from L2)
  uint8_t* dump;              // optional: the activation half of the weight-gradient dump (common.cuh): EX, EV, H0..H7, F, HV as fp16
  unsigned long long* trace;  // debug (NSR_TRACE_FILE): clock64 stamps of CTA 0's first tiles, [tile][step][16]
  // two-tier evaluation (common.cuh "active set"); ctrl == NULL: plain dense launch
  uint32_t* ctrl;             // control block of this pass
  int32_t* list;              // active point indices (tier 1 appends, tier 2 reads)
  int role;                   // AS_ROLE_*
  float tau;                  // tier 1: a point with sigma~ <= -tau is certified empty
  uint32_t verify_bits;       // re-evaluation launch: runs iff ctrl[AS_VMAX] > verify_bits
};

// debug experiments (NSR_EXPERIMENT, debug-hook builds only; results are garbage, timings isolate one cost each):
//   1 no weight streaming after the first lap   2 epilogue skips its TMEM stores   3 epilogue skips the conversion arithmetic
//   4 encoders skip sincosf   5 no MMA is issued (commits only)   6 only every second weight chunk is streamed   100+g: grid limited to g CTAs
#ifdef NSR_DEBUG_HOOKS
#define NSR_EXP(n) (a.experiment == (n))
#else
#define NSR_EXP(n) false
#endif

#define NSR_TR(tl, step, slot)                                                                         \
  do {                                                                                                 \
    if (a.trace != nullptr && blockIdx.x == 0 && (tl) < 4) a.trace[((tl) * 10 + (step)) * 16 + (slot)] = clock64(); \
  } while (0)

// ----------------------------------------------------------------------------- the kernel
// SAVE: also write the ReLU sign bits (a.relu_mask) and, when a.dump != NULL, the activation half of the weight-gradient dump
// for the backward pass.  A separate instantiation, so the plain render kernel carries none of that code.
// CLASSIFY: tier 1 of the two-tier evaluation (SPLIT = 1 only): GEMM steps 0..7 + alpha head, sigma~ and the active list out.
// Warp roles: mlp_common.cuh (12 warps).  The tier-1 kernel (CLASSIFY) issues one MMA per product, so its step is bounded by the
// latency of the epilogue, not by the tensor pipe (with no MMA issued at all it still took 82 % of its time): it runs 16 epilogue
// warps, 8 per accumulator half, so that the two halves drain concurrently and each scheduler has four warps to hide latency behind.
template <bool CLASSIFY>
struct Roles {
  static constexpr int EPI_WARPS = CLASSIFY ? 16 : 8;
  static constexpr int ENC0 = EPI_WARPS, MMA = EPI_WARPS + 2, PROD = EPI_WARPS + 3;
  static constexpr int THREADS = (EPI_WARPS + 4) * 32;
};
static_assert(Roles<false>::THREADS == MLP_THREADS && Roles<false>::ENC0 == ENC_WARP0 && Roles<false>::MMA == MMA_WARP &&
              Roles<false>::PROD == PROD_WARP, "mlp_common.cuh warp roles");

// PAIR (tier 1 only): the kernel runs as clusters of two CTAs on the two SMs of a TPC and issues tcgen05.mma.cta_group::2 -- one
// instruction stream (the leader's, cluster rank 0) drives both tensor cores with M = 256: each CTA keeps its own tile (its TMEM, its
// encodings, its epilogue) but only HALF of every weight chunk (64 of the 128 output rows, 8 KB), so both the TMA writes into shared
// memory and the tensor core's B reads out of it halve.  That is what bounded the single-CTA kernel: 64 B/cycle written + 64 B/cycle
// read is all the shared memory has.  Barriers: weights / accumulators / encodings are released by multicast
// commits (both CTAs); the epilogue and encoder warps of BOTH CTAs arrive, one lane per warp, on the LEADER's barriers; the
// follower's otherwise idle MMA warp relays "my half of the chunk has landed" to the leader's full barrier.
template <int SPLIT, bool SAVE, bool CLASSIFY, bool PAIR = false>
__global__ void __launch_bounds__(Roles<CLASSIFY>::THREADS, 1) nerf_mlp_kernel(MlpArgs a) {
  using C = Cfg<SPLIT>;
  using R = Roles<CLASSIFY>;
  static_assert(!PAIR || CLASSIFY, "the CTA-pair variant is built for the tier-1 kernel");
  constexpr int kStageBytes = PAIR ? C::STAGE_BYTES / 2 : C::STAGE_BYTES;
  constexpr int kStages = PAIR ? 2 * C::STAGES : C::STAGES;   // half-size stages: the same ring memory holds twice as many chunks
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;   // == blockIdx.x & 1
  constexpr bool kSplit = C::kSplit;
  constexpr bool kMixed = C::kMixed;
  constexpr int kSteps = CLASSIFY ? 8 : NUM_STEPS;   // GEMM steps per tile
#ifndef NSR_ACCFREE_SPLIT
#define NSR_ACCFREE_SPLIT 0
#endif
  // hand ACC1 back to the MMA warp as soon as it is in registers (separate barrier) instead of with the activations it turns into
  constexpr bool kAccFree = (SPLIT == 1) || NSR_ACCFREE_SPLIT;
  static_assert(!CLASSIFY || (SPLIT == 1 && !SAVE), "tier 1 is the single-pass fp16 arithmetic, nothing saved");
  // ---- role with respect to the active set: every thread of every CTA takes the same decision from the control block, which no
  // kernel of this launch's role modifies in the words read here
  int num_tiles = a.num_tiles;
  int n_act = 0;
  bool use_list = false;
  if (a.ctrl != nullptr) {
    const uint32_t force = a.ctrl[AS_FORCE_DENSE];
    if (a.role == AS_ROLE_TIER1) {
      if (force) return;
    } else if (a.role == AS_ROLE_TIER2) {
      if (!force) {
        use_list = true;
        n_act = int(a.ctrl[AS_COUNT]);
        num_tiles = (n_act + 127) >> 7;
        if (num_tiles == 0) return;
      }
    } else if (a.role == AS_ROLE_REDO) {
      if (force || !(a.ctrl[AS_VMAX] > a.verify_bits)) return;
    }
  }
  // point index of (tile, row), or -1 for the padding rows of the last tile
  auto point_of = [&](int tile, int row) -> int64_t {
    const int64_t q = int64_t(tile) * 128 + row;
    if (use_list) return q < n_act ? int64_t(a.list[q]) : int64_t(-1);
    return q < a.n_points ? q : int64_t(-1);
  };
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sRing = smem + C::SM_RING;
  const float* sTail = reinterpret_cast<const float*>(smem + C::SM_TAIL);
  float4* sXch = reinterpret_cast<float4*>(smem + C::SM_XCH);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);  // [STAGES]   producer -> MMA (tx bytes)
  uint64_t* empty = full + kStages;                                // [STAGES]   MMA commit -> producer
  uint64_t* acc_ready = empty + kStages;                           // [2]        MMA commit -> epilogue
  uint64_t* a_ready = acc_ready + 2;                               // [2]        epilogue (256) -> MMA
  uint64_t* enc_ready = a_ready + 2;                               // [0] xyz encoding, [1] view-dir encoding: encoders -> MMA
  uint64_t* enc_free = enc_ready + 2;                              // [0] after step 5, [1] after step 9: MMA commit -> encoders
  uint64_t* acc_free1 = enc_free + 2;                              // [1]        epilogue (256): ACC1 is in registers -> MMA (kAccFree)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free1 + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t dumpP = size_t(a.num_tiles) * 128;   // (dump: dense launches only)   // rows of the optional dump
  // PAIR: the two CTAs of a pair take tiles (2i, 2i+1) together, so the tile range is rounded up to even (a tile beyond the last
  // one has no valid row: point_of returns -1 everywhere)
  const int tile_end = PAIR ? ((num_tiles + 1) & ~1) : num_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], (PAIR && crank == 0) ? 2 : 1);   // PAIR leader: own producer (+ bytes) and the follower's relay
      mbar_init(&empty[s], 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&acc_ready[h], 1);
      mbar_init(&a_ready[h], PAIR ? 2 * R::EPI_WARPS : (CLASSIFY ? R::EPI_WARPS * 32 : EPI_THREADS));   // PAIR: one arrival per warp, both CTAs
      mbar_init(&enc_ready[h], PAIR ? 4 : ENC_THREADS);
      mbar_init(&enc_free[h], 1);
    }
    mbar_init(acc_free1, PAIR ? 2 * R::EPI_WARPS : (CLASSIFY ? R::EPI_WARPS * 32 : EPI_THREADS));
    fence_mbar_init();
  }
  if (warp == R::MMA) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  for (int i = tid; i < TAIL_FLOATS; i += R::THREADS)
    reinterpret_cast<float*>(smem + C::SM_TAIL)[i] = reinterpret_cast<const float*>(a.packed + WEIGHT_BYTES)[i];
  if constexpr (CLASSIFY) {   // tier 1 adds its biases in packed fp16: the table lives where the (unused) view-dir encoding would
    static_assert(C::DIR_BYTES >= 8 * 256 * 2, "fp16 bias table");
    for (int i = tid; i < 8 * 128; i += R::THREADS) {
      const float2 b = reinterpret_cast<const float2*>(a.packed + WEIGHT_BYTES)[i];   // TAIL_BIAS == 0
      reinterpret_cast<uint32_t*>(smem + C::SM_INBUF + C::OFF_DIR_HI)[i] = pack_f16x2(b.x, b.y);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / multicast commit
  tc_fence_after_sync();
  if (*tmem_slot != 0u) __trap();  // the whole TMEM of the SM is ours: the allocation must start at 0
  // PAIR: addresses of the LEADER's barriers in the cluster window (the leader maps onto itself)
  [[maybe_unused]] uint32_t lead_a_ready[2] = {0u, 0u}, lead_acc_free1 = 0u, lead_enc_ready0 = 0u;
  if constexpr (PAIR) {
    lead_a_ready[0] = mapa_u32(smem_u32(&a_ready[0]), 0);
    lead_a_ready[1] = mapa_u32(smem_u32(&a_ready[1]), 0);
    lead_acc_free1 = mapa_u32(smem_u32(acc_free1), 0);
    lead_enc_ready0 = mapa_u32(smem_u32(&enc_ready[0]), 0);
  }

  if (warp == R::PROD) {
    // ===================================================================== weight producer (TMA engine)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      bool first_lap = true;
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        int base = kMixed ? MIX_CHUNK0 : 0;          // first packed chunk of the step
        for (int step = 0; step < kSteps; ++step) {
          const int nk = step_k_chunks(step), nhs = step_n_halves(step);
          for (int i = 0; i < nk * nhs; ++i) {         // in the order the MMA warp consumes them (common.cuh issue_slot)
            int nh, kc;
            issue_slot(nk, step_k_early(step), nhs, i, nh, kc);
            if (!first_lap) mbar_wait(&empty[stage], phase ^ 1);
            if ((NSR_EXP(1) || (NSR_EXP(6) && (i & 1))) && !first_lap) {   // 6: every second chunk only
              mbar_arrive(&full[stage]);
            } else {
              // (PAIR: this CTA's half of the chunk's output rows -- the packed chunk keeps 8-row groups 1024 B apart, so rows
              // [64 r, 64 r + 64) are its bytes [8192 r, 8192 r + 8192))
              mbar_arrive_expect_tx(&full[stage], kStageBytes);
              bulk_g2s(sRing + stage * kStageBytes, a.packed + size_t(base + nh * nk + kc) * CHUNK_PAIR_BYTES + (PAIR ? crank * kStageBytes : 0u),
                       kStageBytes, &full[stage]);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
              first_lap = false;
            }
          }
          base += nk * nhs;
        }
      }
    }
  } else if (warp == R::MMA && CLASSIFY) {
    // ===================================================================== MMA issuer of the tier-1 kernel: a static schedule
    // With one MMA per product a chunk is 4 instructions = 256 tensor-pipe cycles, and the generic issue loop below (slot
    // arithmetic, ring bookkeeping, ~90 SASS instructions and three barrier probes per chunk) takes longer than that: measured
    // 98 cycles per MMA for the bare loop (tools/mma_rate_probe1.cu) against 64 for the same instructions issued back to back
    // (tools/operand_reuse_probe.cu).  Tier 1 consumes 60 chunks per tile through a 10-stage ring and runs 8 steps, so every ring
    // stage, barrier parity and descriptor offset of a tile is a compile-time constant: the whole tile is unrolled.
    static_assert(!CLASSIFY || (C::STAGES == 10 && kSteps == 8 && !kSplit), "static tier-1 schedule");
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, 128);
    constexpr uint32_t HI_B = sdesc_hi(1024);
    const uint32_t ring_lo = sdesc_lo(smem_u32(sRing), 128);
    const uint32_t enc_hi = sdesc_lo(smem_u32(smem + C::SM_INBUF) + C::OFF_ENC_HI, 128);
    // PAIR: barriers the peer arrives on are awaited with cluster-scope acquire; commits reach both CTAs
    auto wait_b = [&](uint64_t* bar, uint32_t par) {
      if constexpr (PAIR) mbar_wait_cluster(bar, par);
      else mbar_wait(bar, par);
    };
    auto commit_b = [&](uint64_t* bar) {
      if constexpr (PAIR) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    uint32_t tl = 0;
    if (PAIR && crank != 0) {
      // follower: the leader issues for the pair.  This warp only relays the arrival of this CTA's half of every chunk.
      const uint32_t lead_full = mapa_u32(smem_u32(&full[0]), 0);
      uint32_t rl = 0;
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
#pragma unroll 1
        for (int c = 0; c < 60; ++c) {
          mbar_wait(&full[c % kStages], ((c / kStages) + rl * (60 / kStages)) & 1);
          if (lane == 0) mbar_arrive_cluster(lead_full + (c % kStages) * 8);
        }
        ++rl;
      }
    } else
    for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++tl) {
      wait_b(&enc_ready[0], tl & 1);
      int cbase = 0;   // chunks of the tile consumed so far (compile-time after unrolling)
#pragma unroll
      for (int step = 0; step < 8; ++step) {
        const int nk = step_k_chunks(step), k_early = step_k_early(step);
        const uint32_t par = step & 1;   // 8 steps per tile: the n-th wait of a per-step barrier has parity step & 1
        if (lane == 0) NSR_TR(tl, step, 0);
        wait_b(&a_ready[0], par);        // A[K 0..127] of this step written, ACC0 drained
        if (lane == 0) NSR_TR(tl, step, 1);
        tc_fence_after_sync();
#pragma unroll
        for (int slot = 0; slot < 2 * nk; ++slot) {
          int nh, kc;
          issue_slot(nk, k_early, 2, slot, nh, kc);
          const int c = cbase + slot, stage = c % kStages;
          if (slot == k_early) {         // first chunk of half 1: ACC1 must be back (in the epilogue's registers)
            if (lane == 0) NSR_TR(tl, step, 5);
            wait_b(acc_free1, par);
            if (lane == 0) NSR_TR(tl, step, 6);
            tc_fence_after_sync();
          }
          if (slot == 2 * k_early) {     // first late-K chunk: A[K 128..255] written
            if (lane == 0) NSR_TR(tl, step, 2);
            wait_b(&a_ready[1], par);
            if (lane == 0) NSR_TR(tl, step, 3);
            tc_fence_after_sync();
          }
          // 60 chunks per tile through 10 (20) stages: 6 (3) uses of every stage per tile -- the parity pattern repeats (alternates)
          // from one tile to the next
          wait_b(&full[stage], ((c / kStages) + tl * (60 / kStages)) & 1);
          if (leader && !NSR_EXP(5)) {
            const uint32_t acc = nh ? TM_ACC1 : TM_ACC0;
            const uint32_t b = ring_lo + stage * (kStageBytes >> 4);
            if ((step == 0 || step == 5) && kc == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if constexpr (PAIR) umma_ss2_pair(acc, enc_hi + j * 16, HI_B, b + j * 16, HI_B, idesc, j ? 1u : 0u);
                else umma_ss2(acc, enc_hi + j * 16, HI_B, b + j * 16, HI_B, idesc, j ? 1u : 0u);
              }
            } else {
              const uint32_t at = TM_AHI + (step == 5 ? kc - 1 : kc) * 32;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if constexpr (PAIR) umma_ts2_pair(acc, at + j * 8, b + j * 16, HI_B, idesc, (j || kc) ? 1u : 0u);
                else umma_ts2(acc, at + j * 8, b + j * 16, HI_B, idesc, (j || kc) ? 1u : 0u);
              }
            }
          }
          if (leader) commit_b(&empty[stage]);
          if (slot == last_slot_half0(nk, k_early)) {
            if (lane == 0) NSR_TR(tl, step, 7);
            if (leader) commit_b(&acc_ready[0]);
          }
          if (slot == 2 * nk - 1 && leader) commit_b(&acc_ready[1]);
        }
        if (nk == k_early) wait_b(&a_ready[1], par);   // step 0 has no late-K chunk: consume the phase all the same
        if (lane == 0) NSR_TR(tl, step, 4);
        if (step == 5 && leader) commit_b(&enc_free[0]);  // last reader of the xyz encoding
        cbase += 2 * nk;
      }
    }
  } else if (warp == R::MMA) {
    // ===================================================================== MMA issuer
    // The whole warp runs the (uniform) control flow and waits; one elected lane issues the tcgen05
    // instructions, so descriptors and addresses live in uniform registers.
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, 128), idesc8 = make_idesc_f8(128, 128);
    constexpr uint32_t HI_B = sdesc_hi(1024), HI_DIR = sdesc_hi(512), HI_8 = sdesc_hi(512);  // 8-row groups 1024 B apart (512 B for the K=32 tile)
    const uint32_t ring_lo = sdesc_lo(smem_u32(sRing), 128);
    const uint32_t inbuf = smem_u32(smem + C::SM_INBUF);
    const uint32_t enc_hi = sdesc_lo(inbuf + C::OFF_ENC_HI, 128), enc_lo = sdesc_lo(inbuf + C::OFF_ENC_LO, 128);
    const uint32_t dir_hi = sdesc_lo(inbuf + C::OFF_DIR_HI, 128), dir_lo = sdesc_lo(inbuf + C::OFF_DIR_LO, 128);
    uint32_t stage = 0, phase = 0, tl = 0;
    Waiter w_a[2], w_enc[2], w_f1;
    bool ready = false;  // result of an early try_wait on full[stage]
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      w_enc[0].wait(&enc_ready[0]);
      for (int step = 0; step < kSteps; ++step) {
        if (step == 9) w_enc[1].wait(&enc_ready[1]);
        const int nk = step_k_chunks(step);
        const int nhs = step_n_halves(step);
        bool waited1 = false, waitedf = !kAccFree;
        if (lane == 0) NSR_TR(tl, step, 0);
        w_a[0].wait(&a_ready[0]);  // A[K 0..127] of this step written, ACC0 drained
        if (lane == 0) NSR_TR(tl, step, 1);
        if (step == 9) {           // the single 128-wide half of step 9 accumulates in ACC1
          w_a[1].wait(&a_ready[1]);
          waited1 = true;
          if (kAccFree) w_f1.wait(acc_free1);
          waitedf = true;
        }
        tc_fence_after_sync();
        const int k_early = step_k_early(step), slot_h0 = nhs == 2 ? last_slot_half0(nk, k_early) : -1;
        for (int slot = 0; slot < nk * nhs; ++slot) {
          {
            int nh, kc;
            issue_slot(nk, k_early, nhs, slot, nh, kc);
            const uint32_t acc = (step == 9 || nh == 1) ? TM_ACC1 : TM_ACC0;
            // ---- A source of this K chunk: 0 = xyz encoding, 1 = TMEM activations, 2 = view-dir encoding
            int src = 1, ak = kc;
            if ((step == 0 || step == 5) && kc == 0) src = 0;
            else if (step == 9 && kc == 4) src = 2;
            else if (step == 5) ak = kc - 1;
            // kAccFree: the early-K chunks of half 1 only need ACC1 back (its previous contents are in the epilogue's registers),
            // not the second half of the activations -- they run under the previous step's second drain
            if (!waitedf && nh == 1) {
              if (lane == 0) NSR_TR(tl, step, 5);
              w_f1.wait(acc_free1);
              if (lane == 0) NSR_TR(tl, step, 6);
              tc_fence_after_sync();
              waitedf = true;
            }
            if (!waited1 && ((!kAccFree && nh == 1) || (src == 1 && ak >= 2))) {
              if (lane == 0) NSR_TR(tl, step, 2);
              w_a[1].wait(&a_ready[1]);  // A[K 128..255] written, ACC1 drained
              if (lane == 0) NSR_TR(tl, step, 3);
              tc_fence_after_sync();
              waited1 = true;
            }
            if (!ready) mbar_wait(&full[stage], phase);
            // descriptors: only the start-address field moves (16-byte units): +16 per k-step of 16 elements
            const uint32_t bh = ring_lo + stage * (C::STAGE_BYTES >> 4), bl = bh + (CHUNK_BYTES >> 4);
            const uint32_t acc0 = kc != 0;  // first MMA of a half overwrites the accumulator
            if (leader && !NSR_EXP(5)) {
              if (kMixed && src == 1 && step >= MIX_X3_STEPS) {
                // fp16 main term + the two residual products as e4m3 MMAs (K = 32 each) into the same accumulator
                const uint32_t ah = TM_AHI + ak * 32, a8l = TM_A8L + ak * 16, a8h = TM_A8H + ak * 16;
                const uint32_t b8h = bl, b8l = bl + (F8_TILE_BYTES >> 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) umma_ts2(acc, ah + j * 8, bh + j * 16, HI_B, idesc, j ? 1u : acc0);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  umma_ts2_f8(acc, a8l + j * 8, b8h + j * 16, HI_8, idesc8, 1u);
                  umma_ts2_f8(acc, a8h + j * 8, b8l + j * 16, HI_8, idesc8, 1u);
                }
              } else if (src == 1) {
                const uint32_t ah = TM_AHI + ak * 32, al = TM_ALO + ak * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  umma_ts2(acc, ah + j * 8, bh + j * 16, HI_B, idesc, j ? 1u : acc0);
                  if (kSplit) {
                    umma_ts2(acc, al + j * 8, bh + j * 16, HI_B, idesc, 1u);
                    umma_ts2(acc, ah + j * 8, bl + j * 16, HI_B, idesc, 1u);
                  }
                }
              } else if (src == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  umma_ss2(acc, enc_hi + j * 16, HI_B, bh + j * 16, HI_B, idesc, j ? 1u : acc0);
                  if (kSplit) {
                    umma_ss2(acc, enc_lo + j * 16, HI_B, bh + j * 16, HI_B, idesc, 1u);
                    umma_ss2(acc, enc_hi + j * 16, HI_B, bl + j * 16, HI_B, idesc, 1u);
                  }
                }
              } else {  // view-dir encoding: 27 channels -> two k-steps
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  umma_ss2(acc, dir_hi + j * 16, HI_DIR, bh + j * 16, HI_B, idesc, 1u);
                  if (kSplit) {
                    umma_ss2(acc, dir_lo + j * 16, HI_DIR, bh + j * 16, HI_B, idesc, 1u);
                    umma_ss2(acc, dir_hi + j * 16, HI_DIR, bl + j * 16, HI_B, idesc, 1u);
                  }
                }
              }
            }
            if (leader) umma_commit(&empty[stage]);
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1;
            }
            // probe the next stage AFTER queueing this chunk's MMAs: the probe's latency then overlaps their execution
            ready = mbar_try_wait(&full[stage], phase);
          }
          // accumulator 0 is complete after its late-K chunks (every MMA that read A[K 0..127] has been issued before them: the
          // commit covers those too); accumulator 1 (and the single half of step 9) with the step's last chunk
          if (slot == slot_h0 && lane == 0) NSR_TR(tl, step, 7);
          if (slot == slot_h0 && leader) umma_commit(&acc_ready[0]);
          if (slot == nk * nhs - 1 && leader) umma_commit(&acc_ready[1]);
        }
        if (!waited1) w_a[1].wait(&a_ready[1]);  // keep the parity in step (step 0: no late-K chunk)
        if (!waitedf) w_f1.wait(acc_free1);
        if (lane == 0) NSR_TR(tl, step, 4);
        if (step == 5 && leader) umma_commit(&enc_free[0]);  // last reader of the xyz encoding
      }
      if (!CLASSIFY && leader) umma_commit(&enc_free[1]);
    }
  } else if (warp >= R::ENC0) {
    // ===================================================================== encoders (2 warps, 2 rows per thread):
    // tile i+1's encodings while tile i is in the tensor pipe
    const int er = tid - R::ENC0 * 32;
    uint32_t tl = 0;
    Waiter w_free[2];
    uint8_t* inbuf = smem + C::SM_INBUF;
    for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++tl) {
      float x[2][3], vd[2][3];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t p = point_of(tile, er + rr * 64);
#pragma unroll
        for (int d = 0; d < 3; ++d) x[rr][d] = vd[rr][d] = 0.f;
        if (p >= 0 && (a.flags & NSR_FLAG_EMBEDDED_INPUT)) {
          // NeRF.forward(x) on pre-embedded input (RH:99-122): x[p] = [63 xyz channels | 27 view-dir channels], copied below
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            x[rr][d] = a.z_or_pts[p * 90 + d];
            vd[rr][d] = a.z_or_pts[p * 90 + 63 + d];
          }
        } else if (p >= 0) {
          const int64_t ray = p / a.S;
          const float* rp = a.rays + ray * 11;
          if (a.flags & NSR_FLAG_PTS_INPUT) {
#pragma unroll
            for (int d = 0; d < 3; ++d) x[rr][d] = a.z_or_pts[p * 3 + d];
          } else {
            const float z = a.z_or_pts[p];
#pragma unroll
            for (int d = 0; d < 3; ++d) x[rr][d] = __fadd_rn(rp[d], __fmul_rn(rp[3 + d], z));  // RN:463
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) vd[rr][d] = rp[8 + d];
        }
      }
      // ---- gamma(x) (RH:47-48): [x, sin(2^k x), cos(2^k x)]_k, 63 channels + one zero pad
      if (tl >= 1) w_free[0].wait(&enc_free[0]);  // step 5 of the previous tile has read the old encoding
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = er + rr * 64;
        float e[64];
        e[0] = x[rr][0];
        e[1] = x[rr][1];
        e[2] = x[rr][2];
        if constexpr (CLASSIFY) {
          // tier 1 only certifies (its operands are rounded to 11 bits anyway): one accurate sincosf per coordinate, the octaves by
          // angle doubling -- the error doubles per octave, 3e-5 at 2^9 x, a sixteenth of an fp16 ulp of the encoding
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn, cs;
            sincosf(x[rr][d], &sn, &cs);
#pragma unroll
            for (int k = 0; k < 10; ++k) {
              e[3 + 6 * k + d] = sn;
              e[3 + 6 * k + 3 + d] = cs;
              const float s2 = 2.f * sn * cs, c2 = fmaf(-2.f * sn, sn, 1.f);
              sn = s2;
              cs = c2;
            }
          }
        } else if (a.flags & NSR_FLAG_EMBEDDED_INPUT) {
          const int64_t p = point_of(tile, row);
#pragma unroll
          for (int c = 3; c < 63; ++c) e[c] = p >= 0 ? a.z_or_pts[p * 90 + c] : 0.f;
        } else {
#pragma unroll
          for (int k = 0; k < 10; ++k) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float sn, cs;
              if (NSR_EXP(4)) sn = cs = x[rr][d];
              else sincosf(x[rr][d] * float(1 << k), &sn, &cs);
              e[3 + 6 * k + d] = sn;
              e[3 + 6 * k + 3 + d] = cs;
            }
          }
        }
        e[63] = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2<kSplit>(e[8 * g + 2 * q], e[8 * g + 2 * q + 1], h[q], l[q]);
          st_a8(inbuf + C::OFF_ENC_HI, 1024, row, g, h[0], h[1], h[2], h[3]);
          if (kSplit) st_a8(inbuf + C::OFF_ENC_LO, 1024, row, g, l[0], l[1], l[2], l[3]);
          if (SAVE && a.dump != nullptr) {
            uint8_t* d = a.dump + dump_off_ex(dumpP) + dump_blocked_off(tile, row, 64, g);
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + dump_lo(dumpP)) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async_smem();
      if constexpr (PAIR) {                       // one arrival per warp on the LEADER's barrier (it issues for both CTAs)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_enc_ready0);
      } else {
        mbar_arrive(&enc_ready[0]);
      }
      if (CLASSIFY) continue;                     // tier 1 stops at the alpha head: no view-dir encoding
      if (tl >= 1) w_free[1].wait(&enc_free[1]);  // step 9 of the previous tile has read the old view-dir encoding
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = er + rr * 64;
        float v[32];
        v[0] = vd[rr][0];
        v[1] = vd[rr][1];
        v[2] = vd[rr][2];
        if (a.flags & NSR_FLAG_EMBEDDED_INPUT) {
          const int64_t p = point_of(tile, row);
#pragma unroll
          for (int c = 3; c < 27; ++c) v[c] = p >= 0 ? a.z_or_pts[p * 90 + 63 + c] : 0.f;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float sn, cs;
              sincosf(vd[rr][d] * float(1 << k), &sn, &cs);
              v[3 + 6 * k + d] = sn;
              v[3 + 6 * k + 3 + d] = cs;
            }
          }
        }
#pragma unroll
        for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2<kSplit>(v[8 * g + 2 * q], v[8 * g + 2 * q + 1], h[q], l[q]);
          st_a8(inbuf + C::OFF_DIR_HI, 512, row, g, h[0], h[1], h[2], h[3]);
          if (kSplit) st_a8(inbuf + C::OFF_DIR_LO, 512, row, g, l[0], l[1], l[2], l[3]);
          if (SAVE && a.dump != nullptr) {
            uint8_t* d = a.dump + dump_off_ev(dumpP) + dump_blocked_off(tile, row, 32, g);
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + dump_lo(dumpP)) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&enc_ready[1]);
    }
  } else if constexpr (CLASSIFY) {
    // ===================================================================== tier-1 epilogue: 16 warps, 32 columns per thread
    // thread = (row, 32-column quarter cq): all 512 threads turn ACC0 into A[K 0..127], then ACC1 into A[K 128..255].  The step is a
    // dependency chain (accumulator complete -> converted -> next step's MMAs), so what counts is the latency of one conversion:
    // a quarter of a half per thread, four warps per scheduler to hide the TMEM and shared-memory latencies behind.
    const int quad = warp & 3, cq = warp >> 2;
    const int row = quad * 32 + lane;
    const int col0 = cq * 32;
    const uint32_t tlane = uint32_t(quad * 32) << 16;
    Waiter w_acc[2];
    // arrivals towards the MMA warp: every thread on the CTA's own barrier, or (PAIR) one lane per warp on the leader's
    auto arrive_a = [&](int half) {
      if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_a_ready[half]);
      } else {
        mbar_arrive(&a_ready[half]);
      }
    };
    auto arrive_f1 = [&]() {
      if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_acc_free1);
      } else {
        mbar_arrive(acc_free1);
      }
    };
    arrive_a(0);                                  // initial credits
    arrive_a(1);
    if (kAccFree) arrive_f1();
    const bool tracer = tid == 0;                 // debug timeline
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++tl) {
      const int64_t p = point_of(tile, row);
      float sigma = 0.f;
      for (int step = 0; step < 8; ++step) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t H[16];
          w_acc[half].wait(&acc_ready[half]);
          if (tracer) NSR_TR(tl, step, 8 + 4 * half);
          tc_fence_after_sync();
          {
            uint32_t u[32];
            tmem_ld32(tlane + (half ? TM_ACC1 : TM_ACC0) + col0, u);
            tmem_ld_wait();
            if (tracer) NSR_TR(tl, step, 9 + 4 * half);
            if (half == 1 && kAccFree) {   // ACC1 lives in registers now: the next step's (half 1, K early) chunks may overwrite it
              tc_fence_before_sync();
              arrive_f1();
            }
            if (step < 7) {
              // fp16(acc) + fp16(bias) with the ReLU fused, two columns per instruction: 2.5x fewer issue slots than the fp32 form,
              // and the conversion's latency is what the step waits for.  (One more rounding than the tier-2 arithmetic: tier 1
              // only certifies, and its distance to tier 2 is verified at run time.)
              const uint4* b16 = reinterpret_cast<const uint4*>(smem + C::SM_INBUF + C::OFF_DIR_HI + (step * 256 + 128 * half + col0) * 2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 bb = b16[j];
                H[4 * j] = hfma2_relu_one(pack_f16x2(__uint_as_float(u[8 * j]), __uint_as_float(u[8 * j + 1])), bb.x);
                H[4 * j + 1] = hfma2_relu_one(pack_f16x2(__uint_as_float(u[8 * j + 2]), __uint_as_float(u[8 * j + 3])), bb.y);
                H[4 * j + 2] = hfma2_relu_one(pack_f16x2(__uint_as_float(u[8 * j + 4]), __uint_as_float(u[8 * j + 5])), bb.z);
                H[4 * j + 3] = hfma2_relu_one(pack_f16x2(__uint_as_float(u[8 * j + 6]), __uint_as_float(u[8 * j + 7])), bb.w);
              }
            } else {
              // step 7 feeds the alpha head only (RH:109): fp32 throughout, nothing is stored
              const float* bias = sTail + TAIL_BIAS + 7 * 256 + 128 * half + col0;
              const float* walpha = sTail + TAIL_WALPHA + 128 * half + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * j), wa = *reinterpret_cast<const float4*>(walpha + 4 * j);
                sigma = fmaf(fmaxf(__uint_as_float(u[4 * j]) + bb.x, 0.f), wa.x, sigma);
                sigma = fmaf(fmaxf(__uint_as_float(u[4 * j + 1]) + bb.y, 0.f), wa.y, sigma);
                sigma = fmaf(fmaxf(__uint_as_float(u[4 * j + 2]) + bb.z, 0.f), wa.z, sigma);
                sigma = fmaf(fmaxf(__uint_as_float(u[4 * j + 3]) + bb.w, 0.f), wa.w, sigma);
              }
            }
          }
          if (tracer) NSR_TR(tl, step, 10 + 4 * half);
          if (step < 7 && !NSR_EXP(2)) {              // (nothing consumes the activations of step 7: tier 1 ends at the alpha head)
            tmem_st16(tlane + TM_AHI + 64 * half + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
            tmem_st_wait();
          }
          tc_fence_before_sync();
          arrive_a(half);
          if (tracer) NSR_TR(tl, step, 11 + 4 * half);
        }
      }
      // ---- sigma~ = alpha head of the fp32 post-ReLU activations (RH:109), four partial sums per row; certainly-empty points stop
      // here, the others join the active list (order within the list is irrelevant: every row of a tier-2 tile is independent).
      // The partials are safe in shared memory until the next tile's step 7: that needs a_ready arrivals of every thread of this tile.
      if (cq != 0) reinterpret_cast<float*>(sXch)[row * 4 + cq] = sigma;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (cq == 0) {   // warps 0..3, warp-uniform
        const float4 o = sXch[row];
        const float sg = ((sigma + o.y) + (o.z + o.w)) + sTail[TAIL_MISC];
        const bool act = p >= 0 && !(sg <= -a.tau);       // NaN counts as active
        if (p >= 0) reinterpret_cast<float4*>(a.raw)[p] = make_float4(0.f, 0.f, 0.f, sg);
        const uint32_t m = __ballot_sync(0xffffffffu, act);
        if (m != 0u) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(a.ctrl + AS_COUNT, uint32_t(__popc(m)));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (act) a.list[base + __popc(m & ((1u << lane) - 1u))] = int32_t(p);
        }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    // thread = (row, column half): row = 32 (warp & 3) + lane == TMEM lane; columns [col0, col0 + 64) of ACC0 / ACC1
    const int row = (warp & 3) * 32 + lane;
    const int ch = warp >> 2;
    const int col0 = ch * 64;
    const uint32_t tlane = uint32_t((warp & 3) * 32) << 16;
    Waiter w_acc[2];
    // initial credits: nothing to wait for before the very first step
    mbar_arrive(&a_ready[0]);
    mbar_arrive(&a_ready[1]);
    if (kAccFree) mbar_arrive(acc_free1);
    uint32_t tl = 0;
    uint32_t vbits = 0u;   // tier 2: max |sigma~ - sigma| of this thread's points, as float bits
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int64_t p = point_of(tile, row);
      float sigma = 0.f;
      uint32_t* mrow = (SAVE && a.relu_mask != nullptr) ? a.relu_mask + size_t(tile) * MASK_TILE_WORDS + (ch * 2) * 128 + row : nullptr;
      for (int step = 0; step < 9; ++step) {
        const float* bias = sTail + TAIL_BIAS + step * 256 + col0;
        const float* walpha = (step == 7) ? sTail + TAIL_WALPHA + col0 : nullptr;
        const bool relu = step != 8;  // feature_linear has no activation (RH:110)
        const float sc = kMixed ? sTail[TAIL_MIXSCALE + step] : 1.f;
        const bool f8next = kMixed && step + 1 >= MIX_X3_STEPS;  // operand format the NEXT step consumes
        // H: fp16 hi words.  L: fp16 residual words, or (f8next) [16 words e4m3 residuals | 16 words e4m3 copies]
        uint32_t H[32], L[kSplit ? 32 : 1];
        // bias (+ReLU, + alpha head) and operand split of this thread's 64 columns of accumulator half `half`
        auto convert = [&](const uint32_t(&u0)[32], const uint32_t(&u1)[32], int half) {
          if (NSR_EXP(3)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) H[j] = u0[j] ^ u1[j];
            return;
          }
          const float* bh = bias + 128 * half;
          const float* wa = walpha ? walpha + 128 * half : nullptr;
          if constexpr (kMixed) {
            if (f8next) {
              epi32_mix<true>(u0, bh, sc, relu, wa, sigma, H, nullptr, L, L + 16);
              epi32_mix<true>(u1, bh + 32, sc, relu, wa ? wa + 32 : nullptr, sigma, H + 16, nullptr, L + 8, L + 24);
            } else {
              epi32_mix<false>(u0, bh, sc, relu, wa, sigma, H, L, nullptr, nullptr);
              epi32_mix<false>(u1, bh + 32, sc, relu, wa ? wa + 32 : nullptr, sigma, H + 16, L + 16, nullptr, nullptr);
            }
          } else {
            epi32<kSplit>(u0, bh, relu, wa, sigma, H, L);
            epi32<kSplit>(u1, bh + 32, relu, wa ? wa + 32 : nullptr, sigma, H + 16, L + (kSplit ? 16 : 0));
          }
        };
        // accumulator column c is K index c of the next step: fp16 pairs -> packed column c / 2, e4m3 quads -> c / 4
        auto store = [&](int half) {
          if (NSR_EXP(2)) {
            tc_fence_before_sync();
            return;
          }
          tmem_st16(tlane + TM_AHI + 64 * half + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
          tmem_st16(tlane + TM_AHI + 64 * half + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&H[16]));
          if (kMixed && f8next) {
            tmem_st16(tlane + TM_A8L + 32 * half + col0 / 4, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
            tmem_st16(tlane + TM_A8H + 32 * half + col0 / 4, *reinterpret_cast<const uint32_t(*)[16]>(&L[kSplit ? 16 : 0]));
          } else if (kSplit) {
            tmem_st16(tlane + TM_ALO + 64 * half + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
            tmem_st16(tlane + TM_ALO + 64 * half + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&L[kSplit ? 16 : 0]));
          }
          tmem_st_wait();
          tc_fence_before_sync();
        };
        // ---- first half: drain ACC0 into registers while the second half is still in the tensor pipe
        w_acc[0].wait(&acc_ready[0]);
        if (tid == 0) NSR_TR(tl, step, 8);
        tc_fence_after_sync();
        {
          uint32_t u0[32], u1[32];
          tmem_ld32(tlane + TM_ACC0 + col0, u0);
          tmem_ld32(tlane + TM_ACC0 + col0 + 32, u1);
          tmem_ld_wait();
          convert(u0, u1, 0);
        }
        if (SAVE && mrow != nullptr && step < 8) {
          mrow[(step * 8 + 0) * 128] = sign_bits(H);
          mrow[(step * 8 + 1) * 128] = sign_bits(H + 16);
        }
        if (SAVE && a.dump != nullptr) dump64_hl(a.dump + dump_off_h(dumpP, step), dump_lo(dumpP), tile, row, 256, col0, H, L);   // step 8: F
        // ---- the chunks that read A[K 0..127] were issued before accumulator 0's last ones (common.cuh issue_slot), so they
        // retired with it: the first half of the activations may be overwritten now, while half 1 is still in the tensor pipe
        if (tid == 0) NSR_TR(tl, step, 9);
        store(0);
        mbar_arrive(&a_ready[0]);
        if (tid == 0) NSR_TR(tl, step, 11);
        // ---- every MMA of this step has retired
        w_acc[1].wait(&acc_ready[1]);
        if (tid == 0) NSR_TR(tl, step, 10);
        tc_fence_after_sync();
        // ---- second half: drain ACC1 straight into the operands for K 128..255
        {
          uint32_t u0[32], u1[32];
          tmem_ld32(tlane + TM_ACC1 + col0, u0);
          tmem_ld32(tlane + TM_ACC1 + col0 + 32, u1);
          tmem_ld_wait();
          if (kAccFree) {   // ACC1 lives in registers now: the next step's (half 1, K early) chunks may overwrite it
            tc_fence_before_sync();
            mbar_arrive(acc_free1);
          }
          convert(u0, u1, 1);
        }
        if (SAVE && mrow != nullptr && step < 8) {
          mrow[(step * 8 + 4) * 128] = sign_bits(H);
          mrow[(step * 8 + 5) * 128] = sign_bits(H + 16);
        }
        if (SAVE && a.dump != nullptr) dump64_hl(a.dump + dump_off_h(dumpP, step), dump_lo(dumpP), tile, row, 256, 128 + col0, H, L);
        store(1);
        mbar_arrive(&a_ready[1]);
        if (tid == 0) NSR_TR(tl, step, 12);
      }
      // ---- step 9: views layer accumulates in ACC1; rgb head on CUDA cores (RH:113-117)
      w_acc[1].wait(&acc_ready[1]);
      tc_fence_after_sync();
      // ACC0 / A[K 0..127] are not touched by this step's epilogue: release them now.  (Not before the wait above:
      // the MMA warp must have consumed the previous a_ready[0] phase first -- it has once step 9 was issued.)
      mbar_arrive(&a_ready[0]);
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      uint32_t mv0 = 0u, mv1 = 0u;   // sign bits of this thread's 2 x 32 views-layer columns
      uint32_t HV[32], HVL[32];      // the same columns as packed fp16 hi / residual words (only kept for the dump)
      {
        const float* bias = sTail + TAIL_BIAS + 9 * 256 + col0;
        const float4* wr = reinterpret_cast<const float4*>(sTail + TAIL_WRGB) + col0;
        const float sc9 = kMixed ? sTail[TAIL_MIXSCALE + 9] : 1.f;
        uint32_t u0[32], u1[32];
        tmem_ld32(tlane + TM_ACC1 + col0, u0);
        tmem_ld32(tlane + TM_ACC1 + col0 + 32, u1);
        tmem_ld_wait();
        // ACC1 now lives in registers: hand it back before the head arithmetic, so the next tile's step 0 is not held up
        tc_fence_before_sync();
        mbar_arrive(&a_ready[1]);
        if (kAccFree) mbar_arrive(acc_free1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b0 = *reinterpret_cast<const float4*>(bias + 4 * j), b1 = *reinterpret_cast<const float4*>(bias + 32 + 4 * j);
          const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
          float hq0[4] = {0.f, 0.f, 0.f, 0.f}, hq1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float h0 = fmaxf(kMixed ? fmaf(__uint_as_float(u0[4 * j + q]), sc9, bb0[q]) : __uint_as_float(u0[4 * j + q]) + bb0[q], 0.f);
            const float h1 = fmaxf(kMixed ? fmaf(__uint_as_float(u1[4 * j + q]), sc9, bb1[q]) : __uint_as_float(u1[4 * j + q]) + bb1[q], 0.f);
            const float4 w0 = wr[4 * j + q], w1 = wr[32 + 4 * j + q];
            if (SAVE) {
              hq0[q] = h0;
              hq1[q] = h1;
              const int cc = 4 * j + q;                       // column inside the 32-column word
              const int bit = (cc & 1) ? 16 + (cc >> 1) : (cc >> 1);
              mv0 |= (h0 > 0.f ? 1u : 0u) << bit;
              mv1 |= (h1 > 0.f ? 1u : 0u) << bit;
            }
            r0 = fmaf(h0, w0.x, r0);
            r1 = fmaf(h0, w0.y, r1);
            r2 = fmaf(h0, w0.z, r2);
            r0 = fmaf(h1, w1.x, r0);
            r1 = fmaf(h1, w1.y, r1);
            r2 = fmaf(h1, w1.z, r2);
          }
          if (SAVE) {
            split2<true>(hq0[0], hq0[1], HV[2 * j], HVL[2 * j]);
            split2<true>(hq0[2], hq0[3], HV[2 * j + 1], HVL[2 * j + 1]);
            split2<true>(hq1[0], hq1[1], HV[16 + 2 * j], HVL[16 + 2 * j]);
            split2<true>(hq1[2], hq1[3], HV[16 + 2 * j + 1], HVL[16 + 2 * j + 1]);
          }
        }
      }
      if (SAVE && a.dump != nullptr) dump64_hl(a.dump + dump_off_hv(dumpP), dump_lo(dumpP), tile, row, 128, col0, HV, HVL);
      if (SAVE && mrow != nullptr) {
        mrow[(64 + 0) * 128] = mv0;
        mrow[(64 + 1) * 128] = mv1;
      }
      // ---- the two column halves of a row meet in shared memory; the ch == 0 thread writes raw[p]
      if (ch == 1) sXch[row] = make_float4(r0, r1, r2, sigma);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0 && p >= 0) {
        const float4 o = sXch[row];
        const float* misc = sTail + TAIL_MISC;
        const float sg = (sigma + o.w) + misc[0];
        if (use_list) {   // tier 2 verification: how far was tier 1's sigma~ (still in raw[p]) from this one?
          const uint32_t d = __float_as_uint(fabsf(sg - a.raw[p * 4 + 3]));
          vbits = d > vbits ? d : vbits;
        }
        reinterpret_cast<float4*>(a.raw)[p] = make_float4((r0 + o.x) + misc[1], (r1 + o.y) + misc[2], (r2 + o.z) + misc[3], sg);  // RH:118
      }
    }
    if (use_list && ch == 0) {
      const uint32_t v = __reduce_max_sync(0xffffffffu, vbits);
      if (lane == 0 && v != 0u) atomicMax(a.ctrl + AS_VMAX, v);
    }
  }
  if (a.ctrl != nullptr && a.role == AS_ROLE_REDO && blockIdx.x == 0 && tid == 0) a.ctrl[AS_DENSE_FINAL] = 1u;
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) {
    cluster_sync_all();   // the peer may still arrive on this CTA's barriers / read its shared memory until it is done too
    if (warp == R::MMA) tmem_dealloc_pair(0u, 512);
  } else {
    if (warp == R::MMA) tmem_dealloc(0u, 512);
  }
}

static int& tier1_pair_flag() {
  static int flag = [] {
    const char* e = getenv("NSR_TIER1_PAIR");
    return (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }();
  return flag;
}
int set_tier1_pair(int enabled) {
  const int old = tier1_pair_flag();
  tier1_pair_flag() = enabled ? 1 : 0;
  return old;
}

template <int SPLIT, bool SAVE, bool CLASSIFY = false>
static int launch_variant(const MlpArgs& a, int grid, cudaStream_t st) {
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&nerf_mlp_kernel<SPLIT, SAVE, CLASSIFY>), Cfg<SPLIT>::SM_TOTAL)) return rc;
  nerf_mlp_kernel<SPLIT, SAVE, CLASSIFY><<<grid, Roles<CLASSIFY>::THREADS, Cfg<SPLIT>::SM_TOTAL, st>>>(a);
  count_launch();
  return check_launch("nerf_mlp_kernel");
}

// tier 1 as clusters of two CTAs (tcgen05 ... cta_group::2): an even grid, launched with the cluster attribute
static int launch_tier1_pair(const MlpArgs& a, int grid, cudaStream_t st) {
  auto kernel = &nerf_mlp_kernel<1, false, true, true>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), Cfg<1>::SM_TOTAL)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(grid), 1, 1);
  cfg.blockDim = dim3(Roles<true>::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg<1>::SM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kernel, a) != cudaSuccess) return check_launch("nerf_mlp_kernel (CTA pairs)");
  count_launch();
  return check_launch("nerf_mlp_kernel (CTA pairs)");
}

static int launch_by_flags(const MlpArgs& a, int grid, uint32_t flags, cudaStream_t st) {
  if (a.role == AS_ROLE_TIER1) {
    // one CTA per SM by default; as CTA pairs on request (nsr_set_tier1_pair / NSR_TIER1_PAIR=1: same results bit for bit, and on a
    // power-capped B200 the same speed -- DESIGN.md 3.1a).  The pair kernel wants an even grid of at least 2.
    const int tiles_even = (a.num_tiles + 1) & ~1;
    const int pgrid = (grid & ~1) < tiles_even ? (grid & ~1) : tiles_even;
    if (tier1_pair_flag() && pgrid >= 2 && a.experiment < 100) return launch_tier1_pair(a, pgrid, st);
    return launch_variant<1, false, true>(a, grid, st);
  }
  if (a.relu_mask != nullptr || a.dump != nullptr) {
    if (flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8)) {
      set_error("mlp_forward: sign bits / activations are saved for the backward pass, which is built for the default precision only");
      return NSR_E_INVALID;
    }
    return launch_variant<3, true>(a, grid, st);
  }
  if (flags & NSR_FLAG_FAST_FP16) return launch_variant<1, false>(a, grid, st);
  if (flags & NSR_FLAG_MIXED_F8) return launch_variant<2, false>(a, grid, st);
  return launch_variant<3, false>(a, grid, st);
}

int launch_mlp_forward(const float* rays, const float* z_or_pts, int64_t n, int S, const void* packed, uint32_t flags,
                       float* raw, cudaStream_t st, uint32_t* relu_mask, void* dump, int role, void* active_set) {
  const int64_t n_points = n * S;
  if (n_points == 0) return NSR_OK;
  if (n_points > (int64_t(1) << 31) * 64) {
    set_error("mlp_forward: too many points");
    return NSR_E_UNSUPPORTED;
  }
  int num_sms = 0;
  if (int rc = current_device_sms(&num_sms)) return rc;
  MlpArgs a;
  a.rays = rays;
  a.z_or_pts = z_or_pts;
  a.packed = static_cast<const uint8_t*>(packed);
  a.raw = raw;
  a.n_points = n_points;
  a.S = S;
  a.flags = flags;
  a.num_tiles = int((n_points + 127) / 128);
  a.relu_mask = relu_mask;
  a.dump = static_cast<uint8_t*>(dump);
  a.trace = nullptr;
  a.ctrl = nullptr;
  a.list = nullptr;
  a.role = AS_ROLE_PLAIN;
  a.tau = 0.f;
  a.verify_bits = 0u;
  if (role != AS_ROLE_PLAIN) {
    if (active_set == nullptr || n_points >= (int64_t(1) << 31)) {
      set_error("mlp_forward: two-tier launch without an active set (or more than 2^31 points)");
      return NSR_E_INVALID;
    }
    if ((flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8 | NSR_FLAG_PTS_INPUT | NSR_FLAG_EMBEDDED_INPUT)) || (dump != nullptr)) {
      set_error("mlp_forward: the two-tier evaluation runs the default precision on depth input, without the activation dump");
      return NSR_E_INVALID;
    }
    const TwoTierParams& tt = two_tier_params();
    a.ctrl = static_cast<uint32_t*>(active_set);
    a.list = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(active_set) + AS_CTRL_BYTES);
    a.role = role;
    a.tau = tt.tau;
    memcpy(&a.verify_bits, &tt.verify_max, 4);
  }
#ifdef NSR_DEBUG_HOOKS
  const char* experiment = getenv("NSR_EXPERIMENT");
  a.experiment = experiment != nullptr ? atoi(experiment) : 0;
#else
  a.experiment = 0;
#endif
  int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;
  if (a.experiment >= 100 && a.experiment - 100 < grid) grid = a.experiment - 100;
#ifdef NSR_DEBUG_HOOKS   // compiled out of the production library: `python -m neural_sim_nerf_b200.build --debug-hooks`
  const char* trace_file = getenv("NSR_TRACE_FILE");  // debug only: synchronous, dumps CTA 0's timeline
  if (trace_file != nullptr && a.num_tiles >= 4 * grid && (role == AS_ROLE_PLAIN || role == AS_ROLE_TIER1)) {
    const size_t nb = 4 * 10 * 16 * sizeof(unsigned long long);
    cudaMalloc(&a.trace, nb);
    cudaMemset(a.trace, 0, nb);
    const int rc = launch_by_flags(a, grid, flags, st);
    cudaStreamSynchronize(st);
    unsigned long long host[4 * 10 * 16];
    cudaMemcpy(host, a.trace, nb, cudaMemcpyDeviceToHost);
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_file, "a")) {
      fprintf(f, "# launch split=%d tiles=%d%s\n", (flags & NSR_FLAG_FAST_FP16) ? 1 : ((flags & NSR_FLAG_MIXED_F8) ? 2 : 3), a.num_tiles,
              role == AS_ROLE_TIER1 ? " tier1" : "");
      for (int i = 0; i < 40; ++i) {
        fprintf(f, "%d %d", i / 10, i % 10);
        for (int k = 0; k < 16; ++k) fprintf(f, " %llu", host[i * 16 + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    return rc;
  }
#endif
  return launch_by_flags(a, grid, flags, st);
}

}  // namespace nsr
