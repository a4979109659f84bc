// Data-gradient of encode + NeRF MLP (the autograd tape of RN:26-40 / RH:99-122 seen from RN:177-178):
// given dL/draw [P,4] produce dL/dpoint [P,3] and dL/dviewdir [P,3].  Same tile / TMEM / warp-role design as
// mlp_forward.cu (read that header first); a tile runs 22 GEMM steps:
//   steps 0..9   forward recompute (nothing is stored by the forward pass: activations never leave TMEM);
//                every ReLU's sign is kept as one bit per activation in shared memory (32 KB per tile)
//   step 9'      its epilogue turns dL/drgb_raw into dL/dh_views = W_rgb^T g . [h_views > 0] on CUDA cores
//   steps 10..21 the transposed network (common.cuh "bstep" 0..11): dL/dz_l = (W_{l+1}^T dL/dz_{l+1}) . [h_l > 0],
//                two "side" GEMMs peel off dL/d(xyz encoding) (skip branch and layer 0), one dL/d(dir encoding)
// The encoding's own derivative, dgamma/dx = [1, 2^k cos(2^k x), -2^k sin(2^k x)], is applied on CUDA cores.
// Gradients are linear in dL/draw, so every row is scaled by a power of two to put its largest input at
// [1,2) before entering fp16 hi/lo operands and scaled back at the end.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mlp_common.cuh"

namespace nsr {

constexpr int BWD_STAGES = 4;
constexpr int B_ENC_BYTES = 128 * 64 * 2, B_DIR_BYTES = 128 * 32 * 2;
constexpr int B_OFF_ENC_HI = 0, B_OFF_DIR_HI = B_ENC_BYTES, B_OFF_ENC_LO = B_ENC_BYTES + B_DIR_BYTES, B_OFF_DIR_LO = 2 * B_ENC_BYTES + B_DIR_BYTES;
// MASKED = the forward pass saved the ReLU sign bits (common.cuh MASK_*): no forward recompute, so no encoding buffers; the
// space holds two mask tiles instead (the next tile's bits are bulk-copied in while this one computes).
template <bool MASKED>
struct BCfg {
  static constexpr int SM_INBUF = 0;
  static constexpr int SM_RING = MASKED ? 0 : 2 * (B_ENC_BYTES + B_DIR_BYTES);
  static constexpr int SM_TAIL = SM_RING + BWD_STAGES * CHUNK_PAIR_BYTES;
  static constexpr int SM_XCH = SM_TAIL + TAIL_BYTES;            // [128] x 2 float4
  static constexpr int SM_MASK = SM_XCH + 128 * 32;              // recompute: [8 layers][8 words][128 rows] u32; MASKED: 2 x [68][128]
  static constexpr int SM_BAR = SM_MASK + (MASKED ? 2 * MASK_TILE_BYTES : 8 * 8 * 128 * 4);
  static constexpr int SM_TOTAL = SM_BAR + 256;
  static_assert(SM_TOTAL <= 227 * 1024, "backward kernel shared memory");
  static_assert(SM_MASK % 16 == 0, "bulk-copy destination alignment");
};
constexpr int NUM_GSTEPS = NUM_STEPS + NUM_BSTEPS;               // 22

struct BwdArgs {
  const float* rays;
  const float* z;
  const uint8_t* packed;
  const float* d_raw;
  float* d_pts;
  int64_t n_points;
  int S;
  int num_tiles;
  const uint32_t* relu_mask;  // MASKED kernel: sign bits written by the forward pass, [tile][68][128] u32
  uint8_t* dump;        // optional: fp16 activations / pre-activation gradients of every layer (common.cuh), for dL/dMLP
  const float* gscale;  // with dump: one power-of-two scale for ALL rows (max |dL/draw| of the batch), device scalar
  unsigned long long* trace;  // debug (NSR_TRACE_FILE_BWD): clock64 stamps of CTA 0's first tiles, [tile][gstep][16]
  // active set of the forward pass (common.cuh): only its points carry gradient (every other point has dL/draw == 0 exactly, so its
  // dL/dpoint is 0 and the caller pre-zeroes d_pts); tile t of this launch = entries [128 t, 128 t + 128) of the list, which is also
  // how the forward pass indexed the sign bits it saved.  NULL: every point, in order.
  const uint32_t* ctrl;
  const int32_t* list;
};

#define NSR_TRB(tl, g, slot)                                                                              \
  do {                                                                                                  \
    if (a.trace != nullptr && blockIdx.x == 0 && (tl) < 3) a.trace[((tl) * 22 + (g)) * 16 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ bool gstep_is_side(int g) { return g == 9 || (g >= 10 && bstep_is_side(g - 10)); }
__device__ __forceinline__ int gstep_k_chunks(int g) { return g < 10 ? step_k_chunks(g) : bstep_k_chunks(g - 10); }
__device__ __forceinline__ int gstep_side_n(int g) { return g == 9 ? 128 : bstep_side_n(g - 10); }
__device__ __forceinline__ int gstep_k_early(int g) { return g < 10 ? step_k_early(g) : 2; }   // common.cuh issue_slot

// forward-recompute epilogue of 32 columns: bias + ReLU + hi/lo split (+ sign bits)
__device__ __forceinline__ uint32_t fwd32(const uint32_t (&u)[32], const float* bias, bool relu, uint32_t* H, uint32_t* L) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * j);
    float x0 = __uint_as_float(u[4 * j]) + bb.x, x1 = __uint_as_float(u[4 * j + 1]) + bb.y;
    float x2 = __uint_as_float(u[4 * j + 2]) + bb.z, x3 = __uint_as_float(u[4 * j + 3]) + bb.w;
    if (relu) {
      x0 = fmaxf(x0, 0.f);
      x1 = fmaxf(x1, 0.f);
      x2 = fmaxf(x2, 0.f);
      x3 = fmaxf(x3, 0.f);
    }
    split2<true>(x0, x1, H[2 * j], L[2 * j]);
    split2<true>(x2, x3, H[2 * j + 1], L[2 * j + 1]);
  }
  return sign_bits(H);
}

// backward epilogue of 32 columns: g = acc (+ extra[c] * gsig) masked by the forward ReLU sign bits -> hi/lo
__device__ __forceinline__ void bwd32(const uint32_t (&u)[32], uint32_t mask, bool use_mask, const float* extra, float gsig, uint32_t* H,
                                      uint32_t* L) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x0 = __uint_as_float(u[2 * j]), x1 = __uint_as_float(u[2 * j + 1]);
    if (extra != nullptr) {
      const float2 e = *reinterpret_cast<const float2*>(extra + 2 * j);
      x0 = fmaf(e.x, gsig, x0);
      x1 = fmaf(e.y, gsig, x1);
    }
    if (use_mask) {
      if (!(mask & (1u << j))) x0 = 0.f;
      if (!(mask & (0x10000u << j))) x1 = 0.f;
    }
    split2<true>(x0, x1, H[j], L[j]);
  }
}

// dL/dx += sum over the encoding channels [C0, C0+32) (below N_CH) of g[c] * d gamma_c / dx   (RH:47-48 backwards):
//   d sin(f x)/dx = f cos(f x),  d cos(f x)/dx = -f sin(f x),  f = 2^k.
// One accurate sincosf per coordinate at k = 0, then the octaves by angle doubling (sin 2a = 2 sin a cos a,
// cos 2a = 1 - 2 sin^2 a): the error doubles per octave exactly like the argument's own rounding error does
// (|x| <= 1.1, 2^9 x ~ 563 rad: ~3e-5 either way), and the Jacobian costs ~100 FLOPs instead of 32 sincosf calls.
template <int C0, int N_CH, int N_FREQ>
__device__ __forceinline__ void enc_backward32(const uint32_t (&u)[32], const float (&x)[3], float (&dx)[3]) {
  float sn[3], cs[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    sincosf(x[d], &sn[d], &cs[d]);
    if (d >= C0 && d < C0 + 32 && d < N_CH) dx[d] += __uint_as_float(u[d - C0]);   // the identity channels
  }
#pragma unroll
  for (int k = 0; k < N_FREQ; ++k) {
    const float f = float(1 << k);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int ci_s = 3 + 6 * k + d, ci_c = ci_s + 3;
      if (ci_s >= C0 && ci_s < C0 + 32 && ci_s < N_CH) dx[d] = fmaf(__uint_as_float(u[ci_s - C0]) * f, cs[d], dx[d]);
      if (ci_c >= C0 && ci_c < C0 + 32 && ci_c < N_CH) dx[d] = fmaf(-__uint_as_float(u[ci_c - C0]) * f, sn[d], dx[d]);
      const float s2 = 2.f * sn[d] * cs[d], c2 = fmaf(-2.f * sn[d], sn[d], 1.f);
      sn[d] = s2;
      cs[d] = c2;
    }
  }
}

template <bool MASKED>
__global__ void __launch_bounds__(MLP_THREADS, 1) nerf_mlp_bwd_kernel(BwdArgs a) {
  using C = BCfg<MASKED>;
  constexpr int G0 = MASKED ? 10 : 0;          // first GEMM step of a tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sRing = smem + C::SM_RING;
  const float* sTail = reinterpret_cast<const float*>(smem + C::SM_TAIL);
  float4* sXch = reinterpret_cast<float4*>(smem + C::SM_XCH);
  uint32_t* sMask = reinterpret_cast<uint32_t*>(smem + C::SM_MASK);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);
  uint64_t* empty = full + BWD_STAGES;
  uint64_t* acc_ready = empty + BWD_STAGES;
  uint64_t* a_ready = acc_ready + 2;
  uint64_t* enc_ready = a_ready + 2;
  uint64_t* enc_free = enc_ready + 2;
  uint64_t* mask_full = enc_free + 2;          // [2] MASKED: producer (tx bytes) -> epilogue
  uint64_t* mask_free = mask_full + 2;         // [2] MASKED: epilogue (256) -> producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_free + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t P = size_t(a.num_tiles) * 128;  // rows of the optional dump (dense launches only)
  int num_tiles = a.num_tiles;
  int n_act = 0;
  bool use_list = false;
  if (a.ctrl != nullptr && !(a.ctrl[AS_FORCE_DENSE] | a.ctrl[AS_DENSE_FINAL])) {
    use_list = true;
    n_act = int(a.ctrl[AS_COUNT]);
    num_tiles = (n_act + 127) >> 7;
    if (num_tiles == 0) return;
  }
  auto point_of = [&](int tile, int row) -> int64_t {
    const int64_t q = int64_t(tile) * 128 + row;
    if (use_list) return q < n_act ? int64_t(a.list[q]) : int64_t(-1);
    return q < a.n_points ? q : int64_t(-1);
  };

  if (tid == 0) {
    for (int s = 0; s < BWD_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&acc_ready[h], 1);
      mbar_init(&a_ready[h], EPI_THREADS);
      mbar_init(&enc_ready[h], ENC_THREADS);
      mbar_init(&enc_free[h], 1);
      mbar_init(&mask_full[h], 1);
      mbar_init(&mask_free[h], EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < TAIL_FLOATS; i += MLP_THREADS)
    reinterpret_cast<float*>(smem + C::SM_TAIL)[i] = reinterpret_cast<const float*>(a.packed + WEIGHT_BYTES)[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0u) __trap();
  if (warp == PROD_WARP) {
    // ===================================================================== weight producer: forward chunks then backward chunks
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tl = 0;
      bool first_lap = true;
      auto fetch_mask = [&](int tile, uint32_t t) {   // tile's sign bits -> buffer t & 1 (its t/2-th use)
        const uint32_t b = t & 1;
        if (t >= 2) mbar_wait(&mask_free[b], ((t >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&mask_full[b], MASK_TILE_BYTES);
        bulk_g2s(smem + C::SM_MASK + b * MASK_TILE_BYTES, reinterpret_cast<const uint8_t*>(a.relu_mask) + size_t(tile) * MASK_TILE_BYTES,
                 MASK_TILE_BYTES, &mask_full[b]);
      };
      if (MASKED && int(blockIdx.x) < num_tiles) fetch_mask(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
        int base = MASKED ? NUM_CHUNKS : 0;        // first packed chunk of the step
        for (int g = G0; g < NUM_GSTEPS; ++g) {
          const int nk = gstep_k_chunks(g), nhs = gstep_is_side(g) ? 1 : 2;
          for (int i = 0; i < nk * nhs; ++i) {       // in the order the MMA warp consumes them (common.cuh issue_slot)
            // the next tile's bits: a few chunks in, when the tile before this one has long released the other buffer
            if (MASKED && g == 12 && i == 0 && tile + int(gridDim.x) < num_tiles) fetch_mask(tile + gridDim.x, tl + 1);
            int nh, kc;
            issue_slot(nk, gstep_k_early(g), nhs, i, nh, kc);
            if (!first_lap) mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], CHUNK_PAIR_BYTES);
            bulk_g2s(sRing + stage * CHUNK_PAIR_BYTES, a.packed + size_t(base + nh * nk + kc) * CHUNK_PAIR_BYTES, CHUNK_PAIR_BYTES, &full[stage]);
            if (++stage == BWD_STAGES) {
              stage = 0;
              phase ^= 1;
              first_lap = false;
            }
          }
          base += nk * nhs;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================================================== MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t HI_B = sdesc_hi(1024), HI_DIR = sdesc_hi(512);
    const uint32_t ring_lo = sdesc_lo(smem_u32(sRing), 128);
    const uint32_t inbuf = smem_u32(smem + C::SM_INBUF);
    const uint32_t enc_hi = sdesc_lo(inbuf + B_OFF_ENC_HI, 128), enc_lo = sdesc_lo(inbuf + B_OFF_ENC_LO, 128);
    const uint32_t dir_hi = sdesc_lo(inbuf + B_OFF_DIR_HI, 128), dir_lo = sdesc_lo(inbuf + B_OFF_DIR_LO, 128);
    uint32_t stage = 0, phase = 0, tl = 0;
    Waiter w_a[2], w_enc[2];
    bool ready = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      if (!MASKED) w_enc[0].wait(&enc_ready[0]);
      for (int g = G0; g < NUM_GSTEPS; ++g) {
        if (g == 9) w_enc[1].wait(&enc_ready[1]);
        const bool side = gstep_is_side(g);
        const int nk = gstep_k_chunks(g);
        const int nhs = side ? 1 : 2;
        const uint32_t idesc = side ? make_idesc_f16(128, gstep_side_n(g)) : make_idesc_f16(128, 128);
        bool waited1 = false;
        if (lane == 0) NSR_TRB(tl, g, 0);
        w_a[0].wait(&a_ready[0]);
        if (lane == 0) NSR_TRB(tl, g, 1);
        if (side) {  // a side step accumulates in ACC1
          w_a[1].wait(&a_ready[1]);
          waited1 = true;
        }
        tc_fence_after_sync();
        const int k_early = gstep_k_early(g), slot_h0 = side ? -1 : last_slot_half0(nk, k_early);
        for (int slot = 0; slot < nk * nhs; ++slot) {
          {
            int nh, kc;
            issue_slot(nk, k_early, nhs, slot, nh, kc);
            const uint32_t acc = (side || nh == 1) ? TM_ACC1 : TM_ACC0;
            int src = 1, ak = kc;
            if ((g == 0 || g == 5) && kc == 0) src = 0;
            else if (g == 9 && kc == 4) src = 2;
            else if (g == 5) ak = kc - 1;
            if (!waited1 && (nh == 1 || (src == 1 && ak >= 2))) {
              if (lane == 0) NSR_TRB(tl, g, 2);
              w_a[1].wait(&a_ready[1]);
              if (lane == 0) NSR_TRB(tl, g, 3);
              tc_fence_after_sync();
              waited1 = true;
            }
            if (!ready) {
              if (lane == 0) NSR_TRB(tl, g, 5);   // weight ring not ready: last stamp = a stall on the TMA stream
              mbar_wait(&full[stage], phase);
              if (lane == 0) NSR_TRB(tl, g, 6);
            }
            const uint32_t bh = ring_lo + stage * (CHUNK_PAIR_BYTES >> 4), bl = bh + (CHUNK_BYTES >> 4);
            const uint32_t acc0 = kc != 0;
            if (leader) {
              if (src == 1) {
                const uint32_t ah = TM_AHI + ak * 32, al = TM_ALO + ak * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  umma_ts2(acc, ah + j * 8, bh + j * 16, HI_B, idesc, j ? 1u : acc0);
                  umma_ts2(acc, al + j * 8, bh + j * 16, HI_B, idesc, 1u);
                  umma_ts2(acc, ah + j * 8, bl + j * 16, HI_B, idesc, 1u);
                }
              } else if (src == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  umma_ss2(acc, enc_hi + j * 16, HI_B, bh + j * 16, HI_B, idesc, j ? 1u : acc0);
                  umma_ss2(acc, enc_lo + j * 16, HI_B, bh + j * 16, HI_B, idesc, 1u);
                  umma_ss2(acc, enc_hi + j * 16, HI_B, bl + j * 16, HI_B, idesc, 1u);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  umma_ss2(acc, dir_hi + j * 16, HI_DIR, bh + j * 16, HI_B, idesc, 1u);
                  umma_ss2(acc, dir_lo + j * 16, HI_DIR, bh + j * 16, HI_B, idesc, 1u);
                  umma_ss2(acc, dir_hi + j * 16, HI_DIR, bl + j * 16, HI_B, idesc, 1u);
                }
              }
              umma_commit(&empty[stage]);
            }
            if (++stage == BWD_STAGES) {
              stage = 0;
              phase ^= 1;
            }
            ready = mbar_try_wait(&full[stage], phase);
          }
          if (slot == slot_h0 && leader) umma_commit(&acc_ready[0]);
          if (slot == nk * nhs - 1 && leader) umma_commit(&acc_ready[1]);
        }
        if (!waited1) w_a[1].wait(&a_ready[1]);
        if (lane == 0) NSR_TRB(tl, g, 4);
        if (g == 5 && leader) umma_commit(&enc_free[0]);
        if (g == 9 && leader) umma_commit(&enc_free[1]);
      }
    }
  } else if (warp >= ENC_WARP0) {
    // ===================================================================== encoders (as in the forward kernel)
    const int er = tid - ENC_WARP0 * 32;
    uint32_t tl = 0;
    Waiter w_free[2];
    uint8_t* inbuf = smem + C::SM_INBUF;
    for (int tile = blockIdx.x; !MASKED && tile < num_tiles; tile += gridDim.x, ++tl) {
      float x[2][3], vd[2][3];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t p = point_of(tile, er + rr * 64);
#pragma unroll
        for (int d = 0; d < 3; ++d) x[rr][d] = vd[rr][d] = 0.f;
        if (p >= 0) {
          const float* rp = a.rays + (p / a.S) * 11;
          const float z = a.z[p];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            x[rr][d] = __fadd_rn(rp[d], __fmul_rn(rp[3 + d], z));
            vd[rr][d] = rp[8 + d];
          }
        }
      }
      if (tl >= 1) w_free[0].wait(&enc_free[0]);
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = er + rr * 64;
        float e[64];
        e[0] = x[rr][0];
        e[1] = x[rr][1];
        e[2] = x[rr][2];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn, cs;
            sincosf(x[rr][d] * float(1 << k), &sn, &cs);
            e[3 + 6 * k + d] = sn;
            e[3 + 6 * k + 3 + d] = cs;
          }
        }
        e[63] = 0.f;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2<true>(e[8 * gq + 2 * q], e[8 * gq + 2 * q + 1], h[q], l[q]);
          st_a8(inbuf + B_OFF_ENC_HI, 1024, row, gq, h[0], h[1], h[2], h[3]);
          st_a8(inbuf + B_OFF_ENC_LO, 1024, row, gq, l[0], l[1], l[2], l[3]);
          if (a.dump != nullptr) {
            uint8_t* d = a.dump + dump_off_ex(P) + dump_blocked_off(tile, row, 64, gq);
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + dump_lo(P)) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&enc_ready[0]);
      if (tl >= 1) w_free[1].wait(&enc_free[1]);
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int row = er + rr * 64;
        float v[32];
        v[0] = vd[rr][0];
        v[1] = vd[rr][1];
        v[2] = vd[rr][2];
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
#pragma unroll
        for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2<true>(v[8 * gq + 2 * q], v[8 * gq + 2 * q + 1], h[q], l[q]);
          st_a8(inbuf + B_OFF_DIR_HI, 512, row, gq, h[0], h[1], h[2], h[3]);
          st_a8(inbuf + B_OFF_DIR_LO, 512, row, gq, l[0], l[1], l[2], l[3]);
          if (a.dump != nullptr) {
            uint8_t* d = a.dump + dump_off_ev(P) + dump_blocked_off(tile, row, 32, gq);
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + dump_lo(P)) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&enc_ready[1]);
    }
  } else {
    // ===================================================================== epilogue warps
    const int row = (warp & 3) * 32 + lane;
    const int ch = warp >> 2;
    const int col0 = ch * 64;
    const uint32_t tlane = uint32_t((warp & 3) * 32) << 16;
    Waiter w_acc[2];
    if (!MASKED) {   // initial credits (MASKED: the first arrival is the dL/dh_views operand written at the top of every tile)
      mbar_arrive(&a_ready[0]);
      mbar_arrive(&a_ready[1]);
    }
    // sign-bit words of this thread: layer l, accumulator half h, 32-column group q  ->  sMask[(l*8 + h*4 + ch*2 + q)*128 + row]
    uint32_t* my_mask = sMask + (ch * 2) * 128 + row;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      if (MASKED) my_mask = sMask + (tl & 1) * MASK_TILE_WORDS + (ch * 2) * 128 + row;
      const int64_t p = point_of(tile, row);
      float x[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p >= 0) {
        const float* rp = a.rays + (p / a.S) * 11;
        const float z = a.z[p];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          x[d] = __fadd_rn(rp[d], __fmul_rn(rp[3 + d], z));
          vd[d] = rp[8 + d];
        }
        gr = reinterpret_cast<const float4*>(a.d_raw)[p];
      }
      // power-of-two row scale: largest |dL/draw| component -> [1, 2)
      const float gmax = fmaxf(fmaxf(fabsf(gr.x), fabsf(gr.y)), fmaxf(fabsf(gr.z), fabsf(gr.w)));
      // (in dump mode one scale serves the whole batch, so that dW = scale * G'^T H needs no per-row factor)
      float scale = __uint_as_float(__float_as_uint(a.gscale != nullptr ? *a.gscale : gmax) & 0x7f800000u);
      if (!(scale > 0.f) || !(scale < 3.0e38f)) scale = 1.f;
      const float inv = 1.f / scale;
      gr.x *= inv;
      gr.y *= inv;
      gr.z *= inv;
      gr.w *= inv;
      float dx[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};

      if (MASKED) {
        // ---------------------------------------------------- dL/dh_views = W_rgb^T dL/drgb_raw . [h_views > 0] (RH:117, RH:115 backwards)
        // from the saved sign bits: the operand of the first two backward steps, written without any GEMM.  Every MMA of the previous
        // tile retired before this thread left its last step (it waited on that step's accumulator), so A may be overwritten.
        mbar_wait(&mask_full[tl & 1], (tl >> 1) & 1);
        const float4* wr = reinterpret_cast<const float4*>(sTail + TAIL_WRGB) + col0;
        const uint32_t mv0 = my_mask[64 * 128], mv1 = my_mask[65 * 128];
        uint32_t H[32], L[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float gg[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = (q < 2) ? 2 * j + q : 32 + 2 * j + (q - 2);
            const uint32_t word = (q < 2) ? mv0 : mv1;
            const uint32_t bit = (q & 1) ? (0x10000u << j) : (1u << j);
            const float4 w = wr[c];
            gg[q] = (word & bit) ? (gr.x * w.x + gr.y * w.y + gr.z * w.z) : 0.f;
          }
          split2<true>(gg[0], gg[1], H[j], L[j]);
          split2<true>(gg[2], gg[3], H[16 + j], L[16 + j]);
        }
        if (a.dump != nullptr) dump64_hl(a.dump + dump_off_gv(P), dump_lo(P), tile, row, 128, col0, H, L);   // the activations were dumped by the forward pass
        tc_fence_after_sync();
        tmem_st16(tlane + TM_AHI + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
        tmem_st16(tlane + TM_AHI + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&H[16]));
        tmem_st16(tlane + TM_ALO + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
        tmem_st16(tlane + TM_ALO + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&L[16]));
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&a_ready[0]);
        mbar_arrive(&a_ready[1]);
      }

      for (int g = G0; g < NUM_GSTEPS; ++g) {
        if (!gstep_is_side(g)) {
          // ------------------------------------------------ full step: both accumulator halves -> next A operand
          const bool fwd = g < 10;
          const float* bias = sTail + TAIL_BIAS + g * 256 + col0;   // forward only
          const bool relu = g != 8;
          // backward: which forward layer's ReLU gates this gradient (none for dL/dfeature, g == 11)
          const int mlayer = fwd ? g : (g == 11 ? -1 : (g <= 14 ? 19 - g : 20 - g));  // g=12->h7 ... 14->h5, 16->h4 ... 20->h0
          const float* extra = (g == 12) ? sTail + TAIL_WALPHA + col0 : nullptr;      // alpha head: dL/dh7 += w_alpha * dL/dsigma
          uint32_t H[32], L[32];
          w_acc[0].wait(&acc_ready[0]);
          if (tid == 0) NSR_TRB(tl, g, 8);
          tc_fence_after_sync();
          {
            uint32_t u0[32], u1[32];
            tmem_ld32(tlane + TM_ACC0 + col0, u0);
            tmem_ld32(tlane + TM_ACC0 + col0 + 32, u1);
            tmem_ld_wait();
            if (fwd) {
              const uint32_t m0 = fwd32(u0, bias, relu, H, L);
              const uint32_t m1 = fwd32(u1, bias + 32, relu, H + 16, L + 16);
              if (g < 8) {
                my_mask[(g * 8 + 0) * 128] = m0;
                my_mask[(g * 8 + 1) * 128] = m1;
              }
            } else {
              const uint32_t m0 = mlayer >= 0 ? my_mask[(mlayer * 8 + 0) * 128] : 0u;
              const uint32_t m1 = mlayer >= 0 ? my_mask[(mlayer * 8 + 1) * 128] : 0u;
              bwd32(u0, m0, mlayer >= 0, extra, gr.w, H, L);
              bwd32(u1, m1, mlayer >= 0, extra ? extra + 32 : nullptr, gr.w, H + 16, L + 16);
            }
          }
          // forward step g writes H_g (g = 8: F); backward steps write GF (g = 11) or G_l of the layer whose ReLU gated them
          uint8_t* dump_arr = a.dump == nullptr ? nullptr
                              : a.dump + (fwd ? dump_off_h(P, g) : (g == 11 ? dump_off_gf(P) : dump_off_g(P, mlayer)));
          if (dump_arr != nullptr) dump64_hl(dump_arr, dump_lo(P), tile, row, 256, col0, H, L);
          if (tid == 0) NSR_TRB(tl, g, 9);
          // The chunks that read A[K 0..127] are issued before accumulator 0's last ones (common.cuh issue_slot) and retired with it,
          // so the first operand half can be overwritten while half 1 is still in the tensor pipe -- except in step 11, whose K is
          // 128 wide: both of its halves read A[K 0..127] and accumulator 1's chunks come after accumulator 0 is complete.
          const bool early_store = g != 11;
          if (!early_store) {
            w_acc[1].wait(&acc_ready[1]);
            tc_fence_after_sync();
          }
          tmem_st16(tlane + TM_AHI + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
          tmem_st16(tlane + TM_AHI + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&H[16]));
          tmem_st16(tlane + TM_ALO + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
          tmem_st16(tlane + TM_ALO + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&L[16]));
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&a_ready[0]);
          if (tid == 0) NSR_TRB(tl, g, 11);
          if (early_store) {
            w_acc[1].wait(&acc_ready[1]);
            tc_fence_after_sync();
          }
          if (tid == 0) NSR_TRB(tl, g, 10);
          {
            uint32_t u0[32], u1[32];
            tmem_ld32(tlane + TM_ACC1 + col0, u0);
            tmem_ld32(tlane + TM_ACC1 + col0 + 32, u1);
            tmem_ld_wait();
            if (fwd) {
              const uint32_t m0 = fwd32(u0, bias + 128, relu, H, L);
              const uint32_t m1 = fwd32(u1, bias + 160, relu, H + 16, L + 16);
              if (g < 8) {
                my_mask[(g * 8 + 4) * 128] = m0;
                my_mask[(g * 8 + 5) * 128] = m1;
              }
            } else {
              const uint32_t m0 = mlayer >= 0 ? my_mask[(mlayer * 8 + 4) * 128] : 0u;
              const uint32_t m1 = mlayer >= 0 ? my_mask[(mlayer * 8 + 5) * 128] : 0u;
              bwd32(u0, m0, mlayer >= 0, extra ? extra + 128 : nullptr, gr.w, H, L);
              bwd32(u1, m1, mlayer >= 0, extra ? extra + 160 : nullptr, gr.w, H + 16, L + 16);
            }
          }
          if (dump_arr != nullptr) dump64_hl(dump_arr, dump_lo(P), tile, row, 256, 128 + col0, H, L);
          tmem_st16(tlane + TM_AHI + 64 + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
          tmem_st16(tlane + TM_AHI + 64 + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&H[16]));
          tmem_st16(tlane + TM_ALO + 64 + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
          tmem_st16(tlane + TM_ALO + 64 + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&L[16]));
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&a_ready[1]);
          if (tid == 0) NSR_TRB(tl, g, 12);
        } else if (g == 9) {
          // ------------------------------------------------ forward views layer (ACC1, 128 wide) -> dL/dh_views into A[K 0..127]
          w_acc[1].wait(&acc_ready[1]);
          tc_fence_after_sync();
          const float* bias = sTail + TAIL_BIAS + 9 * 256 + col0;
          const float4* wr = reinterpret_cast<const float4*>(sTail + TAIL_WRGB) + col0;
          uint32_t H[32], L[32], HV[32], HVL[32];
          {
            uint32_t u0[32], u1[32];
            tmem_ld32(tlane + TM_ACC1 + col0, u0);
            tmem_ld32(tlane + TM_ACC1 + col0 + 32, u1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float gg[4], hh[4];
              const float2 bA = *reinterpret_cast<const float2*>(bias + 2 * j), bB = *reinterpret_cast<const float2*>(bias + 32 + 2 * j);
              const float bq[4] = {bA.x, bA.y, bB.x, bB.y};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int c = (q < 2) ? 2 * j + q : 32 + 2 * j + (q - 2);
                const float hv = __uint_as_float((q < 2) ? u0[2 * j + q] : u1[2 * j + (q - 2)]) + bq[q];
                const float4 w = wr[c];
                hh[q] = fmaxf(hv, 0.f);
                gg[q] = (hv > 0.f) ? (gr.x * w.x + gr.y * w.y + gr.z * w.z) : 0.f;   // RH:117 backwards, gated by RH:115
              }
              split2<true>(gg[0], gg[1], H[j], L[j]);
              split2<true>(gg[2], gg[3], H[16 + j], L[16 + j]);
              split2<true>(hh[0], hh[1], HV[j], HVL[j]);
              split2<true>(hh[2], hh[3], HV[16 + j], HVL[16 + j]);
            }
          }
          if (a.dump != nullptr) {
            dump64_hl(a.dump + dump_off_hv(P), dump_lo(P), tile, row, 128, col0, HV, HVL);
            dump64_hl(a.dump + dump_off_gv(P), dump_lo(P), tile, row, 128, col0, H, L);
          }
          tmem_st16(tlane + TM_AHI + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&H[0]));
          tmem_st16(tlane + TM_AHI + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&H[16]));
          tmem_st16(tlane + TM_ALO + col0 / 2, *reinterpret_cast<const uint32_t(*)[16]>(&L[0]));
          tmem_st16(tlane + TM_ALO + col0 / 2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&L[16]));
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&a_ready[0]);
          mbar_arrive(&a_ready[1]);
        } else {
          // ------------------------------------------------ backward side step (ACC1): gradient w.r.t. an encoding
          w_acc[1].wait(&acc_ready[1]);
          tc_fence_after_sync();
          // (MASKED: the tile's last step hands nothing on -- the next arrival is the next tile's dL/dh_views operand)
          const bool hand_on = !(MASKED && g == NUM_GSTEPS - 1);
          if (hand_on) mbar_arrive(&a_ready[0]);  // A is only read by this step
          // ACC1 is copied to registers and handed back at once: the sin/cos Jacobian below takes thousands of cycles and
          // must not hold up the next step's second half
          uint32_t u[32];
          const bool mine = (g != 10) || (ch == 0);   // dL/d(view-dir encoding), 27 channels: the ch == 0 threads take all 32 columns
          if (mine) {
            tmem_ld32(tlane + TM_ACC1 + (g == 10 ? 0 : ch * 32), u);
            tmem_ld_wait();
          }
          tc_fence_before_sync();
          if (hand_on) mbar_arrive(&a_ready[1]);
          if (mine) {
            if (g == 10) enc_backward32<0, 27, 4>(u, vd, dv);
            else if (ch == 0) enc_backward32<0, 63, 10>(u, x, dx);    // dL/d(xyz encoding), 63 channels: 32 per column half
            else enc_backward32<32, 63, 10>(u, x, dx);
          }
        }
      }
      if (MASKED) mbar_arrive(&mask_free[tl & 1]);   // this tile's sign bits are no longer needed
      // ---- the two column halves of a row meet in shared memory; the ch == 0 thread writes d_pts[p]
      if (ch == 1) {
        sXch[2 * row] = make_float4(dx[0], dx[1], dx[2], 0.f);
        sXch[2 * row + 1] = make_float4(dv[0], dv[1], dv[2], 0.f);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0 && p >= 0) {
        const float4 o0 = sXch[2 * row], o1 = sXch[2 * row + 1];
        float4* out = reinterpret_cast<float4*>(a.d_pts) + 2 * p;
        out[0] = make_float4((dx[0] + o0.x) * scale, (dx[1] + o0.y) * scale, (dx[2] + o0.z) * scale, 0.f);
        out[1] = make_float4((dv[0] + o1.x) * scale, (dv[1] + o1.y) * scale, (dv[2] + o1.z) * scale, 0.f);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(0u, 512);
}

size_t mlp_dump_bytes(int64_t n_points) { return dump_total(size_t((n_points + 127) / 128) * 128); }

int launch_mlp_backward(const float* rays, const float* z, int64_t n, int S, const void* packed, const float* d_raw, float* d_pts,
                        void* dump, const float* gscale, cudaStream_t st, const uint32_t* relu_mask, const void* active_set) {
  const int64_t n_points = n * S;
  if (n_points == 0) return NSR_OK;
  int num_sms = 0;
  if (int rc = current_device_sms(&num_sms)) return rc;
  // relu_mask with dump: the forward pass (launch_mlp_forward with relu_mask AND dump) already wrote the activation half of the
  // dump (EX, EV, H0..H7, F, HV); this pass adds the gradient half (GV, GF, G0..G7)
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&nerf_mlp_bwd_kernel<false>), BCfg<false>::SM_TOTAL)) return rc;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&nerf_mlp_bwd_kernel<true>), BCfg<true>::SM_TOTAL)) return rc;
  BwdArgs a;
  a.rays = rays;
  a.z = z;
  a.packed = static_cast<const uint8_t*>(packed);
  a.d_raw = d_raw;
  a.d_pts = d_pts;
  a.n_points = n_points;
  a.S = S;
  a.num_tiles = int((n_points + 127) / 128);
  a.relu_mask = relu_mask;
  a.dump = static_cast<uint8_t*>(dump);
  a.gscale = dump != nullptr ? gscale : nullptr;
  a.trace = nullptr;
  a.ctrl = nullptr;
  a.list = nullptr;
  if (active_set != nullptr) {
    if (dump != nullptr || n_points >= (int64_t(1) << 31)) {
      set_error("mlp_backward: the active-set route computes dL/d(rays) only (no parameter-gradient dump)");
      return NSR_E_INVALID;
    }
    a.ctrl = static_cast<const uint32_t*>(active_set);
    a.list = reinterpret_cast<const int32_t*>(static_cast<const uint8_t*>(active_set) + AS_CTRL_BYTES);
  }
#ifdef NSR_DEBUG_HOOKS
  const char* trace_file = getenv("NSR_TRACE_FILE_BWD");  // debug only: synchronous, dumps CTA 0's timeline
  if (trace_file != nullptr && a.num_tiles >= 3 * num_sms && relu_mask == nullptr) {
    const size_t nb = 3 * 22 * 16 * sizeof(unsigned long long);
    cudaMalloc(&a.trace, nb);
    cudaMemset(a.trace, 0, nb);
    nerf_mlp_bwd_kernel<false><<<num_sms, MLP_THREADS, BCfg<false>::SM_TOTAL, st>>>(a);
    cudaStreamSynchronize(st);
    unsigned long long host[3 * 22 * 16];
    cudaMemcpy(host, a.trace, nb, cudaMemcpyDeviceToHost);
    cudaFree(a.trace);
    if (FILE* f = fopen(trace_file, "a")) {
      fprintf(f, "# bwd launch tiles=%d\n", a.num_tiles);
      for (int i = 0; i < 66; ++i) {
        fprintf(f, "%d %d", i / 22, i % 22);
        for (int k = 0; k < 16; ++k) fprintf(f, " %llu", host[i * 16 + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    count_launch();
    return check_launch("nerf_mlp_bwd_kernel");
  }
#endif
  const int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;
  if (relu_mask != nullptr) nerf_mlp_bwd_kernel<true><<<grid, MLP_THREADS, BCfg<true>::SM_TOTAL, st>>>(a);
  else nerf_mlp_bwd_kernel<false><<<grid, MLP_THREADS, BCfg<false>::SM_TOTAL, st>>>(a);
  count_launch();
  return check_launch("nerf_mlp_bwd_kernel");
}

}  // namespace nsr
