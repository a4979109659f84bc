// Pieces shared by the forward (mlp_forward.cu) and backward (mlp_backward.cu) MLP kernels: warp roles,
// the TMEM column map, operand store / hi-lo split helpers, mbarrier bookkeeping.
#pragma once
#include "common.cuh"
#include "sm100_prims.cuh"

namespace nsr {

// ----------------------------------------------------------------------------- kernel configuration
// Warp roles (384 threads):
//   0-7   epilogue: warp w drains TMEM lane quadrant (w & 3), columns [64 (w >> 2), +64) of each accumulator half
//   8-9   encoders (two rows per thread)
//   10    MMA issuer (all lanes run the control flow, one elected lane issues)
//   11    weight producer
constexpr int MLP_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int ENC_THREADS = 64;
constexpr int ENC_WARP0 = 8, MMA_WARP = 10, PROD_WARP = 11;
// The kernel allocates all 512 TMEM columns of its SM (1 CTA / SM), so the allocation starts at column 0,
// lane 0; the addresses below are absolute.  (Checked at run time: the kernel traps otherwise.)
constexpr uint32_t TM_ACC0 = 0, TM_ACC1 = 128, TM_AHI = 256, TM_ALO = 384;
// mixed precision: the ALO columns hold two 8-bit copies of the 256 activations instead (four K-consecutive bytes per column)
constexpr uint32_t TM_A8L = 384;  // e4m3((x - fp16(x)) 2^11)
constexpr uint32_t TM_A8H = 448;  // e4m3(x)

// 16-byte store of 8 fp16 (4 packed words) into a no-swizzle K-major tile whose 8-row groups are `sbo` bytes apart
__device__ __forceinline__ void st_a8(uint8_t* tile, int sbo, int row, int kgroup, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  *reinterpret_cast<uint4*>(tile + (row >> 3) * sbo + kgroup * 128 + (row & 7) * 16) = make_uint4(w0, w1, w2, w3);
}

// (x0, x1) -> packed fp16 hi word and (SPLIT) the packed fp16 residual word
template <bool kSplit>
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2(x0, x1);
  if (kSplit) {
    const __half2 h = *reinterpret_cast<const __half2*>(&hi);
    const float2 f = __half22float2(h);
    lo = pack_f16x2(x0 - f.x, x1 - f.y);
  } else {
    lo = 0u;
  }
}

// sign bits of 32 freshly rounded activations (16 packed fp16 pairs, all >= 0): bit j <- low half of pair j,
// bit 16 + j <- high half
__device__ __forceinline__ uint32_t sign_bits(const uint32_t* H) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) m |= __vcmpne2(H[j], 0u) & (0x00010001u << j);
  return m;
}

// 64 consecutive features (32 packed fp16 words) of tile-row `row` at feature `col` of a blocked [P, W] dump array
__device__ __forceinline__ void dump64(uint8_t* arr, int tile, int row, int W, int col, const uint32_t* H) {
  uint8_t* dst = arr + dump_blocked_off(tile, row, W, col >> 3);
#pragma unroll
  for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(dst + q * 128) = make_uint4(H[4 * q], H[4 * q + 1], H[4 * q + 2], H[4 * q + 3]);
}

// the same for an operand pair: hi words into the array, residual words into its copy `lo_off` bytes further on
__device__ __forceinline__ void dump64_hl(uint8_t* arr, size_t lo_off, int tile, int row, int W, int col, const uint32_t* H, const uint32_t* L) {
  dump64(arr, tile, row, W, col, H);
  dump64(arr + lo_off, tile, row, W, col, L);
}

struct Waiter {  // one per (thread, barrier): parity follows the number of completed waits
  uint32_t n = 0;
  __device__ __forceinline__ void wait(uint64_t* bar) {
    mbar_wait(bar, n & 1);
    ++n;
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
  return p != 0;
}

// bias + (ReLU) + fp16 hi/lo split of 32 accumulator columns; optional fp32 dot with the alpha head
template <bool kSplit>
__device__ __forceinline__ void epi32(const uint32_t (&u)[32], const float* bias, bool relu, const float* walpha, float& sigma,
                                      uint32_t* H, uint32_t* L) {
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
    if (walpha != nullptr) {  // alpha head on the fp32 post-ReLU activations (RH:109)
      const float4 wa = *reinterpret_cast<const float4*>(walpha + 4 * j);
      sigma = fmaf(x0, wa.x, sigma);
      sigma = fmaf(x1, wa.y, sigma);
      sigma = fmaf(x2, wa.z, sigma);
      sigma = fmaf(x3, wa.w, sigma);
    }
    uint32_t l0, l1;
    split2<kSplit>(x0, x1, H[2 * j], l0);
    split2<kSplit>(x2, x3, H[2 * j + 1], l1);
    if (kSplit) {
      L[2 * j] = l0;
      L[2 * j + 1] = l1;
    }
  }
}

// Mixed-precision variant: x = acc * scale + bias.  F8NEXT = false: fp16 hi words H[16] + fp16 residual words L[16]
// (the next step runs the fp16 hi/lo split); true: H[16] + e4m3 residuals L8[8] + e4m3 copies H8[8] (four K-consecutive
// values per word) for the next step's kind::f8f6f4 residual products.
template <bool F8NEXT>
__device__ __forceinline__ void epi32_mix(const uint32_t (&u)[32], const float* bias, float scale, bool relu, const float* walpha,
                                          float& sigma, uint32_t* H, uint32_t* L, uint32_t* L8, uint32_t* H8) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * j);
    float x0 = fmaf(__uint_as_float(u[4 * j]), scale, bb.x), x1 = fmaf(__uint_as_float(u[4 * j + 1]), scale, bb.y);
    float x2 = fmaf(__uint_as_float(u[4 * j + 2]), scale, bb.z), x3 = fmaf(__uint_as_float(u[4 * j + 3]), scale, bb.w);
    if (relu) {
      x0 = fmaxf(x0, 0.f);
      x1 = fmaxf(x1, 0.f);
      x2 = fmaxf(x2, 0.f);
      x3 = fmaxf(x3, 0.f);
    }
    if (walpha != nullptr) {  // alpha head on the fp32 post-ReLU activations (RH:109)
      const float4 wa = *reinterpret_cast<const float4*>(walpha + 4 * j);
      sigma = fmaf(x0, wa.x, sigma);
      sigma = fmaf(x1, wa.y, sigma);
      sigma = fmaf(x2, wa.z, sigma);
      sigma = fmaf(x3, wa.w, sigma);
    }
    const uint32_t h01 = pack_f16x2(x0, x1), h23 = pack_f16x2(x2, x3);
    H[2 * j] = h01;
    H[2 * j + 1] = h23;
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
    if (F8NEXT) {
      constexpr float kS = float(1 << MIX_XLO_SHIFT);
      L8[j] = pack_e4m3x4((x0 - f01.x) * kS, (x1 - f01.y) * kS, (x2 - f23.x) * kS, (x3 - f23.y) * kS);
      H8[j] = pack_e4m3x4(x0, x1, x2, x3);
    } else {
      L[2 * j] = pack_f16x2(x0 - f01.x, x1 - f01.y);
      L[2 * j + 1] = pack_f16x2(x2 - f23.x, x3 - f23.y);
    }
  }
}

}  // namespace nsr
