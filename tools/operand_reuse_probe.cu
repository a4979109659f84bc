// How does the tcgen05.mma issue rate depend on which operands consecutive instructions share?  One CTA per SM, one thread
// issues a long stream of M128 x N x K16 kind::f16 MMAs (no copies, no epilogue traffic); the stream's (A, B, accumulator)
// sequence is the experiment.  Motivation: the 3-MMAs-per-product stream of the fp16x3 kernels runs at 64 cycles per MMA, the
// 1-MMA-per-product stream of the tier-1 kernel at 98 (tools/mma_rate_probe1.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/operand_reuse_probe tools/operand_reuse_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"
using namespace nsr;

constexpr int B_BYTES = 160 * 1024;   // B operand pool: 40 K16 slices of [128 x 16] (4 KB each) or 20 of [256 x 16]
constexpr int A_BYTES = 32 * 1024;    // A operand pool for the SS form: 8 slices of [128 x 16]
constexpr int SMEM = B_BYTES + A_BYTES + 1024;

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
  return p != 0;
}

// pattern:
//  0 same A, same B            1 A varies (16 TMEM slices), same B        2 same A, B varies        3 both vary (tier-1 stream)
//  4 fp16x3 stream: (ah,bh) (al,bh) (ah,bl)                                5 pairs sharing A: (a,b0->acc0) (a,b1->acc1)
//  6 pairs sharing B: (a0,b) (a1,b) into two accumulators                  7 N = 256, both vary (A used once per 256 columns)
//  8 SS form, both vary        9 SS form, pairs sharing A                  10 triples sharing A  11 quads sharing A (4 accumulators of 64? no: N=128 x4 -> needs 512 cols: uses acc 0,128 alternately)
template <int PATTERN>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + B_BYTES + A_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < (B_BYTES + A_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003800u + (i * 2654435761u >> 28);  // ~1.0 / 0.5 with noise
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  {  // something finite in the A operand columns 256..511
    uint32_t h[16];
    for (int j = 0; j < 16; ++j) h[j] = 0x3c003800u + j;
    const uint32_t tl = uint32_t(warp * 32) << 16;
    for (int c = 256; c < 512; c += 16) tmem_st16(tl + c, h);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc128 = make_idesc_f16(128, 128), idesc256 = make_idesc_f16(128, 256);
    constexpr uint32_t HI_B = sdesc_hi(1024);
    const uint32_t b0 = sdesc_lo(smem_u32(smem), 128), a0 = sdesc_lo(smem_u32(smem + B_BYTES), 128);
    // slice i of the B pool: chunks of [128 x 64] = 16 KB hold 4 K16 slices 256 B apart (k-step = 16 elements = 2 core matrices)
    auto bslice = [&](int i) { return b0 + uint32_t(((i >> 2) * 16384 + (i & 3) * 256) >> 4); };
    auto aslice = [&](int i) { return a0 + uint32_t((((i >> 2) & 1) * 16384 + (i & 3) * 256) >> 4); };
    auto atm = [&](int i) { return 256u + uint32_t(i & 15) * 8u; };   // 16 K16 slices of a TMEM-resident [128 x 256] fp16 operand
    const long long t0 = clock64();
    if (leader) {
      // 240 MMAs per outer iteration, fully unrolled: every descriptor offset is a compile-time constant added to a uniform base
      // (a runtime-indexed stream is issue-bound at ~170 cycles per MMA: the operands have to travel through R2UR)
      for (int it = 0; it < n_mma / 240; ++it) {
#pragma unroll
        for (int i = 0; i < 240; ++i) {
          if constexpr (PATTERN == 0) umma_ts2(0, atm(0), bslice(0), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 1) umma_ts2(0, atm(i), bslice(0), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 2) umma_ts2(0, atm(0), bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 3) umma_ts2((i >> 4 & 1) * 128, atm(i), bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 4) {
            const int k = i / 3, r = i % 3;
            umma_ts2((k >> 4 & 1) * 128, r == 1 ? atm(k) + 128 : atm(k), bslice((2 * k + (r == 2)) % 40), HI_B, idesc128, 1u);
          }
          if constexpr (PATTERN == 5) umma_ts2((i & 1) * 128, atm(i >> 1), bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 6) umma_ts2((i & 1) * 128, atm(i), bslice((i >> 1) % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 7) umma_ts2(0, atm(i), b0 + uint32_t((((i >> 2) % 5) * 32768 + (i & 3) * 256) >> 4), HI_B, idesc256, 1u);
          if constexpr (PATTERN == 8) umma_ss2((i >> 4 & 1) * 128, aslice(i), HI_B, bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 9) umma_ss2((i & 1) * 128, aslice(i >> 1), HI_B, bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 10) umma_ts2((i % 3 == 1) * 128, atm(i / 3), bslice(i % 40), HI_B, idesc128, 1u);
          if constexpr (PATTERN == 11) umma_ts2((i & 1) * 128, atm(i >> 2), bslice(i % 40), HI_B, idesc128, 1u);
        }
      }
      umma_commit(done);
    }
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

template <int P>
static void launch(int grid, int n, long long* out) {
  cudaFuncSetAttribute(probe<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  probe<P><<<grid, 128, SMEM>>>(n, out);
}
static void run(int p, int grid, int n, long long* out) {
  switch (p) {
    case 0: launch<0>(grid, n, out); break;
    case 1: launch<1>(grid, n, out); break;
    case 2: launch<2>(grid, n, out); break;
    case 3: launch<3>(grid, n, out); break;
    case 4: launch<4>(grid, n, out); break;
    case 5: launch<5>(grid, n, out); break;
    case 6: launch<6>(grid, n, out); break;
    case 7: launch<7>(grid, n, out); break;
    case 8: launch<8>(grid, n, out); break;
    case 9: launch<9>(grid, n, out); break;
    case 10: launch<10>(grid, n, out); break;
    case 11: launch<11>(grid, n, out); break;
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);

  const char* names[] = {"same A, same B", "A varies, same B", "same A, B varies", "both vary (tier-1 stream)", "fp16x3 stream (ah,bh)(al,bh)(ah,bl)",
                         "pairs sharing A (two accumulators)", "pairs sharing B (two accumulators)", "N=256, both vary", "SS form, both vary",
                         "SS form, pairs sharing A", "triples sharing A", "quads sharing A"};
  const int n = 48000;
  for (int grid : {1, 148})
    for (int p = 0; p < 12; ++p) {
      run(p, grid, n, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("pattern %d: %s\n", p, cudaGetErrorString(e));
        return 1;
      }
      std::vector<long long> c(grid);
      cudaMemcpy(c.data(), out, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto v : c) mx = v > mx ? v : mx;
      printf("REUSE grid=%3d pattern %2d [%-38s]: %6.1f cycles per MMA%s\n", grid, p, names[p], double(mx) / n, p == 7 ? " (N=256: 2x the work)" : "");
    }
  return 0;
}
