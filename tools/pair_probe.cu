// CTA-pair MMA semantics probe (tcgen05 ... cta_group::2), the building block of the paired tier-1 kernel:
//   cluster of 2 CTAs; each CTA holds its own A [128 x 64] fp16 in TMEM (TS form) and HALF of B [128 x 64]: rows [64 r, 64 r + 64) of the
//   weight chunk, in the K-major no-swizzle layout of the packed network; the leader (rank 0) issues 4 MMAs M=256 N=128 K=16;
//   D [128 x 128] fp32 of each CTA's rows lands in that CTA's TMEM; tcgen05.commit multicast signals both CTAs.
// Checks D = A_cta . B^T exactly (small integers) in both CTAs, and the remote mbarrier arrive the follower uses to report to the leader.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/pair_probe tools/pair_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"
using namespace nsr;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ts2_pair(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` (same offset) in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(uint16_t(3))
               : "memory");
}

// mode 0: B rows [64 r, +64) per CTA (N split).  out[cta][128][128]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* out, int* flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                                       // [64 x 64] fp16, 8-row groups 1024 B apart, K core matrices 128 B apart
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 8192);
  uint64_t* ready = done + 1;                               // leader's: follower reports "my operands are in place"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(done, 1);
    mbar_init(ready, 2);                                    // leader itself + the follower's remote arrive
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc2(tmem_slot, 512);
  // B half of this CTA
  for (int i = tid; i < 64 * 64; i += 128) {
    const int row = i / 64, k = i % 64;
    const int off = (row >> 3) * 1024 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sB + off) = B[(rank * 64 + row) * 64 + k];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  // A of this CTA into TMEM columns [256, 288): lane = row, column j = (k = 2j, 2j+1)
  {
    const int row = tid;
    uint32_t h[16];
    for (int half = 0; half < 2; ++half) {
      for (int j = 0; j < 16; ++j) {
        const __half2 v = __halves2half2(A[(rank * 128 + row) * 64 + 32 * half + 2 * j], A[(rank * 128 + row) * 64 + 32 * half + 2 * j + 1]);
        h[j] = *reinterpret_cast<const uint32_t*>(&v);
      }
      tmem_st16((uint32_t(warp * 32) << 16) + tbase + 256 + 16 * half, h);
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync();   // barriers initialised and visible cluster-wide before any remote arrive / multicast commit
  if (warp == 0) {
    if (rank == 1) {
      if (lane == 0) mbar_arrive_remote(mapa(smem_u32(ready), 0));      // follower: operands in place
    } else {
      if (lane == 0) mbar_arrive(ready);
      mbar_wait_cluster(ready, 0);
      tc_fence_after_sync();
      const bool leader = lane == 0;
      if (leader) {
        const uint32_t idesc = make_idesc_f16(256, 128);
        const uint32_t blo = sdesc_lo(smem_u32(sB), 128);
        constexpr uint32_t HI_B = sdesc_hi(1024);
        for (int j = 0; j < 4; ++j) umma_ts2_pair(tbase + 0, tbase + 256 + j * 8, blo + j * 16, HI_B, idesc, j ? 1u : 0u);
        umma_commit_pair(done);
      }
    }
  }
  mbar_wait_cluster(done, 0);   // both CTAs: the multicast commit
  tc_fence_after_sync();
  {
    uint32_t u[32];
    for (int c = 0; c < 128; c += 32) {
      tmem_ld32((uint32_t(warp * 32) << 16) + tbase + c, u);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(size_t(rank) * 128 + tid) * 128 + c + j] = __uint_as_float(u[j]);
    }
  }
  if (tid == 0) flags[rank] = int(tbase);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tbase, 512);
}

int main() {
  std::vector<__half> hA(256 * 64), hB(128 * 64);
  for (int i = 0; i < 256 * 64; ++i) hA[i] = __float2half(float((i * 7 + (i >> 6)) % 5 - 2));
  for (int i = 0; i < 128 * 64; ++i) hB[i] = __float2half(float((i * 3 + (i >> 5)) % 7 - 3));
  __half *dA, *dB;
  float* dO;
  int* dF;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 2 * 128 * 128 * 4);
  cudaMalloc(&dF, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, 2 * 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 256);
  probe<<<2, 128, 8192 + 256>>>(dA, dB, dO, dF);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PAIR probe: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> o(2 * 128 * 128);
  int f[2];
  cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(f, dF, 8, cudaMemcpyDeviceToHost);
  double worst = 0;
  int bad = 0;
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 128; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += __half2float(hA[(c * 128 + r) * 64 + k]) * __half2float(hB[n * 64 + k]);
        const double d = fabs(double(o[(size_t(c) * 128 + r) * 128 + n]) - ref);
        if (d > worst) worst = d;
        if (d != 0 && bad < 5) {
          printf("  mismatch cta %d row %d col %d: got %g want %g\n", c, r, n, o[(size_t(c) * 128 + r) * 128 + n], ref);
          ++bad;
        }
      }
  printf("PAIR probe: tmem base %d / %d, max |D - A.B^T| = %g  %s\n", f[0], f[1], worst, worst == 0 ? "PASS" : "FAIL");
  return worst == 0 ? 0 : 2;
}
