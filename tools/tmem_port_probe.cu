// Does tcgen05.ld (accumulator drain) share a TMEM read port with the A-operand fetch of TS-form tcgen05.mma?
// One CTA per SM: warp 10 issues a long stream of MMAs (A from TMEM or from shared memory, N = 128 or 256), warps 0-7
// run a tcgen05.ld loop over the accumulator columns; both rates are reported alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tmem_port_probe tools/tmem_port_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"
using namespace nsr;

constexpr int SMEM = 160 * 1024 + 1024;

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
  return p != 0;
}

// mma_mode: 0 none, 1 TS N=128, 2 SS N=128, 3 TS N=256, 4 TS f8 N=128 (K=32)     ld_mode: 0 none, 1 ld only, 2 ld + st
__global__ void __launch_bounds__(384, 1) probe(int n_mma, int mma_mode, int ld_mode, long long fixed_cycles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(done + 1);
  uint32_t* tmem_slot = const_cast<uint32_t*>(stop) + 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(done, 1);
    *stop = 0;
    fence_mbar_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 160 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp < 8) {  // something finite in the A operand columns
    uint32_t h[16];
    for (int j = 0; j < 16; ++j) h[j] = 0x3c003c00u;
    const uint32_t tl = uint32_t((warp & 3) * 32) << 16;
    for (int c = 256 + (warp >> 2) * 128; c < 384 + (warp >> 2) * 128; c += 16) tmem_st16(tl + c, h);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 10) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    if (mma_mode != 0) {
      const uint32_t idesc = mma_mode == 3 ? make_idesc_f16(128, 256) : (mma_mode == 4 ? make_idesc_f8(128, 128) : make_idesc_f16(128, 128));
      constexpr uint32_t HI_B = sdesc_hi(1024), HI_8 = sdesc_hi(512);
      const uint32_t b_lo = sdesc_lo(smem_u32(smem), 128), a_lo = sdesc_lo(smem_u32(smem) + 65536, 128);
      if (leader) {
        for (int i = 0; i < n_mma; i += 16) {
          const uint32_t acc = mma_mode == 3 ? 0u : ((i >> 5) & 1) * 128;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t a_t = 256 + j * 8, koff = (j & 3) * 16;
            if (mma_mode == 2) umma_ss2(acc, a_lo + koff, HI_B, b_lo + koff, HI_B, idesc, 1u);
            else if (mma_mode == 4) umma_ts2_f8(acc, a_t, b_lo + koff, HI_8, idesc, 1u);
            else umma_ts2(acc, a_t, b_lo + koff, HI_B, idesc, 1u);
          }
        }
        umma_commit(done);
      }
      mbar_wait(done, 0);
    } else {
      while (clock64() - t0 < fixed_cycles) {
      }
    }
    const long long t1 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      *stop = 1;
    }
  } else if (warp < 8 && ld_mode != 0) {
    const uint32_t tl = uint32_t((warp & 3) * 32) << 16;
    const uint32_t col0 = (warp >> 2) * 64;
    uint32_t u0[32], u1[32], h[16];
    long long iters = 0;
    uint32_t sink = 0;
    while (!*stop) {
      tmem_ld32(tl + col0, u0);
      tmem_ld32(tl + col0 + 32, u1);
      tmem_ld_wait();
      sink ^= u0[0] ^ u1[31] ^ u0[17];
      if (ld_mode == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = 0x3c003c00u | (sink & 1u);
        tmem_st16(tl + 256 + col0 / 2, h);
        tmem_st16(tl + 384 + col0 / 2, h);
        tmem_st_wait();
      }
      ++iters;
    }
    if (lane == 0) {
      if (sink == 0x12345u) iters = -1;
      atomicAdd(reinterpret_cast<unsigned long long*>(&out[blockIdx.x * 2 + 1]), (unsigned long long)iters);
    }
  }
  __syncthreads();
  if (warp == 10) tmem_dealloc(0u, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int n_mma = 48000;
  const char* mn[] = {"no MMA", "TS N=128 f16", "SS N=128 f16", "TS N=256 f16", "TS N=128 f8"};
  const char* ln[] = {"no ld", "ld", "ld+st"};
  for (int grid : {1, 148})
    for (int mm = 0; mm < 5; ++mm)
      for (int ld = 0; ld < 3; ++ld) {
        if (mm == 0 && ld == 0) continue;
        cudaMemset(out, 0, 148 * 16);
        probe<<<grid, 384, SMEM>>>(n_mma, mm, ld, 3000000, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d/%d: %s\n", mm, ld, cudaGetErrorString(e));
          return 1;
        }
        std::vector<long long> c(grid * 2);
        cudaMemcpy(c.data(), out, grid * 16, cudaMemcpyDeviceToHost);
        // each ld iteration of one warp moves 64 columns x 32 lanes x 4 B = 8 KB
        printf("grid=%3d  %-13s %-6s : %7.1f cycles / MMA   ld %6.1f B/cycle/SM\n", grid, mn[mm], ln[ld],
               mm ? double(c[0]) / n_mma : 0.0, double(c[1]) * 8192.0 / double(c[0]));
      }
  return 0;
}
