// (single-pass variant of mma_rate_probe.cu: PASSES MMAs per (A,B) k-slice, 16 KB or 32 KB stages, optional commit thinning)
// What slows tcgen05.mma below 64 cycles per M128 N128 K16 instruction inside the MLP kernel?  Reproduces its MMA issue
// pattern (A from TMEM hi/lo, B hi/lo from a shared-memory ring, 12 MMAs per chunk) and switches the co-running
// activities on one by one: (1) TMA weight streaming into the ring, (2) epilogue-like tcgen05.ld / tcgen05.st traffic
// from 8 warps, (4) encoder-like sincosf ALU load from 2 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate_probe tools/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"
using namespace nsr;

constexpr int STAGES = 5, STAGE_BYTES = 32768, CHUNK_BYTES = 16384;
// passes: 1 = one MMA per k-slice (tier 1), 3 = hi/lo split; commit_every: tcgen05.commit on the ring's empty barrier every n-th chunk only

constexpr int SMEM = STAGES * STAGE_BYTES + 1024;

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
  return p != 0;
}

__global__ void __launch_bounds__(384, 1) rate_kernel(const uint8_t* blob, int chunks, int mode, long long* out, float* sink, int passes, int commit_every, int acc_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(done + 1);
  uint32_t* tmem_slot = const_cast<uint32_t*>(stop) + 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    *stop = 0;
    fence_mbar_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < STAGES * STAGE_BYTES / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const bool tma = mode & 1, tmem_traffic = mode & 2, alu = mode & 4;
  if (warp == 11) {
    if (lane == 0 && tma) {
      uint32_t stage = 0, phase = 0; bool first = true;
      for (int c = 0; c < chunks; ++c) {
        if (!first) mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        bulk_g2s(smem + stage * STAGE_BYTES, blob + size_t(c % 73) * STAGE_BYTES, STAGE_BYTES, &full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; first = false; }
      }
    }
  } else if (warp == 10) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, 128);
    constexpr uint32_t HI_B = sdesc_hi(1024);
    const uint32_t ring_lo = sdesc_lo(smem_u32(smem), 128);
    uint32_t stage = 0, phase = 0;
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      if (tma) mbar_wait(&full[stage], phase);
      const uint32_t bh = ring_lo + stage * (STAGE_BYTES >> 4), bl = bh + (CHUNK_BYTES >> 4);
      const uint32_t acc = ((c >> 2) & 1) * 128;
      const uint32_t ah = 256 + (c & 3) * 32, al = 384 + (c & 3) * 32;
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          umma_ts2(acc, ah + j * 8, bh + j * 16, HI_B, idesc, (acc_flag && j == 0 && (c & 3) == 0) ? 0u : 1u);
          if (passes == 3) {
            umma_ts2(acc, al + j * 8, bh + j * 16, HI_B, idesc, 1u);
            umma_ts2(acc, ah + j * 8, bl + j * 16, HI_B, idesc, 1u);
          }
        }
        if (tma || commit_every < 0) {
          if (commit_every <= 1 || (c % commit_every) == commit_every - 1 || tma) umma_commit(&empty[stage]);
        }
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(done);
    mbar_wait(done, 0);
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; *stop = 1; }
  } else if (warp >= 8) {
    if (alu) {
      float acc = 0.f, x = 0.001f * tid;
      while (!*stop) {
#pragma unroll 4
        for (int i = 0; i < 16; ++i) { float s, c; sincosf(x, &s, &c); acc += s * c; x += 0.37f; }
      }
      if (acc == 12345.f) sink[0] = acc;
    }
  } else {
    if (tmem_traffic) {
      // unused TMEM columns do not exist (all 512 are operands/accumulators): read the accumulators, rewrite the A operand
      const uint32_t tl = uint32_t((warp & 3) * 32) << 16;
      const uint32_t col0 = (warp >> 2) * 64;
      uint32_t u0[32], u1[32], h[16];
      while (!*stop) {
        tmem_ld32(tl + col0, u0);
        tmem_ld32(tl + col0 + 32, u1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = pack_f16x2(__uint_as_float(u0[2 * j]) * 0.f + 1.f, __uint_as_float(u1[2 * j]) * 0.f + 1.f);
        tmem_st16(tl + 256 + col0 / 2, h);
        tmem_st16(tl + 384 + col0 / 2, h);
        tmem_st_wait();
      }
    }
  }
  __syncthreads();
  if (warp == 10) tmem_dealloc(0u, 512);
}

int main() {
  uint8_t* blob; long long* out; float* sink;
  cudaMalloc(&blob, 73 * STAGE_BYTES); cudaMemset(blob, 0x3c, 73 * STAGE_BYTES);
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int chunks = 4000;
  for (int passes : {1, 3}) for (int commit_every : {0, -1}) for (int acc_flag : {0, 1}) {
  printf("== passes %d, %s, accumulate flag %s\n", passes, commit_every < 0 ? "commit per chunk even without TMA" : "commit only with TMA", acc_flag ? "cleared once per 4 chunks" : "always set");
  const char* names[] = {"MMA only", "+TMA ring", "+TMEM ld/st (8 warps)", "+TMA +TMEM", "+ALU (2 warps sincosf)", "+TMA +ALU", "+TMEM +ALU", "all"};
  for (int grid : {148}) for (int mode = 0; mode < 8; ++mode) {
    rate_kernel<<<grid, 384, SMEM>>>(blob, chunks, mode, out, sink, passes, commit_every, acc_flag);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<long long> c(grid); cudaMemcpy(c.data(), out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    printf("RATE grid=%3d mode %d [%-24s]: %.1f cycles per MMA\n", grid, mode, names[mode], double(mx) / (double(chunks) * 4 * passes));
  }
  }
  return 0;
}
