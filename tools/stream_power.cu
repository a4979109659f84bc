// How much board power does streaming the packed weights L2 -> shared memory cost at the rate the MLP kernel needs
// (~43 B/clk/SM)?  Runs the stream for a few seconds at a throttled rate; sample nvidia-smi power.draw alongside.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/stream_power tools/stream_power.cu
#include <cstdio>
#include <cstdlib>
#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"
using namespace nsr;
constexpr int STAGES = 5, STAGE = 32768;
__global__ void __launch_bounds__(128, 1) stream(const uint8_t* src, long long chunks, int gap_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (long long i = 0; i < chunks + STAGES; ++i) {
      int s = int(i % STAGES);
      if (i >= STAGES) mbar_wait(&full[s], uint32_t(((i - STAGES) / STAGES) & 1));
      if (i < chunks) {
        long long t0 = clock64();
        while (clock64() - t0 < gap_cycles) {}
        mbar_arrive_expect_tx(&full[s], STAGE);
        bulk_g2s(smem + s * STAGE, src + (i % 73) * (long long)STAGE, STAGE, &full[s]);
      }
    }
  }
}
int main(int argc, char** argv) {
  int gap = argc > 1 ? atoi(argv[1]) : 768;       // cycles between copies: 32 KB / 768 = 42.7 B/clk/SM
  double seconds = argc > 2 ? atof(argv[2]) : 4.0;
  uint8_t* src; cudaMalloc(&src, 73 * STAGE); cudaMemset(src, 0x5a, 73 * STAGE);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE + 256);
  long long chunks = (long long)(seconds * 1.9e9 / (gap > 0 ? gap : 350));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  stream<<<148, 128, STAGES * STAGE + 256>>>(src, chunks, gap);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("gap %d cycles: %.2f s, %.2f TB/s aggregate L2->smem\n", gap, ms / 1e3, double(chunks) * STAGE * 148 / (ms * 1e-3) / 1e12);
  return 0;
}
