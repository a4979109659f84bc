// Standalone sm_100a probe: validates the UMMA descriptor conventions used by the render kernels
// (SS / TS operand modes, no-swizzle and 128B-swizzle K-major layouts) against a CPU matmul, and
// measures (a) tcgen05.mma issue rate for the tile shapes the MLP kernel uses, (b) L2 -> shared
// bulk-copy bandwidth when every SM streams the same weight blob.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/umma_probe tools/umma_probe.cu
// Run (GPU box): ./tools/umma_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neural-sim-nerf_b200/csrc/sm100_prims.cuh"

using namespace nsr;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

enum { V_SS_NOSW = 0, V_SS_NOSW_SWAPPED = 1, V_SS_SW128 = 2, V_TS_NOSW = 3, V_TS_NOSW_SWAPPED = 4, V_TS_SW128 = 5, V_SS_MN_NOSW = 6, V_SS_MN_NOSW_SWAPPED = 7 };

// byte offset of element (row r, k) in a K-major operand tile with `K` columns
__device__ __forceinline__ uint32_t off_nosw(int r, int k, int K) {
  return (r >> 3) * (K / 8) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
}
// MN-major, no swizzle: 8x8 core matrices stored [k][mn] (mn contiguous, 16 B per k); mn-blocks 128 B apart,
// k-blocks (rows/8)*128 B apart
__device__ __forceinline__ uint32_t off_mn_nosw(int r, int k, int rows) {
  return (k >> 3) * (rows / 8) * 128 + (r >> 3) * 128 + (k & 7) * 16 + (r & 7) * 2;
}
__device__ __forceinline__ uint32_t off_sw128(int r, int k) {  // K == 64 slab
  int chunk = (k >> 3) ^ (r & 7);
  return (r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (k & 7) * 2;
}

template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const __half* A, const __half* B, float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool sw128 = (variant == V_SS_SW128 || variant == V_TS_SW128);
  const bool swapped = (variant == V_SS_NOSW_SWAPPED || variant == V_TS_NOSW_SWAPPED);
  const bool ts = variant >= V_TS_NOSW && variant <= V_TS_SW128;
  const bool mn = variant >= V_SS_MN_NOSW;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t o = mn ? off_mn_nosw(r, k, 128) : (sw128 ? off_sw128(r, k) : off_nosw(r, k, K));
    *reinterpret_cast<__half*>(sA + o) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t o = mn ? off_mn_nosw(r, k, N) : (sw128 ? off_sw128(r, k) : off_nosw(r, k, K));
    *reinterpret_cast<__half*>(sB + o) = B[i];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t d_tmem = tmem_base;          // columns [0, N)
  const uint32_t a_tmem = tmem_base + 256;    // columns [256, 256 + K/2)

  if (ts) {  // thread = row: pack K halves into K/2 words and store them to TMEM
    uint32_t regs[32];
    static_assert(K == 64, "probe packs exactly 32 words");
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      __half2 h = __halves2half2(A[tid * K + 2 * j], A[tid * K + 2 * j + 1]);
      regs[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st32(a_tmem + (uint32_t(warp * 32) << 16), regs);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N) | (mn ? ((1u << 15) | (1u << 16)) : 0u);
    const uint32_t s_k = 128, s_mn = (K / 8) * 128;
    for (int k = 0; k < K / 16; ++k) {
      uint64_t bdesc, adesc;
      if (mn) {  // k-block stride: (rows/8)*128, mn-block stride: 128; a K=16 step spans two k-blocks
        const uint32_t ka = (128 / 8) * 128, kb = (N / 8) * 128;
        const bool sw = (variant == V_SS_MN_NOSW_SWAPPED);
        adesc = make_sdesc(smem_u32(sA) + k * 2 * ka, sw ? 128 : ka, sw ? ka : 128, 0);
        bdesc = make_sdesc(smem_u32(sB) + k * 2 * kb, sw ? 128 : kb, sw ? kb : 128, 0);
      } else if (sw128) {
        adesc = make_sdesc(smem_u32(sA) + k * 32, 16, 1024, 2);
        bdesc = make_sdesc(smem_u32(sB) + k * 32, 16, 1024, 2);
      } else {
        uint32_t lbo = swapped ? s_mn : s_k, sbo = swapped ? s_k : s_mn;
        adesc = make_sdesc(smem_u32(sA) + k * 2 * s_k, lbo, sbo, 0);
        bdesc = make_sdesc(smem_u32(sB) + k * 2 * s_k, lbo, sbo, 0);
      }
      if (ts)
        umma_ts(d_tmem, a_tmem + k * 8, bdesc, idesc, k > 0);
      else
        umma_ss(d_tmem, adesc, bdesc, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(d_tmem + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark: one thread issues `iters` x (K=256 -> 16) MMAs back to back.
template <int N, bool TS, bool SW>
__global__ void __launch_bounds__(128) tput_kernel(int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < (128 * 256 * 2 + N * 256 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 256 * 2;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        uint64_t adesc, bdesc;
        if (SW) {
          adesc = make_sdesc(smem_u32(sA) + (k >> 2) * 128 * 128 + (k & 3) * 32, 16, 1024, 2);
          bdesc = make_sdesc(smem_u32(sB) + (k >> 2) * N * 128 + (k & 3) * 32, 16, 1024, 2);
        } else {
          adesc = make_sdesc(smem_u32(sA) + k * 256, 128, 32 * 128, 0);
          bdesc = make_sdesc(smem_u32(sB) + k * 256, 128, 32 * 128, 0);
        }
        // D: columns [0,N) for N<=256; A (TS) in columns [256, 384)
        if (TS)
          umma_ts(tmem_base, tmem_base + 256 + k * 8, bdesc, idesc, 1);
        else
          umma_ss(tmem_base, adesc, bdesc, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// L2 -> shared streaming: every CTA walks the same `total` bytes in `chunk`-byte bulk copies
// through a 4-deep ring, `reps` times.  Reports bytes / clock / SM.
__global__ void __launch_bounds__(128) stream_kernel(const uint8_t* src, int total, int chunk, int reps, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int nchunks = total / chunk;
    const long long n = (long long)nchunks * reps;
    long long t0 = clock64();
    // keep 4 copies in flight; a slot is reused as soon as its previous copy has landed
    for (long long i = 0; i < n + 4; ++i) {
      int s = int(i & 3);
      if (i >= 4) mbar_wait(&full[s], uint32_t(((i - 4) >> 2) & 1));
      if (i < n) {
        mbar_arrive_expect_tx(&full[s], chunk);
        bulk_g2s(smem + s * chunk, src + (i % nchunks) * (long long)chunk, chunk, &full[s]);
      }
    }
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
}

// ---------------------------------------------------------------------------------------------
template <int N, int K>
static double run_variant(int variant, const std::vector<__half>& hA, const std::vector<__half>& hB, const std::vector<float>& ref) {
  __half *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, 128 * K * 2));
  CK(cudaMalloc(&dB, N * K * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), 128 * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, 128 * N * 4));
  size_t smem = 128 * K * 2 + N * K * 2 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> out(128 * N);
  CK(cudaMemcpy(out.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int i = 0; i < 128 * N; ++i) {
    double d = std::fabs((double)out[i] - ref[i]);
    if (!(d == d)) d = 1e30;
    if (d > maxerr) maxerr = d;
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return maxerr;
}

template <int N>
static void correctness() {
  constexpr int K = 64;
  std::vector<__half> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K), ref(128 * N);
  srand(1234 + N);
  for (int i = 0; i < 128 * K; ++i) {
    fA[i] = float((rand() % 17) - 8) / 8.0f;
    hA[i] = __float2half(fA[i]);
  }
  for (int i = 0; i < N * K; ++i) {
    fB[i] = float((rand() % 17) - 8) / 8.0f;
    hB[i] = __float2half(fB[i]);
  }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int k = 0; k < K; ++k) acc += fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = acc;
    }
  const char* names[] = {"SS nosw (LBO=K-stride,SBO=MN-stride)", "SS nosw swapped", "SS sw128", "TS nosw", "TS nosw swapped", "TS sw128",
                         "SS MN-major nosw (LBO=K-block stride,SBO=MN-block stride)", "SS MN-major nosw swapped"};
  for (int v = 0; v < 8; ++v) {
    double e = run_variant<N, K>(v, hA, hB, ref);
    printf("PROBE N=%d variant %d [%s]: max_abs_err=%g %s\n", N, v, names[v], e, e == 0.0 ? "PASS" : "FAIL");
  }
}

template <int N, bool TS, bool SW>
static void tput(int grid) {
  long long* dc;
  CK(cudaMalloc(&dc, grid * sizeof(long long)));
  size_t smem = 128 * 256 * 2 + N * 256 * 2 + 1024;
  CK(cudaFuncSetAttribute(tput_kernel<N, TS, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  tput_kernel<N, TS, SW><<<grid, 128, smem>>>(iters, dc);  // warm
  tput_kernel<N, TS, SW><<<grid, 128, smem>>>(iters, dc);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto x : c) mx = x > mx ? x : mx;
  double per = double(mx) / (double(iters) * 16);
  printf("TPUT N=%d %s %s grid=%d: %.1f cycles/MMA(K=16) -> %.0f MAC/clk/SM (ideal 4096)\n", N, TS ? "TS" : "SS", SW ? "sw128" : "nosw", grid,
         per, 128.0 * N * 16 / per);
  cudaFree(dc);
}

static void stream(int grid, int total, int chunk) {
  uint8_t* src;
  long long* dc;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  CK(cudaMalloc(&dc, grid * sizeof(long long)));
  size_t smem = 4 * chunk + 1024;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int reps = 40;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  stream_kernel<<<grid, 128, smem>>>(src, total, chunk, reps, dc);
  cudaEventRecord(e0);
  stream_kernel<<<grid, 128, smem>>>(src, total, chunk, reps, dc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  CK(cudaMemcpy(c.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto x : c) mx = x > mx ? x : mx;
  double bytes = double(total / chunk) * chunk * reps;
  printf("STREAM grid=%d total=%d chunk=%d: %.1f B/clk/SM, aggregate %.2f TB/s (%.3f ms)\n", grid, total, chunk, bytes / double(mx),
         bytes * grid / (ms * 1e-3) / 1e12, ms);
  cudaFree(src);
  cudaFree(dc);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  correctness<128>();
  correctness<256>();
  if (getenv("PROBE_QUICK")) return 0;
  for (int grid : {1, 148}) {
    tput<256, false, false>(grid);
    tput<256, false, true>(grid);
    tput<128, false, false>(grid);
    tput<128, false, true>(grid);
    tput<128, true, false>(grid);
    tput<128, true, true>(grid);
    tput<256, true, false>(grid);
    tput<256, true, true>(grid);
  }
  for (int grid : {1, 148, 296}) {
    stream(grid, 1196032, 16384);
    stream(grid, 1196032, 32768);
  }
  stream(148, 1196032, 8192);
  return 0;
}
