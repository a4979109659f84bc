// Standalone sm_100a probe for the mixed-precision MLP mode (fp16 main term + two fp8 correction terms):
//   (1) kind::f8f6f4 operand conventions: 8-bit K-major no-swizzle shared-memory tiles (core matrix = 8 rows x 16 bytes),
//       8-bit A operand in TMEM (four K-consecutive bytes per column), both against a CPU matmul;
//   (2) kind::f16 and kind::f8f6f4 instructions accumulating into the SAME TMEM accumulator;
//   (3) sustained issue rate (and, with nvidia-smi running beside it, clocks / power) of the per-K-chunk instruction mix
//       of the current kernel (12 x f16) against the mixed one (4 x f16 + 4 x f8) on all SMs with random operands.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp8_probe tools/fp8_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

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

static float e4m3_to_float(uint8_t v) {
  const int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
  float f = e == 0 ? ldexpf(m / 8.f, -6) : ldexpf(1.f + m / 8.f, e - 7);
  return s ? -f : f;
}

__device__ __forceinline__ uint32_t off16(int r, int k, int K) { return (r >> 3) * (K / 8) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }
__device__ __forceinline__ uint32_t off8(int r, int k, int K) { return (r >> 3) * (K / 16) * 128 + (k >> 4) * 128 + (r & 7) * 16 + (k & 15); }

constexpr int K = 64, N = 128;

__device__ __forceinline__ bool elect_one_sync_() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
  return p != 0;
}

// variant 0: SS f8   1: TS f8   2: TS f16 + TS f8 into one accumulator   3: TS f16 only
__global__ void __launch_bounds__(128) check_kernel(const uint8_t* A8, const uint8_t* B8, const __half* A16, const __half* B16, float* D,
                                                    int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA8 = smem;                 // 8 KB
  uint8_t* sB8 = smem + 8192;          // 8 KB
  uint8_t* sB16 = smem + 16384;        // 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    sA8[off8(r, k, K)] = A8[i];
    sB8[off8(r, k, K)] = B8[i];
    *reinterpret_cast<__half*>(sB16 + off16(r, k, K)) = B16[i];
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tmem_base_s != 0u) __trap();
  const uint32_t tlane = uint32_t(warp * 32) << 16;
  {  // thread = row: A operands into TMEM
    uint32_t w16[32], w8[16];
    for (int j = 0; j < 32; ++j) {
      const __half2 h = __halves2half2(A16[tid * K + 2 * j], A16[tid * K + 2 * j + 1]);
      w16[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    for (int j = 0; j < 16; ++j)
      w8[j] = uint32_t(A8[tid * K + 4 * j]) | (uint32_t(A8[tid * K + 4 * j + 1]) << 8) | (uint32_t(A8[tid * K + 4 * j + 2]) << 16) |
              (uint32_t(A8[tid * K + 4 * j + 3]) << 24);
    tmem_st32(tlane + 256, w16);
    tmem_st16(tlane + 384, w8);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t i16 = make_idesc_f16(128, N), i8 = make_idesc_f8(128, N);
    const uint32_t a8 = sdesc_lo(smem_u32(sA8), 128), b8 = sdesc_lo(smem_u32(sB8), 128), b16 = sdesc_lo(smem_u32(sB16), 128);
    constexpr uint32_t HI8 = sdesc_hi(512), HI16 = sdesc_hi(1024);
    uint32_t acc = 0;
    if (variant == 2 || variant == 3)
      for (int j = 0; j < 4; ++j) {
        umma_ts2(0u, 256 + j * 8, b16 + j * 16, HI16, i16, acc);
        acc = 1;
      }
    if (variant == 0)
      for (int j = 0; j < 2; ++j) {
        umma_ss2_f8(0u, a8 + j * 16, HI8, b8 + j * 16, HI8, i8, acc);
        acc = 1;
      }
    if (variant == 1 || variant == 2)
      for (int j = 0; j < 2; ++j) {
        umma_ts2_f8(0u, 384 + j * 8, b8 + j * 16, HI8, i8, acc);
        acc = 1;
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 32) {
    uint32_t u[32];
    tmem_ld32(tlane + c, u);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * N + c + j] = __uint_as_float(u[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

// mode 0: 12 x f16 per K chunk (hi.hi, lo.hi, hi.lo)   1: 4 x f16 + 2 x f8 + 2 x f8   2: 4 x f16 only   3: 8 x f16
__global__ void __launch_bounds__(128) rate_kernel(int mode, int groups, int reps, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  // B operands: [hi fp16 16 KB | lo fp16 16 KB | f8 8 KB | f8 8 KB], pseudo-random contents
  uint32_t seed = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
  auto rnd = [&]() {
    seed ^= seed << 13;
    seed ^= seed >> 17;
    seed ^= seed << 5;
    return seed;
  };
  for (int i = tid; i < 16384; i += 128) {  // fp16 values in (-1, 1): exponent field 8..14, random mantissa / sign
    const uint32_t r = rnd();
    reinterpret_cast<uint16_t*>(smem)[i] = uint16_t((r & 0x83FF) | (((r >> 16) % 7 + 8) << 10));
  }
  for (int i = tid; i < 16384; i += 128) {
    uint8_t b = uint8_t(rnd());
    if ((b & 0x7F) == 0x7F) b ^= 1;
    smem[32768 + i] = b;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tlane = uint32_t(warp * 32) << 16;
  for (int c = 256; c < 512; c += 32) {  // A operands: AHI [256,384) fp16; [384,512) fp16 lo (mode 0) or two f8 arrays (mode 1)
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) {
      const uint32_t r = rnd();
      if (c < 384 || mode != 1) {
        const uint32_t lo = (r & 0x83FF) | (((r >> 10) % 7 + 8) << 10), hi = ((r >> 16) & 0x83FF) | (((r >> 26) % 7 + 8) << 10);
        w[j] = lo | (hi << 16);
      } else {
        uint32_t v = r;
        for (int b = 0; b < 4; ++b)
          if (((v >> (8 * b)) & 0x7F) == 0x7F) v ^= 1u << (8 * b);
        w[j] = v;
      }
    }
    tmem_st32(tlane + c, w);
  }
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const unsigned long long t0 = clock64();
  if (warp == 0) {
    const bool leader = elect_one_sync_();
    const uint32_t i16 = make_idesc_f16(128, 128), i8 = make_idesc_f8(128, 128);
    const uint32_t bh = sdesc_lo(smem_u32(smem), 128), bl = bh + (16384 >> 4), b8h = bh + (32768 >> 4), b8l = b8h + (8192 >> 4);
    constexpr uint32_t HI8 = sdesc_hi(512), HI16 = sdesc_hi(1024);
    uint32_t phase = 0;
    for (int rep = 0; rep < reps; ++rep) {
      if (leader) {
        for (int g = 0; g < groups; ++g) {
          const uint32_t acc = (g & 1) ? 128u : 0u;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t ah = 256 + kc * 32, al = 384 + kc * 32, a8l = 384 + kc * 16, a8h = 448 + kc * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              umma_ts2(acc, ah + j * 8, bh + j * 16, HI16, i16, (kc | j) ? 1u : 0u);
              if (mode == 0 || mode == 3) umma_ts2(acc, al + j * 8, bh + j * 16, HI16, i16, 1u);
              if (mode == 0) umma_ts2(acc, ah + j * 8, bl + j * 16, HI16, i16, 1u);
            }
            if (mode == 1) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                umma_ts2_f8(acc, a8l + j * 8, b8h + j * 16, HI8, i8, 1u);
                umma_ts2_f8(acc, a8h + j * 8, b8l + j * 16, HI8, i8, 1u);
              }
            }
          }
        }
        umma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
  // ---------------------------------------------------------------- (1), (2) numerics
  std::vector<uint8_t> A8(128 * K), B8(128 * K);
  std::vector<__half> A16(128 * K), B16(128 * K);
  srand(7);
  for (int i = 0; i < 128 * K; ++i) {
    auto r8 = []() {
      uint8_t b = uint8_t(rand());
      b = (b & 0x87) | (uint8_t(5 + rand() % 5) << 3);  // |value| in [2^-2, 2^3)
      return b;
    };
    A8[i] = r8();
    B8[i] = r8();
    A16[i] = __float2half((rand() % 2001 - 1000) / 500.f);
    B16[i] = __float2half((rand() % 2001 - 1000) / 500.f);
  }
  uint8_t *dA8, *dB8;
  __half *dA16, *dB16;
  float* dD;
  CK(cudaMalloc(&dA8, 128 * K));
  CK(cudaMalloc(&dB8, 128 * K));
  CK(cudaMalloc(&dA16, 128 * K * 2));
  CK(cudaMalloc(&dB16, 128 * K * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA8, A8.data(), 128 * K, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB8, B8.data(), 128 * K, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA16, A16.data(), 128 * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB16, B16.data(), 128 * K * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const char* names[4] = {"SS f8", "TS f8", "TS f16 + TS f8 (one accumulator)", "TS f16"};
  for (int v = 0; v < 4; ++v) {
    CK(cudaMemset(dD, 0, 128 * N * 4));
    check_kernel<<<1, 128, 32768>>>(dA8, dB8, dA16, dB16, dD, v);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) {
          if (v != 3) ref += double(e4m3_to_float(A8[m * K + k])) * e4m3_to_float(B8[n * K + k]);
          if (v >= 2) ref += double(__half2float(A16[m * K + k])) * __half2float(B16[n * K + k]);
        }
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("check %-36s max|ref| %.3f  max err %.3e  %s\n", names[v], maxref, maxerr, maxerr <= 1e-4 * maxref ? "OK" : "MISMATCH");
  }
  // ---------------------------------------------------------------- (3) sustained rate
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* dC;
  CK(cudaMalloc(&dC, sms * 8));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
  const char* mnames[4] = {"12 x f16 (fp16x3)", "4 x f16 + 4 x f8 (mixed)", "4 x f16 (fast)", "8 x f16"};
  const int order[6] = {0, 1, 2, 3, 1, 0};
  for (int oi = 0; oi < 6; ++oi) {
    const int mode = order[oi];
    const int groups = 64;
    // calibrate reps for ~`seconds`: a K chunk of 12 MMAs is ~0.5 us
    int reps = 50;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int pass = 0; pass < 2; ++pass) {
      CK(cudaEventRecord(e0));
      rate_kernel<<<sms, 128, 49152>>>(mode, groups, reps, dC);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (pass == 0) reps = int(reps * (seconds * 1000.0 / ms)) + 1;
    }
    std::vector<unsigned long long> cyc(sms);
    CK(cudaMemcpy(cyc.data(), dC, sms * 8, cudaMemcpyDeviceToHost));
    double mean_cyc = 0;
    for (auto c : cyc) mean_cyc += double(c) / sms;
    const double chunks = double(reps) * groups * 4;
    printf("rate  %-28s %.3f s  %.1f ns / K-chunk  %.0f cycles / K-chunk  SM clock %.0f MHz\n", mnames[mode], ms / 1000.0,
           ms * 1e6 / chunks, mean_cyc / chunks, mean_cyc / (ms * 1e3));
  }
  return 0;
}
