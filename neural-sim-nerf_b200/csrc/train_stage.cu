// Device-side pieces of one optimisation step of the NeRF training loop (RN:643-716) around the forward /
// backward kernels: counter-based random numbers for the stratified jitter (RN:447-461), the inverse-CDF draws (RH:211) and
// the sigma noise (RN:365-366); img2mse (RH:12) with its gradient; Adam (torch.optim.Adam as RN:287 configures it) over all
// parameter tensors of a network in one launch.
#include <math.h>

#include "common.cuh"

namespace nsr {

// ----------------------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
// counter = (index lo, index hi, stream, 0), key = seed: every element of every random tensor is addressable from
// (seed, stream, index) alone, so results do not depend on the launch shape.
struct U4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint64_t index, uint32_t stream, uint64_t seed) {
  uint32_t c0 = uint32_t(index), c1 = uint32_t(index >> 32), c2 = stream, c3 = 0u;
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u01(uint32_t x) { return float(x >> 8) * 5.9604644775390625e-8f; }             // [0, 1)
__device__ __forceinline__ float u01_open(uint32_t x) { return (float(x >> 8) + 1.0f) * 5.9604644775390625e-8f; }  // (0, 1]

// out[i] ~ U[0,1), four values per Philox call: out[4q + j] = component j of philox(q, stream, seed)
__global__ void uniform_kernel(uint64_t seed, uint32_t stream, float* __restrict__ out, int64_t count) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (q * 4 >= count) return;
  const U4 r = philox4x32_10(uint64_t(q), stream, seed);
  const float v[4] = {u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (q * 4 + j < count) out[q * 4 + j] = v[j];
}

// raw[p].sigma += std * N(0,1)   (RN:365-366: noise = randn(raw[...,3].shape) * raw_noise_std), Box-Muller on one Philox call
__global__ void sigma_noise_kernel(uint64_t seed, uint32_t stream, float4* __restrict__ raw, int64_t n_points, float std) {
  const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p >= n_points) return;
  const U4 r = philox4x32_10(uint64_t(p), stream, seed);
  const float g = sqrtf(-2.0f * logf(u01_open(r.x))) * cospif(2.0f * u01(r.y));
  raw[p].w += std * g;
}

int launch_uniform(uint64_t seed, uint32_t stream, float* out, int64_t count, cudaStream_t st) {
  if (count == 0) return NSR_OK;
  const int64_t quads = (count + 3) / 4;
  uniform_kernel<<<unsigned((quads + 255) / 256), 256, 0, st>>>(seed, stream, out, count);
  count_launch();
  return check_launch("uniform_kernel");
}

int launch_sigma_noise(uint64_t seed, uint32_t stream, float* raw, int64_t n_points, float std, cudaStream_t st) {
  if (n_points == 0) return NSR_OK;
  sigma_noise_kernel<<<unsigned((n_points + 255) / 256), 256, 0, st>>>(seed, stream, reinterpret_cast<float4*>(raw), n_points, std);
  count_launch();
  return check_launch("sigma_noise_kernel");
}

// ----------------------------------------------------------------------------- img2mse (RH:12) and its gradient
// loss = mean((x - y)^2) over `count` values; d_x = 2 (x - y) / count.  One block, fixed reduction order: deterministic.
constexpr int MSE_THREADS = 1024;

__global__ void __launch_bounds__(MSE_THREADS) mse_grad_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t count,
                                                               float* __restrict__ d_x, float* __restrict__ loss) {
  const float inv = 1.0f / float(count);
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < count; i += MSE_THREADS) {
    const float d = x[i] - y[i];
    s = fmaf(d, d, s);
    d_x[i] = 2.0f * d * inv;
  }
  __shared__ float sm[MSE_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *loss = t * inv;
  }
}

int launch_mse_grad(const float* x, const float* y, int64_t count, float* d_x, float* loss, cudaStream_t st) {
  mse_grad_kernel<<<1, MSE_THREADS, 0, st>>>(x, y, count, d_x, loss);
  count_launch();
  return check_launch("mse_grad_kernel");
}

// ----------------------------------------------------------------------------- Adam, all tensors of up to two networks in one launch
// torch.optim.Adam (no weight decay, no amsgrad), the operation order of its single-tensor implementation:
//   m <- m + (g - m)(1 - b1);  v <- v b2 + (1 - b2) g g;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(AdamJobs jobs, float one_minus_b1, float b2, float one_minus_b2, float step_size, float bc2_sqrt, float eps) {
  const AdamJob j = jobs.j[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += gridDim.x * blockDim.x) {
    const float g = j.g[i];
    const float m = j.m[i] + (g - j.m[i]) * one_minus_b1;
    const float v = __fadd_rn(__fmul_rn(j.v[i], b2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
    j.m[i] = m;
    j.v[i] = v;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
    j.p[i] = __fadd_rn(j.p[i], __fmul_rn(-step_size, __fdiv_rn(m, denom)));
  }
}

int launch_adam(const AdamJobs& jobs, float beta1, float beta2, float lr, float eps, int64_t step, cudaStream_t st) {
  if (jobs.count == 0) return NSR_OK;
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  adam_kernel<<<dim3(8, jobs.count), 256, 0, st>>>(jobs, float(1.0 - double(beta1)), beta2, float(1.0 - double(beta2)), float(double(lr) / bc1),
                                                   float(sqrt(bc2)), eps);
  count_launch();
  return check_launch("adam_kernel");
}

}  // namespace nsr
