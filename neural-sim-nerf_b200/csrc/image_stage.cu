// Image-level prologue / epilogue of the renderer:
//   to8b (RH:14) on the device, so an image leaves the GPU as H*W*3 bytes instead of 12 bytes per pixel (RN:246);
//   the closed-form pull-back of dL/d(ray_batch) through get_rays (RH:156-165) and the view-direction normalisation
//   (RN:97) to dL/dc2w [3,4] -- what `torch.autograd.grad(batch_rays, categorical_prob, ...)` (RN:179-181) computes
//   through ~10 autograd nodes over [H,W,3] tensors, as one deterministic two-stage reduction;
//   ray generation from a c2w that lives on the device (the pose sampler's output never visits the host).
#include <math.h>

#include "common.cuh"

namespace nsr {

// ----------------------------------------------------------------------------- RH:14  to8b = (255*clip(x,0,1)).astype(uint8)
__device__ __forceinline__ unsigned to8b_one(float x) {
  // np.clip keeps NaN; NaN.astype(uint8) is 0 on x86 (cvttss2si gives INT_MIN, low byte 0): fmaxf(NaN, 0) = 0 matches
  const float c = fminf(fmaxf(x, 0.f), 1.f);
  return unsigned(__fmul_rn(255.f, c));      // fp32 product, truncation toward zero like astype
}

__global__ void to8b_kernel(const float* __restrict__ x, int64_t n, uint8_t* __restrict__ out) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;   // group of 4 values
  const int64_t i = q * 4;
  if (i + 3 < n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4 v = reinterpret_cast<const float4*>(x)[q];
    const unsigned w = to8b_one(v.x) | (to8b_one(v.y) << 8) | (to8b_one(v.z) << 16) | (to8b_one(v.w) << 24);
    reinterpret_cast<unsigned*>(out)[q] = w;
  } else {
    for (int64_t k = i; k < n && k < i + 4; ++k) out[k] = uint8_t(to8b_one(x[k]));
  }
}

int launch_to8b(const float* x, int64_t n, uint8_t* out, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const int64_t groups = (n + 3) / 4;
  to8b_kernel<<<unsigned((groups + 255) / 256), 256, 0, st>>>(x, n, out);
  count_launch();
  return check_launch("to8b_kernel");
}

// ----------------------------------------------------------------------------- get_rays from a device-resident c2w
__global__ void make_rays_dev_kernel(int H, int W, float fx, float fy, float cx, float cy, const float* __restrict__ c2w,
                                     int ld_c2w, float near_, float far_, float* __restrict__ rays) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const float i = float(idx % W), j = float(idx / W);  // RH:157-159: i = column (x), j = row (y)
  const float d0 = __fdiv_rn(__fsub_rn(i, cx), fx);
  const float d1 = -__fdiv_rn(__fsub_rn(j, cy), fy);
  const float d2 = -1.f;
  float rd[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)  // RH:162: sum(dirs[..., None, :] * c2w[:3,:3], -1)
    rd[a] = __fadd_rn(__fadd_rn(__fmul_rn(d0, c2w[a * ld_c2w + 0]), __fmul_rn(d1, c2w[a * ld_c2w + 1])), __fmul_rn(d2, c2w[a * ld_c2w + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
  float* o = rays + int64_t(idx) * 11;
  o[0] = c2w[3];
  o[1] = c2w[ld_c2w + 3];
  o[2] = c2w[2 * ld_c2w + 3];
  o[3] = rd[0];
  o[4] = rd[1];
  o[5] = rd[2];
  o[6] = near_;
  o[7] = far_;
  o[8] = __fdiv_rn(rd[0], nrm);  // RN:97
  o[9] = __fdiv_rn(rd[1], nrm);
  o[10] = __fdiv_rn(rd[2], nrm);
}

int launch_make_rays_dev(int H, int W, const float* K9, const float* c2w_dev, int ld_c2w, float near_, float far_, float* rays,
                         cudaStream_t st) {
  const int total = H * W;
  if (total == 0) return NSR_OK;
  make_rays_dev_kernel<<<(total + 255) / 256, 256, 0, st>>>(H, W, K9[0], K9[4], K9[2], K9[5], c2w_dev, ld_c2w, near_, far_, rays);
  count_launch();
  return check_launch("make_rays_dev_kernel");
}

// ----------------------------------------------------------------------------- dL/d(ray_batch) -> dL/dc2w
// rays_d = R dirs, rays_o = t, viewdir = rays_d / |rays_d|  (RH:160-164, RN:97), dirs = ((i-cx)/fx, -(j-cy)/fy, -1):
//   g_d   = dL/drays_d + (g_v - v (v . g_v)) / |rays_d|         (normalisation Jacobian, symmetric)
//   dL/dR[a][b] = sum_rays g_d[a] dirs[b],   dL/dt[a] = sum_rays dL/drays_o[a]
// Stage 1: C2W_GRAD_BLOCKS blocks, grid-stride over the rays, fp32 per-thread partials, fp64 from the warp level up;
// stage 2: one warp adds the block partials in a fixed order.  No atomics: bit-reproducible.
constexpr int C2W_GRAD_BLOCKS = 148;
constexpr int C2W_GRAD_THREADS = 256;

__global__ void __launch_bounds__(C2W_GRAD_THREADS)
    c2w_grad_partial_kernel(int W, float fx, float fy, float cx, float cy, const float* __restrict__ rays,
                            const float* __restrict__ d_rays, const int32_t* __restrict__ pixel_idx, int64_t n,
                            double* __restrict__ partials) {
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x) {
    const int pix = pixel_idx ? pixel_idx[r] : int(r);
    const float dirs[3] = {__fdiv_rn(__fsub_rn(float(pix % W), cx), fx), -__fdiv_rn(__fsub_rn(float(pix / W), cy), fy), -1.f};
    const float* ry = rays + r * 11;
    const float* g = d_rays + r * 11;
    const float d[3] = {ry[3], ry[4], ry[5]};
    const float inv = 1.f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float v[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
    const float vg = v[0] * g[8] + v[1] * g[9] + v[2] * g[10];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gd = g[3 + a] + (g[8 + a] - v[a] * vg) * inv;
#pragma unroll
      for (int b = 0; b < 3; ++b) acc[a * 4 + b] += gd * dirs[b];
      acc[a * 4 + 3] += g[a];
    }
  }
  __shared__ double sm[C2W_GRAD_THREADS / 32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    double v = double(acc[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double s = 0.0;
    for (int w = 0; w < C2W_GRAD_THREADS / 32; ++w) s += sm[w][threadIdx.x];
    partials[blockIdx.x * 12 + threadIdx.x] = s;
  }
}

__global__ void c2w_grad_final_kernel(const double* __restrict__ partials, int blocks, float* __restrict__ d_c2w, int accumulate) {
  const int k = threadIdx.x;
  if (k >= 12) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += partials[b * 12 + k];
  d_c2w[k] = accumulate ? d_c2w[k] + float(s) : float(s);
}

int launch_c2w_grad(int W, const float* K9, const float* rays, const float* d_rays, const int32_t* pixel_idx, int64_t n,
                    float* d_c2w, int accumulate, double* partials, cudaStream_t st) {
  c2w_grad_partial_kernel<<<C2W_GRAD_BLOCKS, C2W_GRAD_THREADS, 0, st>>>(W, K9[0], K9[4], K9[2], K9[5], rays, d_rays, pixel_idx, n, partials);
  count_launch();
  int rc = check_launch("c2w_grad_partial_kernel");
  if (rc) return rc;
  c2w_grad_final_kernel<<<1, 32, 0, st>>>(partials, C2W_GRAD_BLOCKS, d_c2w, accumulate);
  count_launch();
  return check_launch("c2w_grad_final_kernel");
}

size_t c2w_grad_workspace_bytes() { return size_t(C2W_GRAD_BLOCKS) * 12 * sizeof(double); }

}  // namespace nsr
