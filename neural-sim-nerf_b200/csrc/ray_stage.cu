// Per-ray stages of the renderer, one warp per ray, exact fp32:
//   coarse depths (RN:439-461), alpha compositing (RN:343-387 raw2outputs), inverse-CDF
//   resampling (RH:199-243 sample_pdf), sorted merge + z_std (RN:477, RN:495), ray generation
//   (RH:156-165 get_rays + RN:91-112 packing).
// These stages move ~20 B per sample against ~1.2 MFLOP per sample in the MLP, so they are
// written for exactness and coalescing, not for the last percent of bandwidth.
#include <math.h>

#include "common.cuh"

namespace nsr {

constexpr int WARPS_PER_BLOCK = 4;
constexpr unsigned FULL = 0xffffffffu;

// torch.linspace(0, 1, n)[i] as ATen computes it (symmetric about the midpoint)
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n == 1) return 0.f;
  const float step = 1.0f / float(n - 1);
  return (i < n / 2) ? __fmul_rn(step, float(i)) : __fsub_rn(1.0f, __fmul_rn(step, float(n - 1 - i)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ----------------------------------------------------------------------------- coarse depths, RN:439-461
__global__ void coarse_z_kernel(const float* __restrict__ rays, int64_t n, int S, uint32_t flags,
                                const float* __restrict__ t_rand, float* __restrict__ z) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= n * S) return;
  const int64_t r = idx / S;
  const int i = int(idx - r * S);
  const float nr = rays[r * 11 + 6], fr = rays[r * 11 + 7];
  auto zat = [&](int k) -> float {
    const float t = linspace01(k, S);
    if (!(flags & NSR_FLAG_LINDISP)) return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));  // RN:441
    return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, nr), __fsub_rn(1.f, t)), __fmul_rn(__fdiv_rn(1.f, fr), t)));  // RN:443
  };
  float zi = zat(i);
  if (t_rand != nullptr) {  // stratified jitter, RN:447-461
    const float lo = (i == 0) ? zi : __fmul_rn(.5f, __fadd_rn(zi, zat(i - 1)));
    const float up = (i == S - 1) ? zi : __fmul_rn(.5f, __fadd_rn(zat(i + 1), zi));
    zi = __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), t_rand[idx]));
  }
  z[idx] = zi;
}

// ----------------------------------------------------------------------------- raw2outputs, RN:343-387
// Lane l owns the C consecutive samples [l*C, l*C + C).
template <int C>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    raw2outputs_kernel(const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int ld,
                       int64_t n, int S, uint32_t flags, float* __restrict__ rgb_map, float* __restrict__ disp_map,
                       float* __restrict__ acc_map, float* __restrict__ weights, float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * int64_t(WARPS_PER_BLOCK) + (threadIdx.x >> 5);
  if (ray >= n) return;
  const float* zr = z + ray * S;
  const float4* rr = raw + ray * S;
  const float dx = rays_d[ray * ld], dy = rays_d[ray * ld + 1], dz = rays_d[ray * ld + 2];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));  // RN:361

  float alpha[C], cr[C], cg[C], cb[C], zz[C];
  float lane_prod = 1.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane * C + j;
    alpha[j] = 0.f;
    cr[j] = cg[j] = cb[j] = zz[j] = 0.f;
    if (i < S) {
      const float zi = zr[i];
      float dist = (i == S - 1) ? 1e10f : __fsub_rn(zr[i + 1], zi);  // RN:358-359
      dist = __fmul_rn(dist, dnorm);
      const float4 q = rr[i];
      const float sig = fmaxf(q.w, 0.f);
      alpha[j] = __fsub_rn(1.f, expf(-__fmul_rn(sig, dist)));  // RN:356
      cr[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.x)));      // sigmoid, RN:363
      cg[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.y)));
      cb[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.z)));
      zz[j] = zi;
      lane_prod *= __fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);  // RN:376
    }
  }
  // exclusive product scan across lanes
  float incl = lane_prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl *= v;
  }
  float T = __shfl_up_sync(FULL, incl, 1);
  if (lane == 0) T = 1.f;

  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane * C + j;
    if (i < S) {
      const float w = __fmul_rn(alpha[j], T);
      T *= __fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
      sr += w * cr[j];
      sg += w * cg[j];
      sb += w * cb[j];
      sd += w * zz[j];
      sa += w;
      if (weights) weights[ray * S + i] = w;
    }
  }
  sr = warp_sum(sr);
  sg = warp_sum(sg);
  sb = warp_sum(sb);
  sd = warp_sum(sd);
  sa = warp_sum(sa);
  if (lane == 0) {
    if (flags & NSR_FLAG_WHITE_BKGD) {  // RN:384-385
      sr += 1.f - sa;
      sg += 1.f - sa;
      sb += 1.f - sa;
    }
    if (rgb_map) {
      rgb_map[ray * 3 + 0] = sr;
      rgb_map[ray * 3 + 1] = sg;
      rgb_map[ray * 3 + 2] = sb;
    }
    if (depth_map) depth_map[ray] = sd;
    if (acc_map) acc_map[ray] = sa;
    if (disp_map) {
      const float q = __fdiv_rn(sd, sa);                 // 0/0 -> NaN when the ray hit nothing
      const float m = (q != q) ? q : fmaxf(1e-10f, q);   // torch.max propagates NaN (RN:381)
      disp_map[ray] = __fdiv_rn(1.f, m);
    }
  }
}

// ----------------------------------------------------------------------------- raw2outputs backward (for RN:177-178)
// Given g = dL/drgb_map: dL/draw [n,S,4] and dL/d||rays_d|| [n] (through dists = dz * ||d||, RN:361).
//   w_i = a_i T_i,  T_i = prod_{j<i} (1 - a_j + 1e-10),  a_i = 1 - exp(-relu(s_i) dist_i),  c_i = sigmoid(raw_rgb_i)
//   dL/dw_i = g . c_i (- sum(g) with white_bkgd);   dL/da_i = T_i dL/dw_i - (sum_{k>i} dL/dw_k w_k) / (1 - a_i + 1e-10)
//   dL/ds_i = dL/da_i dist_i (1 - a_i) [s_i > 0];   dL/ddist_i = dL/da_i relu(s_i) (1 - a_i);   dL/draw_rgb_i = w_i g c_i (1 - c_i)
template <int C>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    raw2outputs_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays, int64_t n, int S,
                           uint32_t flags, const float* __restrict__ d_rgb, float4* __restrict__ d_raw, float* __restrict__ d_dnorm,
                           float* __restrict__ gmax) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * int64_t(WARPS_PER_BLOCK) + (threadIdx.x >> 5);
  if (ray >= n) return;
  const float* zr = z + ray * S;
  const float4* rr = raw + ray * S;
  const float dx = rays[ray * 11 + 3], dy = rays[ray * 11 + 4], dz = rays[ray * 11 + 5];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float g0 = d_rgb[ray * 3], g1 = d_rgb[ray * 3 + 1], g2 = d_rgb[ray * 3 + 2];
  const float gw_bias = (flags & NSR_FLAG_WHITE_BKGD) ? -(g0 + g1 + g2) : 0.f;  // rgb += 1 - acc  (RN:385)

  float alpha[C], sig[C], dzs[C], cr[C], cg[C], cb[C];
  float lane_prod = 1.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int i = lane * C + j;
    alpha[j] = sig[j] = dzs[j] = cr[j] = cg[j] = cb[j] = 0.f;
    if (i < S) {
      const float zi = zr[i];
      dzs[j] = (i == S - 1) ? 1e10f : __fsub_rn(zr[i + 1], zi);
      const float4 q = rr[i];
      sig[j] = q.w;
      alpha[j] = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(q.w, 0.f), __fmul_rn(dzs[j], dnorm))));
      cr[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.x)));
      cg[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.y)));
      cb[j] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.z)));
      lane_prod *= __fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
    }
  }
  float incl = lane_prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl *= v;
  }
  float T = __shfl_up_sync(FULL, incl, 1);
  if (lane == 0) T = 1.f;
  // forward sweep inside the lane: T_i, w_i, dL/dw_i; lane total of dL/dw_k w_k
  float Ti[C], wi[C], dw[C];
  float lane_sum = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    Ti[j] = T;
    wi[j] = alpha[j] * T;
    dw[j] = g0 * cr[j] + g1 * cg[j] + g2 * cb[j] + gw_bias;
    lane_sum += dw[j] * wi[j];
    if (lane * C + j < S) T *= __fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
  }
  // exclusive SUFFIX sum over lanes
  float suf = lane_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_down_sync(FULL, suf, o);
    if (lane + o < 32) suf += v;
  }
  float after = __shfl_down_sync(FULL, suf, 1);  // sum over lanes > this one
  if (lane == 31) after = 0.f;
  float dn = 0.f, amax = 0.f;
#pragma unroll
  for (int j = C - 1; j >= 0; --j) {
    const int i = lane * C + j;
    if (i < S) {
      const float one_m = __fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
      const float dalpha = Ti[j] * dw[j] - after / one_m;
      const float e = 1.f - alpha[j];                       // exp(-relu(s) dist)
      const float dist = dzs[j] * dnorm;
      const float dsig = (sig[j] > 0.f) ? dalpha * dist * e : 0.f;
      const float ddist = dalpha * fmaxf(sig[j], 0.f) * e;
      dn += ddist * dzs[j];
      const float w = wi[j];
      const float4 o = make_float4(w * g0 * cr[j] * (1.f - cr[j]), w * g1 * cg[j] * (1.f - cg[j]), w * g2 * cb[j] * (1.f - cb[j]), dsig);
      d_raw[ray * S + i] = o;
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
      after += dw[j] * w;
    }
  }
  dn = warp_sum(dn);
  if (lane == 0) d_dnorm[ray] = dn;
  if (gmax != nullptr) {  // batch-wide max |dL/draw| (non-negative floats order like unsigned integers)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL, amax, o));
    if (lane == 0 && amax < 3.0e38f) atomicMax(reinterpret_cast<unsigned int*>(gmax), __float_as_uint(amax));
  }
}

// dL/d(ray_batch [n,11]) from the per-sample gradients: pts = o + d z (RN:463), dists = dz ||d|| (RN:361), viewdirs
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    ray_grad_reduce_kernel(const float* __restrict__ rays, const float* __restrict__ z, const float4* __restrict__ d_pts,
                           const float* __restrict__ d_dnorm, int64_t n, int S, float* __restrict__ d_rays) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * int64_t(WARPS_PER_BLOCK) + (threadIdx.x >> 5);
  if (ray >= n) return;
  float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f}, sv[3] = {0.f, 0.f, 0.f};
  for (int i = lane; i < S; i += 32) {
    const float4 p = d_pts[(ray * S + i) * 2], v = d_pts[(ray * S + i) * 2 + 1];
    const float zi = z[ray * S + i];
    so[0] += p.x; so[1] += p.y; so[2] += p.z;
    sd[0] += zi * p.x; sd[1] += zi * p.y; sd[2] += zi * p.z;
    sv[0] += v.x; sv[1] += v.y; sv[2] += v.z;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    so[d] = warp_sum(so[d]);
    sd[d] = warp_sum(sd[d]);
    sv[d] = warp_sum(sv[d]);
  }
  if (lane == 0) {
    const float dx = rays[ray * 11 + 3], dy = rays[ray * 11 + 4], dz = rays[ray * 11 + 5];
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float k = d_dnorm[ray] / nrm;
    float* o = d_rays + ray * 11;
    o[0] = so[0]; o[1] = so[1]; o[2] = so[2];
    o[3] = sd[0] + k * dx; o[4] = sd[1] + k * dy; o[5] = sd[2] + k * dz;
    o[6] = 0.f; o[7] = 0.f;   // near / far: the sampled depths are constants of the graph (python floats at RN:109)
    o[8] = sv[0]; o[9] = sv[1]; o[10] = sv[2];
  }
}

// torch.sum over the last (contiguous) dimension of a CPU fp32 tensor, bit for bit: ATen's
// vectorized_inner_sum reduces 8-lane vectors with four interleaved accumulators, then adds the scalar
// tail and the eight lane partials in order (aten/src/ATen/native/cpu/SumKernel.cpp; no cascade level
// is reached for n < 512).  x in shared memory; every lane returns the sum.
__device__ float aten_row_sum(const float* x, int n, int lane) {
  constexpr int V = 8;
  const int vs = n / V, size_ilp = vs / 4;
  float part = 0.f;
  if (lane < V) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    for (int i = 0; i < size_ilp; ++i) {
      p0 = __fadd_rn(p0, x[(4 * i + 0) * V + lane]);
      p1 = __fadd_rn(p1, x[(4 * i + 1) * V + lane]);
      p2 = __fadd_rn(p2, x[(4 * i + 2) * V + lane]);
      p3 = __fadd_rn(p3, x[(4 * i + 3) * V + lane]);
    }
    for (int i = size_ilp * 4; i < vs; ++i) p0 = __fadd_rn(p0, x[i * V + lane]);
    part = __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3);
  }
  float acc = 0.f;
  for (int k = vs * V; k < n; ++k) acc = __fadd_rn(acc, x[k]);
#pragma unroll
  for (int l = 0; l < V; ++l) acc = __fadd_rn(acc, __shfl_sync(FULL, part, l));
  return acc;
}

// ----------------------------------------------------------------------------- sample_pdf core, RH:199-243
// bins[B], w[B-1] in shared memory (w is overwritten with the pdf); cdf[B] scratch in shared memory.
// Writes N samples to out_s (shared).  u == nullptr -> deterministic linspace.
__device__ void sample_pdf_warp(const float* bins, float* w, float* cdf, int B, int N, const float* u, float* out_s, int lane) {
  const int nw = B - 1;
  for (int i = lane; i < nw; i += 32) w[i] = __fadd_rn(w[i], 1e-5f);  // RH:201
  __syncwarp();
  // The `denom < 1e-5` test below (RH:239) sits right on top of the value empty bins produce
  // (1e-5 / ~1.0006), and u = 1.0 sits on top of cdf[-1]; one ulp in the normaliser or in a prefix moves whole
  // samples by a bin width.  So normaliser and cdf are rounded exactly as ATen's CPU kernels round them:
  // torch.sum in its vector order (aten_row_sum), cumsum in fp64 with every prefix rounded to fp32.
  const float s = aten_row_sum(w, nw, lane);
  for (int i = lane; i < nw; i += 32) w[i] = __fdiv_rn(w[i], s);
  __syncwarp();
  // ATen's CPU cumsum accumulates sequentially in fp64 and rounds every prefix to fp32.  The pdf values are fp32 numbers in
  // [2^-23, 1] and a prefix stays below 2, so every fp64 partial sum is EXACT (24 + 23 + 1 significant bits at most): the sum is
  // associative here and a warp scan gives the sequential loop's prefixes bit for bit (stage test: tests/test_gpu_parity.py).
  if (lane == 0) cdf[0] = 0.f;
  double carry = 0.0;
  for (int i0 = 0; i0 < nw; i0 += 32) {
    const int i = i0 + lane;
    double v = i < nw ? double(w[i]) : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(FULL, v, d);
      if (lane >= d) v += o;
    }
    v += carry;
    if (i < nw) cdf[i + 1] = float(v);
    carry = __shfl_sync(FULL, v, 31);
  }
  __syncwarp();
  for (int k = lane; k < N; k += 32) {
    const float uk = u ? u[k] : linspace01(k, N);
    // inds = searchsorted(cdf, u, right=True) = #(cdf <= u)   (RH:227)
    int lo = 0, hi = B;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uk) lo = mid + 1; else hi = mid;
    }
    const int below = max(0, lo - 1), above = min(B - 1, lo);  // RH:228-229
    const float cb = cdf[below], ca = cdf[above];
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.f;  // RH:239
    const float t = __fdiv_rn(__fsub_rn(uk, cb), denom);
    out_s[k] = __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));  // RH:241
  }
  __syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, int64_t n, int B, int N,
                      const float* __restrict__ u, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ray = blockIdx.x * int64_t(WARPS_PER_BLOCK) + warp;
  if (ray >= n) return;
  float* base = sm + warp * (3 * B + N);
  float *sb = base, *sw = base + B, *sc = base + 2 * B, *so = base + 3 * B;
  for (int i = lane; i < B; i += 32) sb[i] = bins[ray * B + i];
  for (int i = lane; i < B - 1; i += 32) sw[i] = weights[ray * (B - 1) + i];
  __syncwarp();
  sample_pdf_warp(sb, sw, sc, B, N, u ? u + ray * N : nullptr, so, lane);
  for (int i = lane; i < N; i += 32) out[ray * N + i] = so[i];
}

// ----------------------------------------------------------------------------- RN:473-477 + RN:495
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    resample_merge_kernel(const float* __restrict__ z, const float* __restrict__ weights, int64_t n, int S, int Ni,
                          const float* __restrict__ u, float* __restrict__ z_fine, float* __restrict__ z_samples,
                          float* __restrict__ z_std, const uint32_t* __restrict__ ctrl_coarse, uint32_t* __restrict__ ctrl_fine,
                          uint32_t force_count) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // two-tier evaluation (common.cuh "active set"): a coarse pass that found most of its points non-empty (or failed its
  // verification) predicts a fine pass where tier 1 would certify next to nothing -- skip it there
  if (ctrl_fine != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
    ctrl_fine[AS_FORCE_DENSE] = (ctrl_coarse[AS_FORCE_DENSE] | ctrl_coarse[AS_DENSE_FINAL] | uint32_t(ctrl_coarse[AS_COUNT] > force_count)) ? 1u : 0u;
  const int64_t ray = blockIdx.x * int64_t(WARPS_PER_BLOCK) + warp;
  if (ray >= n) return;
  const int B = S - 1, T = S + Ni;
  float* base = sm + warp * (3 * S + T);
  float *sb = base, *sw = base + S, *sc = base + 2 * S, *all = base + 3 * S;  // all[0..S) = z, all[S..T) = samples
  for (int i = lane; i < S; i += 32) all[i] = z[ray * S + i];
  __syncwarp();
  for (int i = lane; i < B; i += 32) sb[i] = __fmul_rn(.5f, __fadd_rn(all[i + 1], all[i]));  // RN:473
  for (int i = lane; i < S - 2; i += 32) sw[i] = weights[ray * S + 1 + i];                    // weights[...,1:-1]
  __syncwarp();
  sample_pdf_warp(sb, sw, sc, B, Ni, u ? u + ray * Ni : nullptr, all + S, lane);
  // z_std = std(z_samples, unbiased=False)   RN:495
  float s = 0.f;
  for (int i = lane; i < Ni; i += 32) s += all[S + i];
  const float mean = warp_sum(s) / float(Ni);
  float v = 0.f;
  for (int i = lane; i < Ni; i += 32) {
    const float d = all[S + i] - mean;
    v += d * d;
    if (z_samples) z_samples[ray * Ni + i] = all[S + i];
  }
  v = warp_sum(v);
  if (lane == 0 && z_std) z_std[ray] = sqrtf(v / float(Ni));
  // sort(cat[z, z_samples])  RN:477.  Both lists are normally already sorted (the coarse depths always; the new
  // samples whenever u is ascending, i.e. det=True): merge by binary search -- rank = own index + number of elements of
  // the other list in front (ties: coarse first, so the ranks form a permutation).  Otherwise fall back to a rank sort.
  const float* zs = all + S;
  bool sorted = true;
  for (int i = lane; i < S - 1; i += 32) sorted = sorted && (all[i] <= all[i + 1]);
  for (int i = lane; i < Ni - 1; i += 32) sorted = sorted && (zs[i] <= zs[i + 1]);
  if (__all_sync(FULL, sorted)) {
    for (int i = lane; i < S; i += 32) {      // coarse element: samples strictly smaller go first
      const float e = all[i];
      int lo = 0, hi = Ni;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zs[mid] < e) lo = mid + 1; else hi = mid;
      }
      z_fine[ray * T + i + lo] = e;
    }
    for (int i = lane; i < Ni; i += 32) {     // new sample: coarse elements smaller or equal go first
      const float e = zs[i];
      int lo = 0, hi = S;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (all[mid] <= e) lo = mid + 1; else hi = mid;
      }
      z_fine[ray * T + i + lo] = e;
    }
  } else {
    // NaNs order last (as torch.sort places them), ties by index: the ranks always form a permutation
    for (int i = lane; i < T; i += 32) {
      const float e = all[i];
      const bool e_nan = e != e;
      int rank = 0;
      for (int j = 0; j < T; ++j) {
        const float o = all[j];
        const bool o_nan = o != o;
        rank += e_nan ? (!o_nan || j < i) : (!o_nan && ((o < e) || (o == e && j < i)));
      }
      z_fine[ray * T + rank] = e;
    }
  }
}

// ----------------------------------------------------------------------------- get_rays + packing
struct Cam {
  float K[9];
  float c2w[12];
};

__global__ void make_rays_kernel(int H, int W, Cam cam, float near_, float far_, float* __restrict__ rays) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const float i = float(idx % W), j = float(idx / W);  // RH:157-159: i = column (x), j = row (y)
  const float d0 = __fdiv_rn(__fsub_rn(i, cam.K[2]), cam.K[0]);
  const float d1 = -__fdiv_rn(__fsub_rn(j, cam.K[5]), cam.K[4]);
  const float d2 = -1.f;
  float rd[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)  // RH:162: sum(dirs[..., None, :] * c2w[:3,:3], -1)
    rd[a] = __fadd_rn(__fadd_rn(__fmul_rn(d0, cam.c2w[a * 4 + 0]), __fmul_rn(d1, cam.c2w[a * 4 + 1])), __fmul_rn(d2, cam.c2w[a * 4 + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])), __fmul_rn(rd[2], rd[2])));
  float* o = rays + int64_t(idx) * 11;
  o[0] = cam.c2w[3];
  o[1] = cam.c2w[7];
  o[2] = cam.c2w[11];
  o[3] = rd[0];
  o[4] = rd[1];
  o[5] = rd[2];
  o[6] = near_;
  o[7] = far_;
  o[8] = __fdiv_rn(rd[0], nrm);  // RN:97
  o[9] = __fdiv_rn(rd[1], nrm);
  o[10] = __fdiv_rn(rd[2], nrm);
}

// RN:91-112 for rays that were generated elsewhere (render(rays=...), the route of RN:163-170): (rays_o, rays_d) [n,3] each ->
// [n,11] = o, d, near, far, d / |d|.  Same arithmetic as make_rays_kernel's tail.
__global__ void pack_rays_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t n, float near_, float far_,
                                 float* __restrict__ rays) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float d0 = d[idx * 3], d1 = d[idx * 3 + 1], d2 = d[idx * 3 + 2];
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
  float* r = rays + idx * 11;
  r[0] = o[idx * 3];
  r[1] = o[idx * 3 + 1];
  r[2] = o[idx * 3 + 2];
  r[3] = d0;
  r[4] = d1;
  r[5] = d2;
  r[6] = near_;
  r[7] = far_;
  r[8] = __fdiv_rn(d0, nrm);  // RN:97
  r[9] = __fdiv_rn(d1, nrm);
  r[10] = __fdiv_rn(d2, nrm);
}

int launch_pack_rays(const float* o, const float* d, int64_t n, float near_, float far_, float* rays, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  pack_rays_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(o, d, n, near_, far_, rays);
  count_launch();
  return check_launch("pack_rays_kernel");
}

// ----------------------------------------------------------------------------- launchers
int launch_coarse_z(const float* rays, int64_t n, int S, uint32_t flags, const float* t_rand, float* z, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const int64_t total = n * S;
  coarse_z_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(rays, n, S, flags, t_rand, z);
  count_launch();
  return check_launch("coarse_z_kernel");
}

int launch_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld, int64_t n, int S, uint32_t flags,
                       float* rgb, float* disp, float* acc, float* weights, float* depth, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const int C = (S + 31) / 32;
  if (C > 8) {
    set_error("raw2outputs: n_samples=%d > 256 not supported", S);
    return NSR_E_UNSUPPORTED;
  }
  const unsigned grid = unsigned((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  const float4* r4 = reinterpret_cast<const float4*>(raw);
#define NSR_R2O(CC)                                                                                                   \
  case CC:                                                                                                            \
    raw2outputs_kernel<CC><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(r4, z, rays_d, ld, n, S, flags, rgb, disp, acc, weights, depth); \
    break;
  switch (C) {
    NSR_R2O(1) NSR_R2O(2) NSR_R2O(3) NSR_R2O(4) NSR_R2O(5) NSR_R2O(6) NSR_R2O(7) NSR_R2O(8)
  }
#undef NSR_R2O
  count_launch();
  return check_launch("raw2outputs_kernel");
}

int launch_sample_pdf(const float* bins, const float* weights, int64_t n, int B, int N, const float* u, float* out, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const size_t smem = size_t(WARPS_PER_BLOCK) * (3 * B + N) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("sample_pdf: n_bins=%d / n_new=%d too large", B, N);
    return NSR_E_UNSUPPORTED;
  }
  const unsigned grid = unsigned((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  sample_pdf_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, st>>>(bins, weights, n, B, N, u, out);
  count_launch();
  return check_launch("sample_pdf_kernel");
}

int launch_resample_merge(const float* z, const float* w, int64_t n, int S, int Ni, const float* u, float* z_fine,
                          float* z_samples, float* z_std, cudaStream_t st, const uint32_t* ctrl_coarse, uint32_t* ctrl_fine,
                          uint32_t force_count) {
  if (n == 0) return NSR_OK;
  const size_t smem = size_t(WARPS_PER_BLOCK) * (3 * S + S + Ni) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("resample_merge: n_samples=%d / n_importance=%d too large", S, Ni);
    return NSR_E_UNSUPPORTED;
  }
  const unsigned grid = unsigned((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  resample_merge_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, st>>>(z, w, n, S, Ni, u, z_fine, z_samples, z_std, ctrl_coarse, ctrl_fine, force_count);
  count_launch();
  return check_launch("resample_merge_kernel");
}

int launch_raw2outputs_backward(const float* raw, const float* z, const float* rays, int64_t n, int S, uint32_t flags,
                                const float* d_rgb, float* d_raw, float* d_dnorm, float* gmax, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const int C = (S + 31) / 32;
  if (C > 8) {
    set_error("raw2outputs backward: n_samples=%d > 256 not supported", S);
    return NSR_E_UNSUPPORTED;
  }
  const unsigned grid = unsigned((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  const float4* r4 = reinterpret_cast<const float4*>(raw);
  float4* d4 = reinterpret_cast<float4*>(d_raw);
#define NSR_R2OB(CC)                                                                                              \
  case CC:                                                                                                        \
    raw2outputs_bwd_kernel<CC><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(r4, z, rays, n, S, flags, d_rgb, d4, d_dnorm, gmax); \
    break;
  switch (C) {
    NSR_R2OB(1) NSR_R2OB(2) NSR_R2OB(3) NSR_R2OB(4) NSR_R2OB(5) NSR_R2OB(6) NSR_R2OB(7) NSR_R2OB(8)
  }
#undef NSR_R2OB
  count_launch();
  return check_launch("raw2outputs_bwd_kernel");
}

int launch_ray_grad_reduce(const float* rays, const float* z, const float* d_pts, const float* d_dnorm, int64_t n, int S,
                           float* d_rays, cudaStream_t st) {
  if (n == 0) return NSR_OK;
  const unsigned grid = unsigned((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  ray_grad_reduce_kernel<<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(rays, z, reinterpret_cast<const float4*>(d_pts), d_dnorm, n, S, d_rays);
  count_launch();
  return check_launch("ray_grad_reduce_kernel");
}

int launch_make_rays(int H, int W, const float* K9, const float* c2w12, float near_, float far_, float* rays, cudaStream_t st) {
  Cam cam;
  for (int i = 0; i < 9; ++i) cam.K[i] = K9[i];
  for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w12[i];
  const int total = H * W;
  if (total == 0) return NSR_OK;
  make_rays_kernel<<<(total + 255) / 256, 256, 0, st>>>(H, W, cam, near_, far_, rays);
  count_launch();
  return check_launch("make_rays_kernel");
}

}  // namespace nsr
