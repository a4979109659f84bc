// Coarse-pass refinement: the density of the few coarse sample points that hierarchical sampling is ill-conditioned in, re-evaluated
// with fp64 accumulation on the CUDA cores.
//
// Why.  sample_pdf (RH:199-243) places the fine samples by the coarse weights normalised over the ray: pdf_i = (w_i + 1e-5) / sum.
// On a ray that only grazes the object the sum is tiny (acc0 ~ 1e-3) and carried by one or two samples whose density is barely
// positive (sigma ~ 0.02): the tensor-core arithmetic's ABSOLUTE error on sigma (~1e-4..3e-4: fp16 hi/lo operands, tensor-core
// accumulation) is then a per-cent error of that weight, the pdf shifts, most of the 128 fine samples move by a fraction of a bin, and
// at a silhouette the pixel moves by up to 4e-2 -- 4 of the 160 000 rays of the test image were outside the 1e-3 bar that way
// (tools/parity_full_image.py, tools/parity_outliers.py), while the reference on the CPU and the reference on CUDA agree with each
// other to 3e-4 on the same rays.  Empty rays are immune (all weights 0); opaque ones nearly so (the sum is ~1: what is left there is a
// fine sample in the low-density bin in front of the surface moving by ~1e-5, 1.1e-3 on one ray of the test image).
//
// What.  After the coarse network pass: (1) select_refine_kernel, one warp per ray, sums the ray's optical depth from raw0; on rays
// that are not opaque (optical depth below the limit, default 2.3 i.e. acc0 < 0.9) every sample whose sigma is not clearly negative goes on a
// list; (2) refine_sigma_kernel evaluates pts_linears.0-7 + the alpha head for the listed points (fp64-accumulated sums over K, fp32 layer outputs, accurate
// sincosf encoding, the fp32 weights kept TRANSPOSED behind the packed tail: common.cuh REF_*), eight points per 256-thread block
// pass, one output unit per thread, and overwrites raw0[p].sigma.  ~12 000 points per 400x400 image: 0.75 ms next to 54 ms.
#include <math.h>

#include "common.cuh"

namespace nsr {

static float g_refine_tau_limit = 2.3026f;     // optical depth of acc0 = 0.9 (nsr_set_coarse_refine_limit)
void set_refine_tau_limit(float tau) { g_refine_tau_limit = tau; }
constexpr float REFINE_SIGMA_MIN = -0.01f;    // samples with sigma above this on such a ray are re-evaluated
static float g_refine_sigma_hi = 10.f;        // and, on EVERY ray, samples with sigma in (SIGMA_MIN, this): the surface-entry samples (0: none)
void set_refine_sigma_hi(float s) { g_refine_sigma_hi = s; }
constexpr int REFINE_POINTS = 8;              // points per block pass: the 1.97 MB of fp32 weights are re-read from L2 once per pass
constexpr unsigned FULLMASK = 0xffffffffu;

// workspace: [count u32, padded to 256 B][list: int32 x cap]
static inline int64_t refine_cap(int64_t n_rays) { return 2 * n_rays + 1024; }
size_t refine_workspace_bytes(int64_t n_rays) { return 256 + size_t(refine_cap(n_rays)) * 4; }

__global__ void __launch_bounds__(256) select_refine_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays,
                                                            int64_t n, int S, float tau_limit, float sigma_hi, uint32_t* __restrict__ count, int32_t* __restrict__ list,
                                                            uint32_t cap) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (ray >= n) return;
  const float* rp = rays + ray * 11;
  const float nrm = sqrtf(rp[3] * rp[3] + rp[4] * rp[4] + rp[5] * rp[5]);
  float tau = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float sg = raw[(ray * S + i) * 4 + 3];
    const float dist = (i + 1 < S ? z[ray * S + i + 1] - z[ray * S + i] : 1e10f) * nrm;   // RN:358-361
    tau += fmaxf(sg, 0.f) * dist;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) tau += __shfl_xor_sync(FULLMASK, tau, d);
  // opaque (or NaN) rays: the normaliser is ~1 and a 1e-4 error of a saturated sample's sigma moves nothing; only their
  // low-density samples (the surface entry: alpha far from 1, so the absolute error of sigma passes into the weight undamped) count
  const bool translucent = tau < tau_limit;
  if (!translucent && !(sigma_hi > REFINE_SIGMA_MIN)) return;
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    const float sg = i < S ? raw[(ray * S + i) * 4 + 3] : -1.f;
    const bool pick = i < S && sg > REFINE_SIGMA_MIN && (translucent || sg < sigma_hi);
    const uint32_t m = __ballot_sync(FULLMASK, pick);
    if (m == 0u) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, uint32_t(__popc(m)));
    base = __shfl_sync(FULLMASK, base, 0);
    const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
    if (pick && pos < cap) list[pos] = int32_t(ray * S + i);
  }
}

__global__ void __launch_bounds__(256) refine_sigma_kernel(const int32_t* __restrict__ list, const uint32_t* __restrict__ count, uint32_t cap,
                                                           const float* __restrict__ rays, const float* __restrict__ z, int S,
                                                           const uint8_t* __restrict__ packed, float* __restrict__ raw) {
  constexpr int R = REFINE_POINTS;
  // activations are kept [k][point]: a thread needs h[k] of all R points for every k, i.e. R consecutive floats = four broadcast
  // LDS.128 instead of sixteen LDS.32 (the scalar version was bound by shared-memory load instructions, not by the fp64 pipe)
  __shared__ __align__(16) float enc[64][R];
  __shared__ __align__(16) float hbuf[2][256][R];
  const float* W32 = reinterpret_cast<const float*>(packed + REF_OFF);
  const float* tail = reinterpret_cast<const float*>(packed + WEIGHT_BYTES);
  const int j = threadIdx.x, warp = j >> 5, lane = j & 31;
  // more candidates than the list holds (a fog-like scene: every ray translucent): WHICH of them made it onto the list depends on the
  // order of the atomics, so refining a subset would make the render irreproducible -- refine none
  if (*count > cap) return;
  const uint32_t n = *count;
  for (uint32_t base = blockIdx.x * R; base < n; base += gridDim.x * R) {
    __syncthreads();   // the previous pass is done with the shared buffers
    // ---- gamma(x) (RH:47-48) of the R points: 64 channels each (63 + a zero)
    for (int t = j; t < R * 64; t += 256) {
      const int r = t >> 6, c = t & 63;
      float v = 0.f;
      if (base + r < n && c < 63) {
        const int64_t p = list[base + r];
        const float* rp = rays + (p / S) * 11;
        const int d = c < 3 ? c : (c - 3) % 3;
        const float x = __fadd_rn(rp[d], __fmul_rn(rp[3 + d], z[p]));   // RN:463
        if (c < 3) {
          v = x;
        } else {
          const int k = (c - 3) / 6;
          const float a = x * float(1 << k);
          v = ((c - 3) % 6) < 3 ? sinf(a) : cosf(a);
        }
      }
      enc[c][r] = v;
    }
    __syncthreads();
    // ---- pts_linears.0-7: thread j = output unit j of all R points, sums accumulated in fp64 (see below), one rounding to fp32 per
    // unit as in the reference: what separates this evaluation from the reference's is mostly the reference's own fp32 rounding
    int cur = 0;
#pragma unroll 1
    for (int l = 0; l < 8; ++l) {
      const float* W = W32 + ref_layer_off(l);
      // (vector fp64 runs at ~1/10 of the fp32 rate on this part -- measured: an all-fp64 version of this loop took 1.5 ms for 9 000
      // points -- so the products of 8 consecutive k are summed by fp32 FMAs and only the block sums are accumulated in fp64: a block
      // of 8 carries ~1e-7 of its own size, and the long-range accumulation, where an fp32 chain loses its bits, is exact)
      double acc[R];
      float part[R];
      const double b = double(tail[TAIL_BIAS + l * 256 + j]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] = b;
        part[r] = 0.f;
      }
      auto step = [&](const float* hk, float w) {   // hk: the R values of input k
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
          const float4 h4 = *reinterpret_cast<const float4*>(hk + 4 * q);
          part[4 * q] = fmaf(h4.x, w, part[4 * q]);
          part[4 * q + 1] = fmaf(h4.y, w, part[4 * q + 1]);
          part[4 * q + 2] = fmaf(h4.z, w, part[4 * q + 2]);
          part[4 * q + 3] = fmaf(h4.w, w, part[4 * q + 3]);
        }
      };
      auto flush = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r] += double(part[r]);
          part[r] = 0.f;
        }
      };
      if (l == 0 || l == 5) {
#pragma unroll 1
        for (int k0 = 0; k0 < 64; k0 += 8) {
#pragma unroll
          for (int k = k0; k < k0 + 8; ++k)
            if (k < 63) step(enc[k], W[k * 256 + j]);
          flush();
        }
        W += 63 * 256;
      }
      if (l != 0) {
#pragma unroll 1
        for (int k0 = 0; k0 < 256; k0 += 8) {
#pragma unroll
          for (int k = k0; k < k0 + 8; ++k) step(hbuf[cur][k], W[k * 256 + j]);
          flush();
        }
      }
      const int nxt = l == 0 ? cur : cur ^ 1;
#pragma unroll
      for (int q = 0; q < R / 4; ++q)
        *reinterpret_cast<float4*>(&hbuf[nxt][j][4 * q]) = make_float4(fmaxf(float(acc[4 * q]), 0.f), fmaxf(float(acc[4 * q + 1]), 0.f),
                                                                       fmaxf(float(acc[4 * q + 2]), 0.f), fmaxf(float(acc[4 * q + 3]), 0.f));
      cur = nxt;
      __syncthreads();
    }
    // ---- alpha head (RH:109): warp w reduces points w, w + 8
    for (int r = warp; r < R; r += 8) {
      const float* wa = W32 + ref_layer_off(8);
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fma(double(hbuf[cur][lane + 32 * q][r]), double(wa[lane + 32 * q]), s);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULLMASK, s, d);
      if (lane == 0 && base + r < n) raw[int64_t(list[base + r]) * 4 + 3] = float(s + double(tail[TAIL_MISC]));
    }
  }
}

int launch_coarse_refine(const float* rays, const float* z, int64_t n, int S, const void* packed, float* raw, void* workspace, cudaStream_t st) {
  if (n == 0 || n * int64_t(S) >= (int64_t(1) << 31)) return NSR_OK;   // (the list holds int32 point indices)
  uint32_t* count = static_cast<uint32_t*>(workspace);
  int32_t* list = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(workspace) + 256);
  const uint32_t cap = uint32_t(refine_cap(n));
  if (cudaMemsetAsync(count, 0, 4, st) != cudaSuccess) return check_launch("refine count init");
  select_refine_kernel<<<unsigned((n * 32 + 255) / 256), 256, 0, st>>>(raw, z, rays, n, S, g_refine_tau_limit, g_refine_sigma_hi, count, list, cap);
  count_launch();
  if (int rc = check_launch("select_refine_kernel")) return rc;
  int sms = 0;
  if (int rc = current_device_sms(&sms)) return rc;
  refine_sigma_kernel<<<4 * sms, 256, 0, st>>>(list, count, cap, rays, z, S, static_cast<const uint8_t*>(packed), raw);
  count_launch();
  return check_launch("refine_sigma_kernel");
}

}  // namespace nsr
