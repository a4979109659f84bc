// C ABI of libnsr_b200 (include/nsr_b200.h): argument checking, error strings, and the
// forward orchestration of render_rays (RN:390-501) over the kernels in ray_stage.cu / mlp_forward.cu.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace nsr {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return NSR_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return NSR_E_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// SM count of the CURRENT device (must be sm_100), cached per device ordinal
int current_device_sms(int* sms) {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return check_launch("cudaGetDevice");
  if (dev >= 0 && dev < 64 && cache[dev] > 0) {
    *sms = cache[dev];
    return NSR_OK;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return check_launch("cudaGetDeviceProperties");
  if (prop.major != 10) {
    set_error("libnsr_b200 needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
    return NSR_E_DEVICE;
  }
  if (dev >= 0 && dev < 64) cache[dev] = prop.multiProcessorCount;
  *sms = prop.multiProcessorCount;
  return NSR_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: apply it once per (kernel, device)
int ensure_dynamic_smem(const void* func, int bytes) {
  struct Entry { const void* f; int dev; };
  static Entry done[256];
  static int n_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < n_done; ++i)
    if (done[i].f == func && done[i].dev == dev) return NSR_OK;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
    return check_launch("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  if (n_done < 256) done[n_done++] = Entry{func, dev};
  return NSR_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace nsr

using namespace nsr;

#define NSR_REQUIRE(cond, ...)   \
  do {                           \
    if (!(cond)) {               \
      set_error(__VA_ARGS__);    \
      return NSR_E_INVALID;      \
    }                            \
  } while (0)

extern "C" {

int nsr_version(void) { return 100; }

const char* nsr_last_error(void) { return g_err; }

uint64_t nsr_launch_count(void) { return g_launches.load(); }

size_t nsr_packed_net_bytes(void) { return PACKED_BYTES; }

int nsr_pack_net(const float* const* weights, const float* const* biases, void* packed_out, void* stream) {
  NSR_REQUIRE(weights && biases && packed_out, "nsr_pack_net: null argument");
  for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) NSR_REQUIRE(weights[i] && biases[i], "nsr_pack_net: tensor %d is null", i);
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(packed_out) & 127) == 0, "nsr_pack_net: packed_out must be 128-byte aligned");
  return launch_pack_net(weights, biases, packed_out, static_cast<cudaStream_t>(stream));
}

int nsr_mlp_forward(const float* rays, const float* z_or_pts, int64_t n_rays, int n_samples, const void* packed_net,
                    uint32_t flags, float* raw_out, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_samples > 0, "nsr_mlp_forward: bad sizes n_rays=%lld n_samples=%d", (long long)n_rays, n_samples);
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(rays && z_or_pts && packed_net && raw_out, "nsr_mlp_forward: null argument");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(raw_out) & 15) == 0, "nsr_mlp_forward: raw_out must be 16-byte aligned");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(packed_net) & 127) == 0, "nsr_mlp_forward: packed_net must be 128-byte aligned");
  return launch_mlp_forward(rays, z_or_pts, n_rays, n_samples, packed_net, flags, raw_out, static_cast<cudaStream_t>(stream));
}

int nsr_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int ld_rays_d, int64_t n_rays,
                    int n_samples, uint32_t flags, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                    float* depth_map, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_samples > 0 && ld_rays_d >= 3, "nsr_raw2outputs: bad sizes");
  // with one sample the reference's dists tensor is EMPTY (RN:358-359 expands to dists[..., :1].shape == [n, 0]) and every
  // output collapses to a sum over nothing; that degenerate case is not reproduced
  NSR_REQUIRE(n_samples >= 2, "nsr_raw2outputs: needs at least 2 samples per ray");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(raw && z_vals && rays_d, "nsr_raw2outputs: null input");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "nsr_raw2outputs: raw must be 16-byte aligned");
  return launch_raw2outputs(raw, z_vals, rays_d, ld_rays_d, n_rays, n_samples, flags, rgb_map, disp_map, acc_map, weights,
                            depth_map, static_cast<cudaStream_t>(stream));
}

int nsr_sample_pdf(const float* bins, const float* weights, int64_t n_rays, int n_bins, int n_new, const float* u,
                   float* samples_out, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_bins >= 2 && n_new > 0, "nsr_sample_pdf: bad sizes");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(bins && weights && samples_out, "nsr_sample_pdf: null argument");
  return launch_sample_pdf(bins, weights, n_rays, n_bins, n_new, u, samples_out, static_cast<cudaStream_t>(stream));
}

int nsr_resample_merge(const float* z_coarse, const float* weights, int64_t n_rays, int n_samples, int n_importance,
                       const float* u, float* z_fine, float* z_samples, float* z_std, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_samples >= 3 && n_importance > 0, "nsr_resample_merge: bad sizes");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(z_coarse && weights && z_fine, "nsr_resample_merge: null argument");
  return launch_resample_merge(z_coarse, weights, n_rays, n_samples, n_importance, u, z_fine, z_samples, z_std,
                               static_cast<cudaStream_t>(stream));
}

// workspace layout: z0 [n,S] | w0 [n,S] | raw0 [n,S,4] | z1 [n,T] | raw1 [n,T,4]   (T = S + Ni)
size_t nsr_render_workspace_bytes(int64_t n, int S, int Ni) {
  const size_t T = size_t(S) + size_t(Ni);
  size_t b = 0;
  b += align_up(size_t(n) * S * 4, 256);
  b += align_up(size_t(n) * S * 4, 256);
  b += align_up(size_t(n) * S * 16, 256);
  b += align_up(size_t(n) * T * 4, 256);
  b += align_up(size_t(n) * T * 16, 256);
  return b;
}

int nsr_render_rays_forward(const float* rays, int64_t n, const void* packed_coarse, const void* packed_fine, int S,
                            int Ni, uint32_t flags, const float* t_rand, const float* u, float* rgb_map, float* disp_map,
                            float* acc_map, float* rgb0, float* disp0, float* acc0, float* z_std, float* raw,
                            float* z_vals_out, float* weights_out, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(n >= 0 && S >= 2 && Ni >= 0, "nsr_render_rays_forward: bad sizes (needs at least 2 samples per ray)");
  if (n == 0) return NSR_OK;
  NSR_REQUIRE(rays && packed_coarse, "nsr_render_rays_forward: null rays / weights");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_render_workspace_bytes(n, S, Ni), "nsr_render_rays_forward: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "nsr_render_rays_forward: workspace must be 256-byte aligned");
  NSR_REQUIRE(Ni == 0 || S >= 3, "nsr_render_rays_forward: hierarchical sampling needs n_samples >= 3");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = S + Ni;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* z0 = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * S * 4, 256);
  float* w0 = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * S * 4, 256);
  float* raw0 = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * S * 16, 256);
  float* z1 = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * T * 4, 256);
  float* raw1 = reinterpret_cast<float*>(ws);
  const uint32_t cflags = flags & NSR_FLAG_WHITE_BKGD;
  const uint32_t mflags = flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8);
  int rc;

  if ((rc = launch_coarse_z(rays, n, S, flags, t_rand, z0, st))) return rc;                          // RN:439-461
  if ((rc = launch_mlp_forward(rays, z0, n, S, packed_coarse, mflags, raw0, st))) return rc;              // RN:463-466
  if (Ni == 0) {
    if ((rc = launch_raw2outputs(raw0, z0, rays + 3, 11, n, S, cflags, rgb_map, disp_map, acc_map, weights_out, nullptr, st))) return rc;
    if (raw) cudaMemcpyAsync(raw, raw0, size_t(n) * S * 16, cudaMemcpyDeviceToDevice, st);
    if (z_vals_out) cudaMemcpyAsync(z_vals_out, z0, size_t(n) * S * 4, cudaMemcpyDeviceToDevice, st);
    return check_launch("render_rays_forward(coarse only)");
  }
  if ((rc = launch_raw2outputs(raw0, z0, rays + 3, 11, n, S, cflags, rgb0, disp0, acc0, w0, nullptr, st))) return rc;  // RN:467
  float* zf = z_vals_out ? z_vals_out : z1;
  if ((rc = launch_resample_merge(z0, w0, n, S, Ni, u, zf, nullptr, z_std, st))) return rc;          // RN:473-477, 495
  float* rawf = raw ? raw : raw1;
  if ((rc = launch_mlp_forward(rays, zf, n, T, packed_fine ? packed_fine : packed_coarse, mflags, rawf, st))) return rc;  // RN:478-483
  if ((rc = launch_raw2outputs(rawf, zf, rays + 3, 11, n, T, cflags, rgb_map, disp_map, acc_map, weights_out, nullptr, st))) return rc;  // RN:485
  return NSR_OK;
}

// workspace layout: d_raw [n,T,4] | d_pts [n,T,8] | d_dnorm [n] | gmax (1 float)
size_t nsr_render_backward_workspace_bytes(int64_t n, int T) {
  return align_up(size_t(n) * T * 16, 256) + align_up(size_t(n) * T * 32, 256) + align_up(size_t(n) * 4, 256) + 256;
}

size_t nsr_mlp_dump_bytes(int64_t n_rays, int n_total_samples) { return mlp_dump_bytes(n_rays * n_total_samples); }

int nsr_render_rays_backward(const float* rays, const float* z_vals, const float* raw, int64_t n, int T, const void* packed_net,
                             uint32_t flags, const float* d_rgb_map, float* d_rays, void* dump, float* const* dW,
                             float* const* dB, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(n >= 0 && T > 0, "nsr_render_rays_backward: bad sizes");
  if (n == 0) return NSR_OK;
  NSR_REQUIRE(rays && z_vals && raw && packed_net && d_rgb_map && d_rays, "nsr_render_rays_backward: null argument");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_render_backward_workspace_bytes(n, T), "nsr_render_rays_backward: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0,
              "nsr_render_rays_backward: workspace must be 256-byte and raw 16-byte aligned");
  NSR_REQUIRE(!(flags & NSR_FLAG_FAST_FP16), "nsr_render_rays_backward: only the default (fp16 hi/lo split) precision is built");
  NSR_REQUIRE((dW == nullptr) == (dB == nullptr), "nsr_render_rays_backward: dW and dB go together");
  NSR_REQUIRE(dW == nullptr || dump != nullptr, "nsr_render_rays_backward: parameter gradients need the dump scratch buffer");
  if (dW)
    for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) NSR_REQUIRE(dW[i] && dB[i], "nsr_render_rays_backward: gradient tensor %d is null", i);
  NSR_REQUIRE(dump == nullptr || (reinterpret_cast<uintptr_t>(dump) & 127) == 0, "nsr_render_rays_backward: dump must be 128-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* d_raw = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * T * 16, 256);
  float* d_pts = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * T * 32, 256);
  float* d_dnorm = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * 4, 256);
  float* gmax = reinterpret_cast<float*>(ws);
  int rc;
  if (dW) cudaMemsetAsync(gmax, 0, 4, st);
  if ((rc = launch_raw2outputs_backward(raw, z_vals, rays, n, T, flags & NSR_FLAG_WHITE_BKGD, d_rgb_map, d_raw, d_dnorm,
                                        dW ? gmax : nullptr, st))) return rc;
  if ((rc = launch_mlp_backward(rays, z_vals, n, T, packed_net, d_raw, d_pts, dW ? dump : nullptr, gmax, st))) return rc;
  if ((rc = launch_ray_grad_reduce(rays, z_vals, d_pts, d_dnorm, n, T, d_rays, st))) return rc;
  if (dW) return launch_weight_grads(dump, d_raw, n * int64_t(T), gmax, dW, dB, st);
  return NSR_OK;
}

int nsr_make_rays(int H, int W, const float* K_host, const float* c2w_host, float near_, float far_, float* rays_out, void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host && c2w_host && rays_out, "nsr_make_rays: bad argument");
  return launch_make_rays(H, W, K_host, c2w_host, near_, far_, rays_out, static_cast<cudaStream_t>(stream));
}

int nsr_make_rays_dev(int H, int W, const float* K_host, const float* c2w_dev, int ld_c2w, float near_, float far_, float* rays_out,
                      void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host && c2w_dev && rays_out && ld_c2w >= 4, "nsr_make_rays_dev: bad argument");
  return launch_make_rays_dev(H, W, K_host, c2w_dev, ld_c2w, near_, far_, rays_out, static_cast<cudaStream_t>(stream));
}

int nsr_to8b(const float* x, int64_t n_values, uint8_t* out, void* stream) {
  NSR_REQUIRE(n_values >= 0, "nsr_to8b: bad size");
  if (n_values == 0) return NSR_OK;
  NSR_REQUIRE(x && out, "nsr_to8b: null argument");
  return launch_to8b(x, n_values, out, static_cast<cudaStream_t>(stream));
}

size_t nsr_c2w_grad_workspace_bytes(void) { return c2w_grad_workspace_bytes(); }

int nsr_rays_grad_to_c2w(int H, int W, const float* K_host, const float* rays, const float* d_rays, const int32_t* pixel_idx,
                         int64_t n_rays, float* d_c2w, int accumulate, void* workspace, void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host && d_c2w && workspace && n_rays >= 0, "nsr_rays_grad_to_c2w: bad argument");
  NSR_REQUIRE(n_rays == 0 || (rays && d_rays), "nsr_rays_grad_to_c2w: null rays / d_rays");
  NSR_REQUIRE(pixel_idx != nullptr || n_rays == int64_t(H) * W, "nsr_rays_grad_to_c2w: without pixel_idx the rays must be the whole image in row-major order");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "nsr_rays_grad_to_c2w: workspace must be 8-byte aligned");
  return launch_c2w_grad(W, K_host, rays, d_rays, pixel_idx, n_rays, d_c2w, accumulate, static_cast<double*>(workspace),
                         static_cast<cudaStream_t>(stream));
}

// workspace layout: rays [H*W,11] | rgb [H*W,3] (used when rgb8 is wanted without rgb_map) | nsr_render_rays_forward's workspace
size_t nsr_render_image_workspace_bytes(int H, int W, int S, int Ni) {
  const int64_t n = int64_t(H) * W;
  return align_up(size_t(n) * 44, 256) + align_up(size_t(n) * 12, 256) + nsr_render_workspace_bytes(n, S, Ni);
}

int nsr_render_image_forward(int H, int W, const float* K_host, const float* c2w_host, const float* c2w_dev, int ld_c2w, float near_,
                             float far_, const void* packed_coarse, const void* packed_fine, int S, int Ni, uint32_t flags,
                             uint8_t* rgb8, float* rgb_map, float* disp_map, float* acc_map, float* rgb0, float* disp0, float* acc0,
                             float* z_std, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host, "nsr_render_image_forward: bad camera");
  NSR_REQUIRE((c2w_host != nullptr) != (c2w_dev != nullptr), "nsr_render_image_forward: give exactly one of c2w_host / c2w_dev");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_render_image_workspace_bytes(H, W, S, Ni), "nsr_render_image_forward: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "nsr_render_image_forward: workspace must be 256-byte aligned");
  const int64_t n = int64_t(H) * W;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* rays = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * 44, 256);
  float* rgb_tmp = reinterpret_cast<float*>(ws);
  ws += align_up(size_t(n) * 12, 256);
  int rc;
  if (c2w_host) rc = nsr_make_rays(H, W, K_host, c2w_host, near_, far_, rays, stream);
  else rc = nsr_make_rays_dev(H, W, K_host, c2w_dev, ld_c2w, near_, far_, rays, stream);
  if (rc) return rc;
  float* rgb = rgb_map ? rgb_map : (rgb8 ? rgb_tmp : nullptr);
  rc = nsr_render_rays_forward(rays, n, packed_coarse, packed_fine, S, Ni, flags, nullptr, nullptr, rgb, disp_map, acc_map, rgb0, disp0,
                               acc0, z_std, nullptr, nullptr, nullptr, ws, workspace_bytes - size_t(ws - static_cast<uint8_t*>(workspace)),
                               stream);
  if (rc) return rc;
  if (rgb8) return nsr_to8b(rgb, n * 3, rgb8, stream);      // RN:246: to8b(rgbs[-1]) -> HWC uint8
  return NSR_OK;
}

}  // extern "C"
