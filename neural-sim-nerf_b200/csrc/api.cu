// C ABI of libnsr_b200 (include/nsr_b200.h): argument checking, error strings, and the
// forward orchestration of render_rays (RN:390-501) over the kernels in ray_stage.cu / mlp_forward.cu.
#include <atomic>
#include <mutex>
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
  static std::mutex mu;                       // host threads driving different devices may get here together
  std::lock_guard<std::mutex> lock(mu);
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

// two-tier evaluation knobs (process-wide; set them before launching work, not concurrently with it)
static TwoTierParams g_two_tier = {4.0f, 1.5f, 0.30f};   // tau, verify_max (measured max |sigma~ - sigma|: 0.74), force_fraction
static bool g_two_tier_enabled = true;
static bool g_coarse_refine = true;
const TwoTierParams& two_tier_params() { return g_two_tier; }

// byte offsets of the blocks inside nsr_render_rays_forward's workspace
struct FwdLayout {
  size_t z0, w0, raw0, z1, raw1, as0, as1, rf, total;
};
static FwdLayout fwd_layout(int64_t n, int S, int Ni) {
  const size_t T = size_t(S) + size_t(Ni);
  FwdLayout L;
  size_t b = 0;
  L.z0 = b;
  b += align_up(size_t(n) * S * 4, 256);
  L.w0 = b;
  b += align_up(size_t(n) * S * 4, 256);
  L.raw0 = b;
  b += align_up(size_t(n) * S * 16, 256);
  L.z1 = b;
  b += align_up(size_t(n) * T * 4, 256);
  L.raw1 = b;
  b += align_up(size_t(n) * T * 16, 256);
  L.as0 = b;
  b += n > 0 ? align_up(active_set_bytes(n * int64_t(S)), 256) : 0;
  L.as1 = b;
  b += n > 0 ? align_up(active_set_bytes(n * int64_t(T)), 256) : 0;
  L.rf = b;                                        // coarse-pass refinement list (refine.cu); after the blocks the layout call reports
  b += (n > 0 && Ni > 0) ? align_up(refine_workspace_bytes(n), 256) : 0;
  L.total = b;
  return L;
}

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

int nsr_chunk_issue_order(int step, int* half_out, int* kc_out, int capacity) {
  NSR_REQUIRE(step >= 0 && step < NUM_STEPS && half_out && kc_out, "nsr_chunk_issue_order: bad argument");
  const int nk = step_k_chunks(step), nhs = step_n_halves(step);
  NSR_REQUIRE(capacity >= nk * nhs, "nsr_chunk_issue_order: capacity %d < %d", capacity, nk * nhs);
  for (int i = 0; i < nk * nhs; ++i) issue_slot(nk, step_k_early(step), nhs, i, half_out[i], kc_out[i]);
  return nk * nhs;
}

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
  NSR_REQUIRE((rays || (flags & NSR_FLAG_EMBEDDED_INPUT)) && z_or_pts && packed_net && raw_out, "nsr_mlp_forward: null argument");
  NSR_REQUIRE(!((flags & NSR_FLAG_EMBEDDED_INPUT) && (flags & NSR_FLAG_PTS_INPUT)), "nsr_mlp_forward: PTS_INPUT and EMBEDDED_INPUT exclude each other");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(raw_out) & 15) == 0, "nsr_mlp_forward: raw_out must be 16-byte aligned");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(packed_net) & 127) == 0, "nsr_mlp_forward: packed_net must be 128-byte aligned");
  return launch_mlp_forward(rays, z_or_pts, n_rays, n_samples, packed_net, flags, raw_out, static_cast<cudaStream_t>(stream));
}

int nsr_mlp_two_tier(const float* rays, const float* z_vals, int64_t n_rays, int n_samples, const void* packed_net, float* raw_out,
                     void* active_set, void* relu_mask, int stages, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_samples > 0 && stages > 0 && stages < 8, "nsr_mlp_two_tier: bad sizes / stages");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(rays && z_vals && packed_net && raw_out && active_set, "nsr_mlp_two_tier: null argument");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(raw_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(active_set) & 255) == 0 &&
                  (reinterpret_cast<uintptr_t>(packed_net) & 127) == 0 && (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0,
              "nsr_mlp_two_tier: raw_out must be 16-byte, active_set 256-byte, packed_net 128-byte, relu_mask 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* mask = static_cast<uint32_t*>(relu_mask);
  int rc;
  if (stages & 1) {
    if (cudaMemsetAsync(active_set, 0, AS_CTRL_BYTES, st) != cudaSuccess) return check_launch("active_set init");
    if ((rc = launch_mlp_forward(rays, z_vals, n_rays, n_samples, packed_net, 0, raw_out, st, nullptr, nullptr, AS_ROLE_TIER1, active_set))) return rc;
  }
  if ((stages & 2) && (rc = launch_mlp_forward(rays, z_vals, n_rays, n_samples, packed_net, 0, raw_out, st, mask, nullptr, AS_ROLE_TIER2, active_set))) return rc;
  if ((stages & 4) && (rc = launch_mlp_forward(rays, z_vals, n_rays, n_samples, packed_net, 0, raw_out, st, mask, nullptr, AS_ROLE_REDO, active_set))) return rc;
  return NSR_OK;
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

// workspace layout: z0 [n,S] | w0 [n,S] | raw0 [n,S,4] | z1 [n,T] | raw1 [n,T,4] | active set of the coarse pass | of the fine pass
size_t nsr_render_workspace_bytes(int64_t n, int S, int Ni) { return fwd_layout(n, S, Ni).total; }

int nsr_render_workspace_layout(int64_t n, int S, int Ni, size_t* offsets_out, int capacity) {
  NSR_REQUIRE(n >= 0 && S > 0 && Ni >= 0 && offsets_out && capacity >= 8, "nsr_render_workspace_layout: bad argument (capacity >= 8)");
  const FwdLayout L = fwd_layout(n, S, Ni);
  const size_t v[8] = {L.z0, L.w0, L.raw0, L.z1, L.raw1, L.as0, L.as1, L.total};
  for (int i = 0; i < 8; ++i) offsets_out[i] = v[i];
  return 8;
}

size_t nsr_active_set_bytes(int64_t n_rays, int n_total_samples) { return active_set_bytes(n_rays * int64_t(n_total_samples)); }

int nsr_set_two_tier(int enabled, float tau, float verify_max, float force_fraction) {
  NSR_REQUIRE(tau > 0.f && verify_max >= 0.f && verify_max < tau && force_fraction >= 0.f,
              "nsr_set_two_tier: need tau > verify_max >= 0 and force_fraction >= 0");
  g_two_tier_enabled = enabled != 0;
  g_two_tier = TwoTierParams{tau, verify_max, force_fraction};
  return NSR_OK;
}

int nsr_set_tier1_pair(int enabled) { return set_tier1_pair(enabled); }

int nsr_set_coarse_refine(int enabled) {
  const int old = g_coarse_refine ? 1 : 0;
  g_coarse_refine = enabled != 0;
  return old;
}

int nsr_coarse_refine(const float* rays, const float* z_vals, int64_t n_rays, int n_samples, const void* packed_net, float* raw, void* workspace,
                      size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_samples >= 2, "nsr_coarse_refine: bad sizes");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(rays && z_vals && packed_net && raw && workspace, "nsr_coarse_refine: null argument");
  NSR_REQUIRE(workspace_bytes >= refine_workspace_bytes(n_rays) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "nsr_coarse_refine: workspace too small or not 256-byte aligned");
  return launch_coarse_refine(rays, z_vals, n_rays, n_samples, packed_net, raw, workspace, static_cast<cudaStream_t>(stream));
}

size_t nsr_coarse_refine_workspace_bytes(int64_t n_rays) { return refine_workspace_bytes(n_rays); }

int nsr_set_coarse_refine_limit(float acc_limit) {
  NSR_REQUIRE(acc_limit > 0.f && acc_limit < 1.f, "nsr_set_coarse_refine_limit: acc_limit must be in (0, 1)");
  set_refine_tau_limit(-logf(1.f - acc_limit));
  return NSR_OK;
}

int nsr_set_coarse_refine_sigma(float sigma_hi) {
  set_refine_sigma_hi(sigma_hi);
  return NSR_OK;
}

int nsr_get_two_tier(int* enabled, float* tau, float* verify_max, float* force_fraction) {
  if (enabled) *enabled = g_two_tier_enabled ? 1 : 0;
  if (tau) *tau = g_two_tier.tau;
  if (verify_max) *verify_max = g_two_tier.verify_max;
  if (force_fraction) *force_fraction = g_two_tier.force_frac;
  return NSR_OK;
}

size_t nsr_relu_mask_bytes(int64_t n_rays, int n_total_samples) { return relu_mask_bytes(n_rays * int64_t(n_total_samples)); }

int nsr_render_rays_forward(const float* rays, int64_t n, const void* packed_coarse, const void* packed_fine, int S,
                            int Ni, uint32_t flags, const float* t_rand, const float* u, float* rgb_map, float* disp_map,
                            float* acc_map, float* rgb0, float* disp0, float* acc0, float* z_std, float* raw,
                            float* z_vals_out, float* weights_out, void* workspace, size_t workspace_bytes, void* stream) {
  return nsr_render_rays_forward_ex(rays, n, packed_coarse, packed_fine, S, Ni, flags, t_rand, u, rgb_map, disp_map, acc_map, rgb0, disp0,
                                    acc0, z_std, raw, z_vals_out, weights_out, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

// One network pass over (n, S) points into raw: dense fp16x3 (as = NULL), or tier 1 -> tier 2 -> conditional re-evaluation.
static int mlp_pass(const float* rays, const float* z, int64_t n, int S, const void* packed, uint32_t mflags, float* raw, cudaStream_t st,
                    uint32_t* mask, void* dump, void* as) {
  if (as == nullptr) return launch_mlp_forward(rays, z, n, S, packed, mflags, raw, st, mask, dump);
  int rc;
  if ((rc = launch_mlp_forward(rays, z, n, S, packed, 0, raw, st, nullptr, nullptr, AS_ROLE_TIER1, as))) return rc;
  if ((rc = launch_mlp_forward(rays, z, n, S, packed, 0, raw, st, mask, nullptr, AS_ROLE_TIER2, as))) return rc;
  return launch_mlp_forward(rays, z, n, S, packed, 0, raw, st, mask, nullptr, AS_ROLE_REDO, as);   // exits at once unless verification failed
}

int nsr_render_rays_forward_ex(const float* rays, int64_t n, const void* packed_coarse, const void* packed_fine, int S,
                               int Ni, uint32_t flags, const float* t_rand, const float* u, float* rgb_map, float* disp_map,
                               float* acc_map, float* rgb0, float* disp0, float* acc0, float* z_std, float* raw,
                               float* z_vals_out, float* weights_out, void* relu_mask, void* dump_out, void* active_set,
                               void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(n >= 0 && S >= 2 && Ni >= 0, "nsr_render_rays_forward: bad sizes (needs at least 2 samples per ray)");
  NSR_REQUIRE(relu_mask == nullptr || (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0, "nsr_render_rays_forward: relu_mask must be 16-byte aligned");
  NSR_REQUIRE(dump_out == nullptr || relu_mask != nullptr, "nsr_render_rays_forward: dump_out goes with relu_mask (the backward pass needs both)");
  NSR_REQUIRE(dump_out == nullptr || (reinterpret_cast<uintptr_t>(dump_out) & 127) == 0, "nsr_render_rays_forward: dump_out must be 128-byte aligned");
  NSR_REQUIRE(dump_out == nullptr || !(flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8)), "nsr_render_rays_forward: dump_out needs the default precision");
  if (n == 0) return NSR_OK;
  NSR_REQUIRE(rays && packed_coarse, "nsr_render_rays_forward: null rays / weights");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_render_workspace_bytes(n, S, Ni), "nsr_render_rays_forward: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "nsr_render_rays_forward: workspace must be 256-byte aligned");
  NSR_REQUIRE(Ni == 0 || S >= 3, "nsr_render_rays_forward: hierarchical sampling needs n_samples >= 3");
  NSR_REQUIRE(active_set == nullptr || (reinterpret_cast<uintptr_t>(active_set) & 255) == 0, "nsr_render_rays_forward: active_set must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = S + Ni;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const FwdLayout lay = fwd_layout(n, S, Ni);
  float* z0 = reinterpret_cast<float*>(ws + lay.z0);
  float* w0 = reinterpret_cast<float*>(ws + lay.w0);
  float* raw0 = reinterpret_cast<float*>(ws + lay.raw0);
  float* z1 = reinterpret_cast<float*>(ws + lay.z1);
  float* raw1 = reinterpret_cast<float*>(ws + lay.raw1);
  const uint32_t cflags = flags & NSR_FLAG_WHITE_BKGD;
  const uint32_t mflags = flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8);
  int rc;
  // Two-tier evaluation (common.cuh "active set"): same outputs as the dense fp16x3 passes, bit for bit, except that `raw` keeps
  // (0, 0, 0, sigma~) for the certified-empty points -- so a caller that wants raw (or the sign bits, whose order follows the
  // active list) must say it knows by passing an active_set buffer.
  const bool two_tier = g_two_tier_enabled && !(flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8 | NSR_FLAG_DENSE)) && dump_out == nullptr &&
                        ((raw == nullptr && relu_mask == nullptr) || active_set != nullptr) && n * int64_t(T) < (int64_t(1) << 31);
    void* as0 = two_tier ? ws + lay.as0 : nullptr;                                     // coarse pass (when it is not the last one)
  void* as_last = two_tier ? (active_set ? active_set : static_cast<void*>(ws + (Ni > 0 ? lay.as1 : lay.as0))) : nullptr;
  if (active_set != nullptr && !two_tier) {
    // the caller will hand this buffer to the backward pass: mark it "everything, in order"
    const uint32_t dense_ctrl[AS_CTRL_WORDS] = {0u, 0u, 1u, 0u};
    if (cudaMemcpyAsync(active_set, dense_ctrl, sizeof(dense_ctrl), cudaMemcpyHostToDevice, st) != cudaSuccess) return check_launch("active_set init");
  }
  if (two_tier) {
    if (cudaMemsetAsync(Ni > 0 ? as0 : as_last, 0, AS_CTRL_BYTES, st) != cudaSuccess) return check_launch("active_set init");
    if (Ni > 0 && cudaMemsetAsync(as_last, 0, AS_CTRL_BYTES, st) != cudaSuccess) return check_launch("active_set init");
  }

  uint32_t* mask = static_cast<uint32_t*>(relu_mask);      // sign bits of the LAST pass: the only one that carries gradient to the rays
  if ((rc = launch_coarse_z(rays, n, S, flags, t_rand, z0, st))) return rc;                          // RN:439-461
  if ((rc = mlp_pass(rays, z0, n, S, packed_coarse, mflags, raw0, st, Ni == 0 ? mask : nullptr, Ni == 0 ? dump_out : nullptr,
                     Ni == 0 ? as_last : as0))) return rc;                                            // RN:463-466
  if (Ni == 0) {
    if ((rc = launch_raw2outputs(raw0, z0, rays + 3, 11, n, S, cflags, rgb_map, disp_map, acc_map, weights_out, nullptr, st))) return rc;
    if (raw) cudaMemcpyAsync(raw, raw0, size_t(n) * S * 16, cudaMemcpyDeviceToDevice, st);
    if (z_vals_out) cudaMemcpyAsync(z_vals_out, z0, size_t(n) * S * 4, cudaMemcpyDeviceToDevice, st);
    return check_launch("render_rays_forward(coarse only)");
  }
  // hierarchical sampling is ill-conditioned in the density of barely-hit rays: those few points again, in fp32 (refine.cu)
  if (g_coarse_refine && !mflags && (rc = launch_coarse_refine(rays, z0, n, S, packed_coarse, raw0, ws + lay.rf, st))) return rc;
  if ((rc = launch_raw2outputs(raw0, z0, rays + 3, 11, n, S, cflags, rgb0, disp0, acc0, w0, nullptr, st))) return rc;  // RN:467
  float* zf = z_vals_out ? z_vals_out : z1;
  const uint32_t force_count = uint32_t(double(g_two_tier.force_frac) * double(n) * double(S));
  if ((rc = launch_resample_merge(z0, w0, n, S, Ni, u, zf, nullptr, z_std, st, static_cast<const uint32_t*>(as0),
                                  static_cast<uint32_t*>(as_last), force_count))) return rc;        // RN:473-477, 495
  float* rawf = raw ? raw : raw1;
  if ((rc = mlp_pass(rays, zf, n, T, packed_fine ? packed_fine : packed_coarse, mflags, rawf, st, mask, dump_out, as_last))) return rc;  // RN:478-483
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
  return nsr_render_rays_backward_ex(rays, z_vals, raw, n, T, packed_net, flags, d_rgb_map, d_rays, dump, dW, dB, nullptr, nullptr, workspace,
                                     workspace_bytes, stream);
}

int nsr_render_rays_backward_ex(const float* rays, const float* z_vals, const float* raw, int64_t n, int T, const void* packed_net,
                                uint32_t flags, const float* d_rgb_map, float* d_rays, void* dump, float* const* dW,
                                float* const* dB, const void* relu_mask, const void* active_set, void* workspace, size_t workspace_bytes,
                                void* stream) {
  NSR_REQUIRE(n >= 0 && T > 0, "nsr_render_rays_backward: bad sizes");
  NSR_REQUIRE(relu_mask == nullptr || (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0, "nsr_render_rays_backward: relu_mask must be 16-byte aligned");
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
  NSR_REQUIRE(active_set == nullptr || (dW == nullptr && dump == nullptr), "nsr_render_rays_backward: the active-set route gives dL/d(rays) only");
  NSR_REQUIRE(active_set == nullptr || (reinterpret_cast<uintptr_t>(active_set) & 255) == 0, "nsr_render_rays_backward: active_set must be 256-byte aligned");
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
  // active set: only its points are back-propagated (dL/draw is exactly 0 everywhere else: alpha == 0 there), the rest stays 0
  if (active_set != nullptr && cudaMemsetAsync(d_pts, 0, size_t(n) * T * 32, st) != cudaSuccess) return check_launch("d_pts memset");
  if ((rc = launch_mlp_backward(rays, z_vals, n, T, packed_net, d_raw, d_pts, dW ? dump : nullptr, gmax, st,
                                static_cast<const uint32_t*>(relu_mask), active_set))) return rc;
  if ((rc = launch_ray_grad_reduce(rays, z_vals, d_pts, d_dnorm, n, T, d_rays, st))) return rc;
  if (dW) return launch_weight_grads(dump, d_raw, n * int64_t(T), gmax, dW, dB, st);
  return NSR_OK;
}

int nsr_mlp_backward(const float* rays, const float* z_vals, int64_t n_rays, int n_total_samples, const void* packed_net, const float* d_raw,
                     float* d_pts, const void* relu_mask, const void* active_set, void* stream) {
  NSR_REQUIRE(n_rays >= 0 && n_total_samples > 0, "nsr_mlp_backward: bad sizes");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(rays && z_vals && packed_net && d_raw && d_pts, "nsr_mlp_backward: null argument");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(d_raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_pts) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(active_set) & 255) == 0,
              "nsr_mlp_backward: d_raw / d_pts / relu_mask must be 16-byte, active_set 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (active_set != nullptr && cudaMemsetAsync(d_pts, 0, size_t(n_rays) * n_total_samples * 32, st) != cudaSuccess) return check_launch("d_pts memset");
  return launch_mlp_backward(rays, z_vals, n_rays, n_total_samples, packed_net, d_raw, d_pts, nullptr, nullptr, st,
                             static_cast<const uint32_t*>(relu_mask), active_set);
}

int nsr_make_rays(int H, int W, const float* K_host, const float* c2w_host, float near_, float far_, float* rays_out, void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host && c2w_host && rays_out, "nsr_make_rays: bad argument");
  return launch_make_rays(H, W, K_host, c2w_host, near_, far_, rays_out, static_cast<cudaStream_t>(stream));
}

int nsr_pack_rays(const float* rays_o, const float* rays_d, int64_t n_rays, float near_, float far_, float* rays_out, void* stream) {
  NSR_REQUIRE(n_rays >= 0, "nsr_pack_rays: bad size");
  if (n_rays == 0) return NSR_OK;
  NSR_REQUIRE(rays_o && rays_d && rays_out, "nsr_pack_rays: null argument");
  return launch_pack_rays(rays_o, rays_d, n_rays, near_, far_, rays_out, static_cast<cudaStream_t>(stream));
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

// workspace layout: rays [n,11] | z [n,T] | raw [n,T,4] | d_rays [n,11] | rgb [n,3] (when rgb_map is NULL) | c2w-gradient partials |
//   ReLU sign bits (n, T) | forward workspace | backward workspace
size_t nsr_render_image_grad_workspace_bytes(int H, int W, int S, int Ni) {
  const int64_t n = int64_t(H) * W;
  const int T = S + Ni;
  return align_up(size_t(n) * 44, 256) + align_up(size_t(n) * T * 4, 256) + align_up(size_t(n) * T * 16, 256) + align_up(size_t(n) * 44, 256) +
         align_up(size_t(n) * 12, 256) + align_up(nsr_c2w_grad_workspace_bytes(), 256) + align_up(nsr_relu_mask_bytes(n, T), 256) +
         align_up(nsr_active_set_bytes(n, T), 256) + nsr_render_workspace_bytes(n, S, Ni) + nsr_render_backward_workspace_bytes(n, T);
}

int nsr_render_image_grad(int H, int W, const float* K_host, const float* c2w_dev, int ld_c2w, float near_, float far_,
                          const void* packed_coarse, const void* packed_fine, int S, int Ni, uint32_t flags, const float* d_rgb_map,
                          float* rgb_map, float* d_c2w, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(H > 0 && W > 0 && K_host && c2w_dev && ld_c2w >= 4 && d_rgb_map && d_c2w, "nsr_render_image_grad: bad argument");
  NSR_REQUIRE(!(flags & (NSR_FLAG_FAST_FP16 | NSR_FLAG_MIXED_F8)), "nsr_render_image_grad: gradient passes run the default precision");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_render_image_grad_workspace_bytes(H, W, S, Ni), "nsr_render_image_grad: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "nsr_render_image_grad: workspace must be 256-byte aligned");
  const int64_t n = int64_t(H) * W;
  const int T = S + Ni;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto carve = [&](size_t bytes) {
    uint8_t* p = ws;
    ws += align_up(bytes, 256);
    return p;
  };
  float* rays = reinterpret_cast<float*>(carve(size_t(n) * 44));
  float* zv = reinterpret_cast<float*>(carve(size_t(n) * T * 4));
  float* raw = reinterpret_cast<float*>(carve(size_t(n) * T * 16));
  float* d_rays = reinterpret_cast<float*>(carve(size_t(n) * 44));
  float* rgb_tmp = reinterpret_cast<float*>(carve(size_t(n) * 12));
  void* cws = carve(nsr_c2w_grad_workspace_bytes());
  void* bits = carve(nsr_relu_mask_bytes(n, T));
  void* aset = carve(nsr_active_set_bytes(n, T));
  const size_t fwd_bytes = nsr_render_workspace_bytes(n, S, Ni), bwd_bytes = nsr_render_backward_workspace_bytes(n, T);
  void* fwd = carve(fwd_bytes);
  void* bwd = carve(bwd_bytes);
  int rc;
  if ((rc = nsr_make_rays_dev(H, W, K_host, c2w_dev, ld_c2w, near_, far_, rays, stream))) return rc;                       // RN:148
  if ((rc = nsr_render_rays_forward_ex(rays, n, packed_coarse, packed_fine, S, Ni, flags, nullptr, nullptr, rgb_map ? rgb_map : rgb_tmp, nullptr,
                                       nullptr, nullptr, nullptr, nullptr, nullptr, raw, zv, nullptr, bits, nullptr, aset, fwd, fwd_bytes, stream))) return rc;   // RN:168-170
  const void* last = (Ni > 0 && packed_fine) ? packed_fine : packed_coarse;
  if ((rc = nsr_render_rays_backward_ex(rays, zv, raw, n, Ni > 0 ? T : S, last, flags & NSR_FLAG_WHITE_BKGD, d_rgb_map, d_rays, nullptr, nullptr,
                                        nullptr, bits, aset, bwd, bwd_bytes, stream))) return rc;                                                           // RN:177-178
  return nsr_rays_grad_to_c2w(H, W, K_host, rays, d_rays, nullptr, n, d_c2w, accumulate, cws, stream);                                                      // RN:179-181
}

// ----------------------------------------------------------------------------- one optimisation step (RN:643-716)
static const int kParamRows[NSR_NET_NUM_TENSORS] = {256, 256, 256, 256, 256, 256, 256, 256, 128, 256, 1, 3};
static const int kParamCols[NSR_NET_NUM_TENSORS] = {63, 256, 256, 256, 256, 319, 256, 256, 283, 256, 256, 128};
static const size_t kNetParamsPadded = 595968;  // multiple of 64 floats

int nsr_random_uniform(uint64_t seed, uint32_t stream_id, float* out, int64_t count, void* stream) {
  NSR_REQUIRE(count >= 0 && (count == 0 || out), "nsr_random_uniform: bad argument");
  return launch_uniform(seed, stream_id, out, count, static_cast<cudaStream_t>(stream));
}

int nsr_add_sigma_noise(uint64_t seed, uint32_t stream_id, float* raw, int64_t n_points, float std, void* stream) {
  NSR_REQUIRE(n_points >= 0 && (n_points == 0 || raw), "nsr_add_sigma_noise: bad argument");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "nsr_add_sigma_noise: raw must be 16-byte aligned");
  return launch_sigma_noise(seed, stream_id, raw, n_points, std, static_cast<cudaStream_t>(stream));
}

// workspace layout (each block padded to 256 B):
//   forward workspace (nsr_render_workspace_bytes) | t_rand [n,S] | u [n,Ni] | rgb [n,3] | rgb0 [n,3] | d_rgb [n,3] | d_rgb0 [n,3] |
//   d_rays [n,11] | backward workspace (n, T) | dump + sign bits of the last pass (n, T) | (Ni > 0) dump + sign bits of the coarse pass
//   (n, S) | gradients 2 x 595 968 floats.  Both forward passes write their activations (fp16) and ReLU sign bits, so neither backward
//   pass recomputes anything.
size_t nsr_train_workspace_bytes(int64_t n, int S, int Ni) {
  const int T = S + Ni;
  size_t b = nsr_render_workspace_bytes(n, S, Ni);
  b += align_up(size_t(n) * S * 4, 256) + align_up(size_t(n) * (Ni > 0 ? Ni : 1) * 4, 256);
  b += 4 * align_up(size_t(n) * 12, 256) + align_up(size_t(n) * 44, 256);
  b += nsr_render_backward_workspace_bytes(n, T);
  b += align_up(nsr_mlp_dump_bytes(n, T), 256) + align_up(nsr_relu_mask_bytes(n, T), 256);
  if (Ni > 0) b += align_up(nsr_mlp_dump_bytes(n, S), 256) + align_up(nsr_relu_mask_bytes(n, S), 256);
  b += 2 * kNetParamsPadded * 4;
  return b;
}

int nsr_train_step(const float* rays, const float* target, int64_t n, const nsr_train_net* coarse, const nsr_train_net* fine, int S, int Ni,
                   uint32_t flags, int perturb, float raw_noise_std, uint64_t seed, float lr, float beta1, float beta2, float eps,
                   int64_t step, float* losses_out, float* rgb_out, void* workspace, size_t workspace_bytes, void* stream) {
  NSR_REQUIRE(n > 0 && S >= 2 && Ni >= 0 && step >= 1, "nsr_train_step: bad sizes (n_rays=%lld, step=%lld)", (long long)n, (long long)step);
  NSR_REQUIRE(Ni == 0 || S >= 3, "nsr_train_step: hierarchical sampling needs n_samples >= 3");
  NSR_REQUIRE(rays && target && coarse && losses_out, "nsr_train_step: null argument");
  NSR_REQUIRE(workspace && workspace_bytes >= nsr_train_workspace_bytes(n, S, Ni), "nsr_train_step: workspace too small");
  NSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "nsr_train_step: workspace must be 256-byte aligned");
  if (fine != nullptr && Ni == 0) fine = nullptr;   // RN:272: no fine pass, the fine network is never evaluated
  const nsr_train_net* nets[2] = {coarse, fine};
  for (int k = 0; k < 2; ++k) {
    if (!nets[k]) continue;
    NSR_REQUIRE(nets[k]->params && nets[k]->exp_avg && nets[k]->exp_avg_sq && nets[k]->packed, "nsr_train_step: null network field");
    for (int i = 0; i < 2 * NSR_NET_NUM_TENSORS; ++i)
      NSR_REQUIRE(nets[k]->params[i] && nets[k]->exp_avg[i] && nets[k]->exp_avg_sq[i], "nsr_train_step: null tensor %d", i);
    NSR_REQUIRE((reinterpret_cast<uintptr_t>(nets[k]->packed) & 127) == 0, "nsr_train_step: packed blob must be 128-byte aligned");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = S + Ni;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto carve = [&](size_t bytes) {
    uint8_t* p = ws;
    ws += align_up(bytes, 256);
    return p;
  };
  const size_t fwd_bytes = nsr_render_workspace_bytes(n, S, Ni);
  uint8_t* fwd = carve(fwd_bytes);
  float* z0 = reinterpret_cast<float*>(fwd);                                      // layout of nsr_render_workspace_bytes
  float* w0 = reinterpret_cast<float*>(fwd + align_up(size_t(n) * S * 4, 256));
  float* raw0 = reinterpret_cast<float*>(fwd + 2 * align_up(size_t(n) * S * 4, 256));
  float* z1 = reinterpret_cast<float*>(fwd + 2 * align_up(size_t(n) * S * 4, 256) + align_up(size_t(n) * S * 16, 256));
  float* raw1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(z1) + align_up(size_t(n) * T * 4, 256));
  float* t_rand = reinterpret_cast<float*>(carve(size_t(n) * S * 4));
  float* u = reinterpret_cast<float*>(carve(size_t(n) * (Ni > 0 ? Ni : 1) * 4));
  float* rgb = reinterpret_cast<float*>(carve(size_t(n) * 12));
  float* rgb0 = reinterpret_cast<float*>(carve(size_t(n) * 12));
  float* d_rgb = reinterpret_cast<float*>(carve(size_t(n) * 12));
  float* d_rgb0 = reinterpret_cast<float*>(carve(size_t(n) * 12));
  float* d_rays = reinterpret_cast<float*>(carve(size_t(n) * 44));
  const size_t bwd_bytes = nsr_render_backward_workspace_bytes(n, T);
  void* bwd = carve(bwd_bytes);
  void* dump = carve(nsr_mlp_dump_bytes(n, T));
  uint32_t* bits = reinterpret_cast<uint32_t*>(carve(nsr_relu_mask_bytes(n, T)));
  void* dump0 = Ni > 0 ? carve(nsr_mlp_dump_bytes(n, S)) : nullptr;
  uint32_t* bits0 = Ni > 0 ? reinterpret_cast<uint32_t*>(carve(nsr_relu_mask_bytes(n, S))) : nullptr;
  float* grads = reinterpret_cast<float*>(carve(2 * kNetParamsPadded * 4));
  if (rgb_out) rgb = rgb_out;

  // gradient tensors of both networks inside the flat buffer, reference shapes
  float* dW[2][NSR_NET_NUM_TENSORS];
  float* dB[2][NSR_NET_NUM_TENSORS];
  for (int k = 0; k < 2; ++k) {
    float* g = grads + k * kNetParamsPadded;
    for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) {
      dW[k][i] = g;
      g += size_t(kParamRows[i]) * kParamCols[i];
    }
    for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) {
      dB[k][i] = g;
      g += kParamRows[i];
    }
  }
  const uint32_t cflags = flags & NSR_FLAG_WHITE_BKGD;
  const uint32_t zflags = flags & NSR_FLAG_LINDISP;
  const void* pc = coarse->packed;
  const void* pl = fine ? fine->packed : coarse->packed;   // network of the last pass (RN:481)
  const int last = fine ? 1 : 0;
  int rc;
  // ---- forward (RN:390-501 with perturb / raw_noise_std as the training kwargs set them)
  if (perturb) {
    if ((rc = launch_uniform(seed, 0u, t_rand, n * int64_t(S), st))) return rc;                                   // RN:455
    if (Ni > 0 && (rc = launch_uniform(seed, 1u, u, n * int64_t(Ni), st))) return rc;                             // RH:211
  }
  if ((rc = launch_coarse_z(rays, n, S, zflags, perturb ? t_rand : nullptr, z0, st))) return rc;
  if ((rc = launch_mlp_forward(rays, z0, n, S, pc, 0, raw0, st, Ni > 0 ? bits0 : bits, Ni > 0 ? dump0 : dump))) return rc;
  if (Ni > 0 && g_coarse_refine && (rc = launch_coarse_refine(rays, z0, n, S, pc, raw0, fwd + fwd_layout(n, S, Ni).rf, st))) return rc;   // as render() does
  if (raw_noise_std > 0.f && (rc = launch_sigma_noise(seed, 2u, raw0, n * int64_t(S), raw_noise_std, st))) return rc;  // RN:365-366
  if ((rc = launch_raw2outputs(raw0, z0, rays + 3, 11, n, S, cflags, Ni > 0 ? rgb0 : rgb, nullptr, nullptr, w0, nullptr, st))) return rc;
  if (Ni > 0) {
    if ((rc = launch_resample_merge(z0, w0, n, S, Ni, perturb ? u : nullptr, z1, nullptr, nullptr, st))) return rc;
    if ((rc = launch_mlp_forward(rays, z1, n, T, pl, 0, raw1, st, bits, dump))) return rc;
    if (raw_noise_std > 0.f && (rc = launch_sigma_noise(seed, 3u, raw1, n * int64_t(T), raw_noise_std, st))) return rc;
    if ((rc = launch_raw2outputs(raw1, z1, rays + 3, 11, n, T, cflags, rgb, nullptr, nullptr, nullptr, nullptr, st))) return rc;
  }
  // ---- loss = img2mse(rgb, target) [+ img2mse(rgb0, target)]  (RN:696-704) and its gradient
  if ((rc = launch_mse_grad(rgb, target, n * 3, d_rgb, losses_out, st))) return rc;
  if (Ni > 0) {
    if ((rc = launch_mse_grad(rgb0, target, n * 3, d_rgb0, losses_out + 1, st))) return rc;
  } else {
    cudaMemsetAsync(losses_out + 1, 0, 4, st);
  }
  // ---- loss.backward() (RN:705): the last pass through rgb, the coarse pass through rgb0
  cudaMemsetAsync(grads, 0, 2 * kNetParamsPadded * 4, st);
  if ((rc = nsr_render_rays_backward_ex(rays, Ni > 0 ? z1 : z0, Ni > 0 ? raw1 : raw0, n, Ni > 0 ? T : S, pl, cflags, d_rgb, d_rays, dump, dW[last],
                                        dB[last], bits, nullptr, bwd, bwd_bytes, stream))) return rc;
  if (Ni > 0 && (rc = nsr_render_rays_backward_ex(rays, z0, raw0, n, S, pc, cflags, d_rgb0, d_rays, dump0, dW[0], dB[0], bits0, nullptr, bwd, bwd_bytes,
                                                  stream))) return rc;
  // ---- optimizer.step() (RN:707)
  AdamJobs jobs;
  jobs.count = 0;
  for (int k = 0; k < 2; ++k) {
    if (!nets[k]) continue;
    for (int i = 0; i < 2 * NSR_NET_NUM_TENSORS; ++i) {
      const int t = i % NSR_NET_NUM_TENSORS;
      const bool is_w = i < NSR_NET_NUM_TENSORS;
      jobs.j[jobs.count++] = AdamJob{nets[k]->params[i], is_w ? dW[k][t] : dB[k][t], nets[k]->exp_avg[i], nets[k]->exp_avg_sq[i],
                                     is_w ? kParamRows[t] * kParamCols[t] : kParamRows[t]};
    }
  }
  if ((rc = launch_adam(jobs, beta1, beta2, lr, eps, step, st))) return rc;
  // ---- the kernels read packed operands: re-pack what changed
  for (int k = 0; k < 2; ++k) {
    if (!nets[k]) continue;
    if ((rc = launch_pack_net(nets[k]->params, nets[k]->params + NSR_NET_NUM_TENSORS, nets[k]->packed, st))) return rc;
  }
  return NSR_OK;
}

}  // extern "C"
