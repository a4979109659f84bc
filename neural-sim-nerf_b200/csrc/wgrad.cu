// Weight gradients of one network pass (the loss.backward() of RN:691-707 as far as the MLP parameters go):
//   dW_l = G_l^T H_{l-1}   ([out x points] . [points x in]),   db_l = column sums of G_l,
// from the fp16 hi / residual operand pairs the backward kernel dumped (common.cuh "dump"), as the error-compensated product
// G_hi H_hi + G_lo H_hi + G_hi H_lo (three MMAs into one fp32 accumulator, like the forward pass).  The dump is already in the UMMA MN-major
// canonical layout, so 32-point slices of it go by cp.async.bulk straight into shared memory and serve as BOTH
// operands of tcgen05.mma with the POINT index as K: D[out feature (lane), in feature (column)] accumulates in
// TMEM over all the tiles a CTA owns (split-K over CTAs), and is added to the fp32 gradient with atomics at the end.
// The 1-row alpha head, the 3-row rgb head and the biases are skinny reductions done on CUDA cores.
#include "common.cuh"
#include "sm100_prims.cuh"

namespace nsr {

constexpr int WG_STAGES = 3;
constexpr int WG_STAGE_BYTES = 64 * 1024;   // [G hi | G lo | H hi | H lo], each a slice of 32 points x <=256 features (16 KB)
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 256;
constexpr int WG_NUM_JOBS = 12;

struct WJob {
  unsigned long long g_off, h_off;  // byte offsets of the two dump arrays
  int g_w, h_w;                     // their widths (features): g_w in {128, 256}, h_w in {32, 64, 256}
  float* out;                       // dW [g_w, ld] fp32
  int ld, col_off, n_valid;         // output columns [col_off, col_off + n_valid)
};
struct WJobs {
  WJob j[WG_NUM_JOBS];
};

__global__ void __launch_bounds__(128, 1) wgrad_gemm_kernel(const uint8_t* __restrict__ dump, size_t lo_off, WJobs jobs, int n_tiles,
                                                            const float* gscale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* done = empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const WJob job = jobs.j[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int my_tiles = (n_tiles > int(blockIdx.x)) ? (n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
  if (my_tiles == 0) return;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int m_halves = job.g_w / 128;

  if (tid == 0) {
    const uint32_t g_bytes = 32u * job.g_w * 2u, h_bytes = 32u * job.h_w * 2u;
    const uint32_t lbo_g = (job.g_w / 8) * 128, lbo_h = (job.h_w / 8) * 128;       // next 8 points
    const uint32_t idesc = make_idesc_f16(128, job.h_w) | (1u << 15) | (1u << 16);  // A and B MN-major
    const int n_slices = my_tiles * 4;
    const uint8_t* gbase = dump + job.g_off;
    const uint8_t* hbase = dump + job.h_off;
    for (int it = 0; it < n_slices + WG_STAGES - 1; ++it) {
      if (it < n_slices) {  // producer side: slice `it` -> stage it % STAGES
        const int s = it % WG_STAGES;
        if (it >= WG_STAGES) mbar_wait(&empty[s], ((it / WG_STAGES) - 1) & 1);
        const int tile = int(blockIdx.x) + (it >> 2) * int(gridDim.x), sl = it & 3;
        uint8_t* dst = smem + s * WG_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], 2 * (g_bytes + h_bytes));
        const uint8_t* gsrc = gbase + size_t(tile) * 128 * job.g_w * 2 + size_t(sl) * g_bytes;
        const uint8_t* hsrc = hbase + size_t(tile) * 128 * job.h_w * 2 + size_t(sl) * h_bytes;
        bulk_g2s(dst, gsrc, g_bytes, &full[s]);
        bulk_g2s(dst + 16384, gsrc + lo_off, g_bytes, &full[s]);
        bulk_g2s(dst + 32768, hsrc, h_bytes, &full[s]);
        bulk_g2s(dst + 49152, hsrc + lo_off, h_bytes, &full[s]);
      }
      const int jt = it - (WG_STAGES - 1);
      if (jt >= 0) {        // consumer side: slice `jt`
        const int s = jt % WG_STAGES;
        mbar_wait(&full[s], (jt / WG_STAGES) & 1);
        const uint32_t sg = smem_u32(smem + s * WG_STAGE_BYTES), sgl = sg + 16384, sh = sg + 32768, shl = sg + 49152;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // 16 points per MMA
          for (int mh = 0; mh < m_halves; ++mh) {
            const uint32_t ao = mh * 16 * 128 + ks * 2 * lbo_g, bo = ks * 2 * lbo_h;
            const uint64_t a_hi = make_sdesc(sg + ao, lbo_g, 128, 0), a_lo = make_sdesc(sgl + ao, lbo_g, 128, 0);
            const uint64_t b_hi = make_sdesc(sh + bo, lbo_h, 128, 0), b_lo = make_sdesc(shl + bo, lbo_h, 128, 0);
            umma_ss(tmem + mh * job.h_w, a_hi, b_hi, idesc, (jt | ks) != 0);
            umma_ss(tmem + mh * job.h_w, a_lo, b_hi, idesc, 1u);
            umma_ss(tmem + mh * job.h_w, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
      }
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after_sync();
  float scale = __uint_as_float(__float_as_uint(*gscale) & 0x7f800000u);
  if (!(scale > 0.f) || !(scale < 3.0e38f)) scale = 1.f;
  const uint32_t tlane = tmem + (uint32_t(warp * 32) << 16);
  for (int mh = 0; mh < m_halves; ++mh) {
    float* orow = job.out + size_t(mh * 128 + tid) * job.ld + job.col_off;
    for (int c0 = 0; c0 < job.h_w; c0 += 16) {
      uint32_t u[16];
      tmem_ld16(tlane + mh * job.h_w + c0, u);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < job.n_valid) atomicAdd(orow + c0 + j, __uint_as_float(u[j]) * scale);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ----------------------------------------------------------------------------- skinny reductions
// out[j][c] += sum_p coef[p][j] * X[p][c]  (coef = dL/draw columns, or 1 for the bias sums), X a blocked dump array
struct SJob {
  unsigned long long x_off;
  int w;              // width of X
  int coef_col;       // first column of d_raw used as coefficient, -1: coefficient 1 (scaled by gscale)
  int n_out;          // rows of out (1..3)
  float* out;         // [n_out][ld]
  int ld;
  float* out_coef;    // optional [n_out]: plain sums of the coefficients (bias of the head), written by thread 0
};
constexpr int SK_NUM_JOBS = 12;
struct SJobs {
  SJob j[SK_NUM_JOBS];
};

// One block = 256 threads walking whole 128-point tiles in 16-byte units (8 consecutive features of one point).  Inside a tile
// an array is (128/8) row blocks of W units each (unit index inside a row block = feature group * 8 + row % 8), so thread t of a
// warp reads 16 bytes next to its neighbours' (512 B per warp) and -- W being 128 or 256 -- always meets the same feature group
// t % W / 8: its 8 x n_out partial sums stay in registers over all tiles, are folded over the 8 rows of a feature group with
// shuffles, over the warps through shared memory, and leave the block as one atomicAdd per output element.
__global__ void __launch_bounds__(256) wgrad_skinny_kernel(const uint8_t* __restrict__ dump, size_t lo_off, const float* __restrict__ d_raw,
                                                           long long n_points, SJobs jobs, int n_tiles, const float* gscale) {
  const SJob job = jobs.j[blockIdx.y];
  const int t = threadIdx.x;
  const int W = job.w;                       // 128 or 256
  const int within = t % W;                  // unit inside a row block: constant per thread
  const int fg = within >> 3, rlow = within & 7;
  const int rb0 = t / W, rb_step = 256 / W;  // row blocks this thread visits: rb0, rb0 + rb_step, ...
  float acc[3][8];
  float csum[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[j][q] = 0.f;
  float scale = __uint_as_float(__float_as_uint(*gscale) & 0x7f800000u);
  if (!(scale > 0.f) || !(scale < 3.0e38f)) scale = 1.f;
  const bool has_coef = job.coef_col >= 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint8_t* base = dump + job.x_off + size_t(tile) * (size_t(128) * W * 2);
#pragma unroll 4
    for (int rb = rb0; rb < 16; rb += rb_step) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t(rb) * W + within) * 16);
      const uint4 ul = *reinterpret_cast<const uint4*>(base + lo_off + (size_t(rb) * W + within) * 16);
      float co[3] = {1.f, 0.f, 0.f};
      if (has_coef) {
        const long long p = (long long)tile * 128 + rb * 8 + rlow;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < n_points) d = reinterpret_cast<const float4*>(d_raw)[p];
        if (job.coef_col == 0) {          // rgb head: the three colour columns of dL/draw
          co[0] = d.x;
          co[1] = d.y;
          co[2] = d.z;
        } else {                          // alpha head: the sigma column
          co[0] = d.w;
        }
        if (fg == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) csum[j] += co[j];
        }
      }
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w}, l4[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[q]));
        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&l4[q]));
        f.x += fl.x;
        f.y += fl.y;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          acc[j][2 * q] = fmaf(co[j], f.x, acc[j][2 * q]);
          acc[j][2 * q + 1] = fmaf(co[j], f.y, acc[j][2 * q + 1]);
        }
      }
    }
  }
  // fold the 8 rows of a feature group (adjacent lanes), then the warps that share it
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = acc[j][q];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      acc[j][q] = v;
    }
    float c = csum[j];
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    csum[j] = c;
  }
  __shared__ float red[2][3][256];          // [row-block parity for W = 128][out row][feature]
  if (rlow == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int q = 0; q < 8; ++q) red[rb0][j][fg * 8 + q] = acc[j][q];
  }
  __shared__ float cred[2][3];
  if (within == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) cred[rb0][j] = csum[j];
  }
  __syncthreads();
  const float mul = has_coef ? 1.f : scale;
  if (t < W) {
    for (int j = 0; j < job.n_out; ++j) {
      float v = red[0][j][t];
      if (W == 128) v += red[1][j][t];
      atomicAdd(job.out + size_t(j) * job.ld + t, v * mul);
    }
  }
  if (t == 0 && job.out_coef != nullptr) {
    for (int j = 0; j < job.n_out; ++j) atomicAdd(job.out_coef + j, cred[0][j] + (W == 128 ? cred[1][j] : 0.f));
  }
}

int launch_weight_grads(const void* dump_v, const float* d_raw, int64_t n_points, const float* gscale, float* const* dW,
                        float* const* dB, cudaStream_t st) {
  if (n_points == 0) return NSR_OK;
  const uint8_t* dump = static_cast<const uint8_t*>(dump_v);
  const int n_tiles = int((n_points + 127) / 128);
  const size_t P = size_t(n_tiles) * 128;
  int num_sms = 0;
  if (int rc = current_device_sms(&num_sms)) return rc;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&wgrad_gemm_kernel), WG_SMEM)) return rc;
  // parameter order (include/nsr_b200.h): 0..7 pts_linears, 8 views_linears.0, 9 feature_linear, 10 alpha_linear, 11 rgb_linear
  WJobs g;
  int k = 0;
  auto add = [&](size_t g_off, int g_w, size_t h_off, int h_w, float* out, int ld, int col_off, int n_valid) {
    g.j[k++] = WJob{(unsigned long long)g_off, (unsigned long long)h_off, g_w, h_w, out, ld, col_off, n_valid};
  };
  add(dump_off_g(P, 0), 256, dump_off_ex(P), 64, dW[0], 63, 0, 63);                                   // pts_linears.0 [256,63]
  for (int l = 1; l <= 7; ++l) {
    if (l == 5) {                                                                                      // [256,319] = cat[xyz(63), h(256)]
      add(dump_off_g(P, 5), 256, dump_off_ex(P), 64, dW[5], 319, 0, 63);
      add(dump_off_g(P, 5), 256, dump_off_h(P, 4), 256, dW[5], 319, 63, 256);
    } else {
      add(dump_off_g(P, l), 256, dump_off_h(P, l - 1), 256, dW[l], 256, 0, 256);
    }
  }
  add(dump_off_gv(P), 128, dump_off_h(P, 8), 256, dW[8], 283, 0, 256);                                 // views_linears.0 [128,283] = cat[feature, dirs]
  add(dump_off_gv(P), 128, dump_off_ev(P), 32, dW[8], 283, 256, 27);
  add(dump_off_gf(P), 256, dump_off_h(P, 7), 256, dW[9], 256, 0, 256);                                 // feature_linear
  const int splits = 12;
  wgrad_gemm_kernel<<<dim3(splits, WG_NUM_JOBS), 128, WG_SMEM, st>>>(dump, dump_lo(P), g, n_tiles, gscale);
  count_launch();
  int rc = check_launch("wgrad_gemm_kernel");
  if (rc) return rc;

  SJobs s;
  k = 0;
  for (int l = 0; l <= 7; ++l) s.j[k++] = SJob{(unsigned long long)dump_off_g(P, l), 256, -1, 1, dB[l], 256, nullptr};   // db_l
  s.j[k++] = SJob{(unsigned long long)dump_off_gv(P), 128, -1, 1, dB[8], 128, nullptr};
  s.j[k++] = SJob{(unsigned long long)dump_off_gf(P), 256, -1, 1, dB[9], 256, nullptr};
  s.j[k++] = SJob{(unsigned long long)dump_off_h(P, 7), 256, 3, 1, dW[10], 256, dB[10]};                                   // alpha_linear: sigma column
  s.j[k++] = SJob{(unsigned long long)dump_off_hv(P), 128, 0, 3, dW[11], 128, dB[11]};                                     // rgb_linear: rgb columns
  wgrad_skinny_kernel<<<dim3(n_tiles < 74 ? n_tiles : 74, SK_NUM_JOBS), 256, 0, st>>>(dump, dump_lo(P), d_raw, (long long)n_points, s, n_tiles, gscale);
  count_launch();
  return check_launch("wgrad_skinny_kernel");
}

}  // namespace nsr
