// Positional encoding + 8x256 NeRF MLP (RH:18-66, RH:99-122, RN:26-40) as one tcgen05 kernel.
//
// A CTA owns 128 sample points at a time (one UMMA M=128 tile; thread r of the four compute
// warps == row r == TMEM lane r).  All ten GEMM steps of the network run back to back on that
// tile without touching HBM:
//   * the weights arrive as pre-packed 16 KB fp16 operand chunks (common.cuh) streamed by the
//     TMA engine (cp.async.bulk, mbarrier complete_tx) through a 4-deep shared-memory ring;
//   * the A operand (encodings / activations, fp16) lives in shared memory in the UMMA K-major
//     canonical layout and is rewritten in place by the epilogue of the previous step;
//   * accumulators live in TMEM (128 lanes x 256 fp32 columns); the epilogue reads them with
//     tcgen05.ld, adds the bias, applies ReLU, rounds to fp16 and writes the next A operand;
//   * the 1-wide alpha head and 3-wide rgb head are dot products evaluated in fp32 on CUDA
//     cores inside the step-7 / step-9 epilogues (they would waste a whole MMA N-tile).
// Warp roles: warps 0-3 compute (encode + epilogue), warp 4 MMA issuer, warp 5 weight producer.
// Only rays, depths and the packed weights are read from HBM, only raw [P,4] is written.
#include <math.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace nsr {

// ----------------------------------------------------------------------------- weight packing
// canonical no-swizzle K-major offset (bytes) of element (row, k) inside a [128 x 64] fp16 chunk
__host__ __device__ __forceinline__ int chunk_off(int row, int k) { return (row >> 3) * 1024 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

struct NetPtrs {
  const float* w[NSR_NET_NUM_TENSORS];
  const float* b[NSR_NET_NUM_TENSORS];
};

// chunk index -> (step, n_half, k_chunk)
__device__ __forceinline__ void chunk_decode(int c, int& step, int& nh, int& kc) {
  int s = 0;
  for (; s < NUM_STEPS; ++s) {
    const int cnt = step_n_halves(s) * step_k_chunks(s);
    if (c < cnt) break;
    c -= cnt;
  }
  step = s;
  nh = c / step_k_chunks(s);
  kc = c % step_k_chunks(s);
}

__global__ void pack_net_kernel(NetPtrs p, uint8_t* __restrict__ out) {
  const int c = blockIdx.x;
  if (c < NUM_CHUNKS) {
    int step, nh, kc;
    chunk_decode(c, step, nh, kc);
    __half* dst = reinterpret_cast<__half*>(out + size_t(c) * CHUNK_BYTES);
    for (int e = threadIdx.x; e < CHUNK_ROWS * CHUNK_K; e += blockDim.x) {
      const int nl = e / CHUNK_K, kl = e % CHUNK_K;
      const int n = nh * 128 + nl;
      float v = 0.f;
      if (step == 0) {
        if (kl < 63) v = p.w[0][n * 63 + kl];
      } else if (step == 5) {  // cat[input_pts(63), h(256)]  RH:106
        if (kc == 0) {
          if (kl < 63) v = p.w[5][n * 319 + kl];
        } else {
          v = p.w[5][n * 319 + 63 + (kc - 1) * 64 + kl];
        }
      } else if (step <= 7) {
        v = p.w[step][n * 256 + kc * 64 + kl];
      } else if (step == 8) {  // feature_linear
        v = p.w[9][n * 256 + kc * 64 + kl];
      } else {  // views_linears.0 on cat[feature(256), dirs(27)]  RH:111
        if (kc < 4) v = p.w[8][n * 283 + kc * 64 + kl];
        else if (kl < 27) v = p.w[8][n * 283 + 256 + kl];
      }
      dst[chunk_off(nl, kl) >> 1] = __float2half_rn(v);
    }
  } else {  // fp32 tail
    float* t = reinterpret_cast<float*>(out + WEIGHT_BYTES);
    for (int i = threadIdx.x; i < TAIL_FLOATS; i += blockDim.x) {
      float v = 0.f;
      if (i < TAIL_WALPHA) {
        const int s = i / 256, j = i % 256;
        if (s <= 7) v = p.b[s][j];
        else if (s == 8) v = p.b[9][j];
        else if (j < 128) v = p.b[8][j];
      } else if (i < TAIL_WRGB) {
        v = p.w[10][i - TAIL_WALPHA];
      } else if (i < TAIL_MISC) {
        const int j = (i - TAIL_WRGB) / 4, ch = (i - TAIL_WRGB) % 4;
        if (ch < 3) v = p.w[11][ch * 128 + j];
      } else {
        const int j = i - TAIL_MISC;
        if (j == 0) v = p.b[10][0];
        else if (j < 4) v = p.b[11][j - 1];
      }
      t[i] = v;
    }
  }
}

int launch_pack_net(const float* const* weights, const float* const* biases, void* packed, cudaStream_t st) {
  NetPtrs p;
  for (int i = 0; i < NSR_NET_NUM_TENSORS; ++i) {
    p.w[i] = weights[i];
    p.b[i] = biases[i];
  }
  pack_net_kernel<<<NUM_CHUNKS + 1, 256, 0, st>>>(p, static_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("pack_net_kernel");
}

// ----------------------------------------------------------------------------- the MLP kernel
constexpr int RING = 4;
constexpr int SM_ACT = 0;                         // [128 x 256] fp16, 8-row groups 4096 B apart
constexpr int SM_ENC = SM_ACT + 128 * 256 * 2;    // [128 x 64]
constexpr int SM_DIR = SM_ENC + 128 * 64 * 2;     // [128 x 64]
constexpr int SM_RING = SM_DIR + 128 * 64 * 2;    // RING x 16 KB
constexpr int SM_TAIL = SM_RING + RING * CHUNK_BYTES;
constexpr int SM_BAR = SM_TAIL + TAIL_BYTES;      // mbarriers
constexpr int SM_TOTAL = SM_BAR + 128;
constexpr int NUM_COMPUTE = 128;
constexpr int MLP_THREADS = 192;

struct MlpArgs {
  const float* rays;
  const float* z_or_pts;
  const uint8_t* packed;
  float* raw;
  int64_t n_points;
  int S;
  uint32_t flags;
  int num_tiles;
  int desc_swap;  // debug: swap the LBO / SBO fields of the shared-memory descriptors
};

__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t sbo, int swap) {
  return swap ? make_sdesc(saddr, sbo, 128, 0) : make_sdesc(saddr, 128, sbo, 0);
}

// 16-byte store of 8 fp16 (4 packed words) into a no-swizzle K-major tile whose 8-row groups are `sbo` bytes apart
__device__ __forceinline__ void st_a8(uint8_t* tile, int sbo, int row, int kgroup, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  *reinterpret_cast<uint4*>(tile + (row >> 3) * sbo + kgroup * 128 + (row & 7) * 16) = make_uint4(w0, w1, w2, w3);
}

__global__ void __launch_bounds__(MLP_THREADS, 1) nerf_mlp_kernel(MlpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAct = smem + SM_ACT;
  uint8_t* sEnc = smem + SM_ENC;
  uint8_t* sDir = smem + SM_DIR;
  uint8_t* sRing = smem + SM_RING;
  const float* sTail = reinterpret_cast<const float*>(smem + SM_TAIL);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM_BAR);  // [RING]
  uint64_t* empty = full + RING;                                // [RING]
  uint64_t* act_ready = empty + RING;                           // compute -> MMA (count 128)
  uint64_t* acc_ready = act_ready + 1;                          // MMA -> compute (count 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(act_ready, NUM_COMPUTE);
    mbar_init(acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  for (int i = tid; i < TAIL_FLOATS; i += MLP_THREADS)
    reinterpret_cast<float*>(smem + SM_TAIL)[i] = reinterpret_cast<const float*>(a.packed + WEIGHT_BYTES)[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 5) {
    // ===================================================================== weight producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        for (int c = 0; c < NUM_CHUNKS; ++c, ++it) {
          const int s = it % RING;
          if (it >= RING) mbar_wait(&empty[s], ((it / RING) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], CHUNK_BYTES);
          bulk_g2s(sRing + s * CHUNK_BYTES, a.packed + size_t(c) * CHUNK_BYTES, CHUNK_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == 4) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      const uint32_t aAct = smem_u32(sAct), aEnc = smem_u32(sEnc), aDir = smem_u32(sDir), aRing = smem_u32(sRing);
      uint32_t it = 0, act_phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        for (int step = 0; step < NUM_STEPS; ++step) {
          mbar_wait(act_ready, act_phase);
          act_phase ^= 1;
          tc_fence_after_sync();
          const int nk = step_k_chunks(step);
          for (int nh = 0; nh < step_n_halves(step); ++nh) {
            for (int kc = 0; kc < nk; ++kc, ++it) {
              const int s = it % RING;
              mbar_wait(&full[s], (it / RING) & 1);
              // A source of this K-chunk
              uint32_t abase, asbo;
              if ((step == 0 || step == 5) && kc == 0) {
                abase = aEnc;
                asbo = 1024;
              } else if (step == 9 && kc == 4) {
                abase = aDir;
                asbo = 1024;
              } else {
                const int ak = (step == 5) ? kc - 1 : kc;
                abase = aAct + ak * 1024;
                asbo = 4096;
              }
              const uint32_t bbase = aRing + s * CHUNK_BYTES;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_ss(tmem + nh * 128, kdesc(abase + j * 256, asbo, a.desc_swap), kdesc(bbase + j * 256, 1024, a.desc_swap), idesc,
                        (kc | j) != 0);
              umma_commit(&empty[s]);
            }
          }
          umma_commit(acc_ready);
        }
      }
    }
  } else {
    // ===================================================================== compute warps: encode + epilogues
    const int row = tid;  // == TMEM lane
    const uint32_t tlane = tmem + (uint32_t(warp * 32) << 16);
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int64_t p = int64_t(tile) * 128 + row;
      const bool valid = p < a.n_points;
      // ---- encode (RH:47-48): [x, sin(2^k x), cos(2^k x)]_k
      {
        float x[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t ray = p / a.S;
          const float* rp = a.rays + ray * 11;
          if (a.flags & NSR_FLAG_PTS_INPUT) {
            x[0] = a.z_or_pts[p * 3 + 0];
            x[1] = a.z_or_pts[p * 3 + 1];
            x[2] = a.z_or_pts[p * 3 + 2];
          } else {
            const float z = a.z_or_pts[p];
#pragma unroll
            for (int d = 0; d < 3; ++d) x[d] = __fadd_rn(rp[d], __fmul_rn(rp[3 + d], z));  // RN:463
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) vd[d] = rp[8 + d];
        }
        float e[64];
        e[0] = x[0];
        e[1] = x[1];
        e[2] = x[2];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn, cs;
            sincosf(x[d] * float(1 << k), &sn, &cs);
            e[3 + 6 * k + d] = sn;
            e[3 + 6 * k + 3 + d] = cs;
          }
        }
        e[63] = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          st_a8(sEnc, 1024, row, g, pack_f16x2(e[8 * g], e[8 * g + 1]), pack_f16x2(e[8 * g + 2], e[8 * g + 3]),
                pack_f16x2(e[8 * g + 4], e[8 * g + 5]), pack_f16x2(e[8 * g + 6], e[8 * g + 7]));
        float v[32];
        v[0] = vd[0];
        v[1] = vd[1];
        v[2] = vd[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn, cs;
            sincosf(vd[d] * float(1 << k), &sn, &cs);
            v[3 + 6 * k + d] = sn;
            v[3 + 6 * k + 3 + d] = cs;
          }
        }
#pragma unroll
        for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          st_a8(sDir, 1024, row, g, pack_f16x2(v[8 * g], v[8 * g + 1]), pack_f16x2(v[8 * g + 2], v[8 * g + 3]),
                pack_f16x2(v[8 * g + 4], v[8 * g + 5]), pack_f16x2(v[8 * g + 6], v[8 * g + 7]));
#pragma unroll
        for (int g = 4; g < 8; ++g) st_a8(sDir, 1024, row, g, 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(act_ready);

      float sigma = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
      for (int step = 0; step < NUM_STEPS; ++step) {
        mbar_wait(acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
        const float* bias = sTail + TAIL_BIAS + step * 256;
        const int ncols = (step == 9) ? 128 : 256;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t u[32];
          tmem_ld32(tlane + c0, u);
          tmem_ld_wait();
          float f[32];
          const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            f[4 * j + 0] = __uint_as_float(u[4 * j + 0]) + b.x;
            f[4 * j + 1] = __uint_as_float(u[4 * j + 1]) + b.y;
            f[4 * j + 2] = __uint_as_float(u[4 * j + 2]) + b.z;
            f[4 * j + 3] = __uint_as_float(u[4 * j + 3]) + b.w;
          }
          if (step == 7) {  // alpha head on the fp32 post-ReLU activations (RH:109)
            const float4* wa = reinterpret_cast<const float4*>(sTail + TAIL_WALPHA + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w = wa[j];
              sigma = fmaf(fmaxf(f[4 * j + 0], 0.f), w.x, sigma);
              sigma = fmaf(fmaxf(f[4 * j + 1], 0.f), w.y, sigma);
              sigma = fmaf(fmaxf(f[4 * j + 2], 0.f), w.z, sigma);
              sigma = fmaf(fmaxf(f[4 * j + 3], 0.f), w.w, sigma);
            }
          }
          if (step == 9) {  // rgb head (RH:117)
            const float4* wr = reinterpret_cast<const float4*>(sTail + TAIL_WRGB) + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float h = fmaxf(f[j], 0.f);
              const float4 w = wr[j];
              r0 = fmaf(h, w.x, r0);
              r1 = fmaf(h, w.y, r1);
              r2 = fmaf(h, w.z, r2);
            }
          } else if (step == 8) {  // feature_linear: no activation (RH:110)
#pragma unroll
            for (int g = 0; g < 4; ++g)
              st_a8(sAct, 4096, row, (c0 >> 3) + g, pack_f16x2(f[8 * g], f[8 * g + 1]), pack_f16x2(f[8 * g + 2], f[8 * g + 3]),
                    pack_f16x2(f[8 * g + 4], f[8 * g + 5]), pack_f16x2(f[8 * g + 6], f[8 * g + 7]));
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              st_a8(sAct, 4096, row, (c0 >> 3) + g, pack_f16x2_relu(f[8 * g], f[8 * g + 1]), pack_f16x2_relu(f[8 * g + 2], f[8 * g + 3]),
                    pack_f16x2_relu(f[8 * g + 4], f[8 * g + 5]), pack_f16x2_relu(f[8 * g + 6], f[8 * g + 7]));
          }
        }
        if (step < 9) {
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(act_ready);
        }
      }
      if (valid) {
        const float* misc = sTail + TAIL_MISC;
        reinterpret_cast<float4*>(a.raw)[p] = make_float4(r0 + misc[1], r1 + misc[2], r2 + misc[3], sigma + misc[0]);  // RH:118
      }
      tc_fence_before_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

static int g_num_sms = 0;

int launch_mlp_forward(const float* rays, const float* z_or_pts, int64_t n, int S, const void* packed, uint32_t flags,
                       float* raw, cudaStream_t st) {
  const int64_t n_points = n * S;
  if (n_points == 0) return NSR_OK;
  if (n_points > (int64_t(1) << 31) * 64) {
    set_error("mlp_forward: too many points");
    return NSR_E_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return check_launch("cudaGetDeviceProperties");
    if (prop.major != 10) {
      set_error("libnsr_b200 needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
      return NSR_E_DEVICE;
    }
    g_num_sms = prop.multiProcessorCount;
    if (cudaFuncSetAttribute(nerf_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute(nerf_mlp_kernel)");
    configured = true;
  }
  MlpArgs a;
  a.rays = rays;
  a.z_or_pts = z_or_pts;
  a.packed = static_cast<const uint8_t*>(packed);
  a.raw = raw;
  a.n_points = n_points;
  a.S = S;
  a.flags = flags;
  a.num_tiles = int((n_points + 127) / 128);
  const char* sw = getenv("NSR_DESC_SWAP");
  a.desc_swap = (sw && sw[0] == '1') ? 1 : 0;
  const int grid = a.num_tiles < g_num_sms ? a.num_tiles : g_num_sms;
  nerf_mlp_kernel<<<grid, MLP_THREADS, SM_TOTAL, st>>>(a);
  count_launch();
  return check_launch("nerf_mlp_kernel");
}

}  // namespace nsr
