// yt8m_b200 -- backward of the NetVLAD soft assignment on the tensor cores: ONE kernel replaces the recompute of the logits
// (a GEMM over all B*T frame rows that wrote z to HBM), the SIMT batched product da = X . dV (403 us of a 3.1 ms step,
// profiles/r02d_train_step_launches.txt) and the softmax-backward row kernel (yt8m_netvlad_bwd.cu, steps (3) and (4)).
//
// Per (video b, 128-frame tile), with the frames as the M rows of two accumulators in TMEM:
//   S[t, k] = sum_d x[t, d] Cw[d, k]          A = frame tile (K-major), B = Cw^T [K, D] (K-major)
//   G[t, k] = sum_d x[t, d] dV[b, d, k]       A = the same tile,        B = dV[b] [D, K] as bf16 hi + lo (MN-major: k contiguous)
// and in the epilogue, a thread per frame:  z = scale S + shift;  a = softmax_k(z);  g = G + dasum[b];
//   dz = a (g - sum_k a g) for t < num_frames, else 0;  dshift += sum_t dz;  (dz * scale) as bf16 hi / lo for the dCw GEMM.
// The frames are read once (tiles past num_frames are skipped: their dz rows stay zero from the memset); Cw and dV[b] come
// from L2 per 64-feature block.  Definition: oracle/yt8m_oracle.py:netvlad_pool (autograd); parity: tests/test_gpu_train.py.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {

constexpr int kTileF = 128;                    // frames per tile (UMMA M)
constexpr int kKb = 64;                        // features per pipeline stage (one 128-byte swizzle atom of bf16)
constexpr int kThreads = 192;                  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

template <int KC>
struct BwdCfg {
  static constexpr int kStages = KC == 64 ? 2 : 3;           // KC = 64: 81 KB -> two CTAs per SM (one's epilogue under the other's loop)
  static constexpr int kXBytes = kTileF * 128;               // 16 KB
  static constexpr int kWBytes = KC * 128;                   // Cw block / one half of dV block
  static constexpr int kStageBytes = kXBytes + 3 * kWBytes;
  static constexpr int kSmallFloats = 3 * KC;                // scale, shift, dasum[b]
  static constexpr int kTotal = kStages * kStageBytes + kSmallFloats * 4 + 128 + 1024;
};

// 32 values per lane, 32 lanes -> lane L returns sum over lanes of v[L]   (31 shuffles)
__device__ __forceinline__ float transpose_reduce32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = upper ? v[j] : v[j + off];
      const float keep = upper ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int KC>
__global__ void __launch_bounds__(kThreads, KC == 64 ? 2 : 1)
netvlad_bwd_assign_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cw,
                             const __grid_constant__ CUtensorMap tm_dv_hi, const __grid_constant__ CUtensorMap tm_dv_lo,
                             const int* __restrict__ num_frames, const float* __restrict__ scale, const float* __restrict__ shift,
                             const float* __restrict__ dasum, int T, int D, __nv_bfloat16* __restrict__ dzs_hi,
                             __nv_bfloat16* __restrict__ dzs_lo, float* __restrict__ dshift) {
  using C = BwdCfg<KC>;
  const int b = blockIdx.y, t0 = blockIdx.x * kTileF;
  const int nf = min(max(__ldg(num_frames + b), 0), T);
  if (t0 >= nf) return;                                     // nothing live in this tile: its dz rows are already zero
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* scale_s = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float* shift_s = scale_s + KC;
  float* dasum_s = shift_s + KC;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(dasum_s + KC);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* acc_full = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int nkb = D / kKb;
  constexpr uint32_t kTmemCols = 2 * KC;                    // S | G

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_cw);
    tma_prefetch_desc(&tm_dv_hi);
    tma_prefetch_desc(&tm_dv_lo);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* st = tiles + stage * C::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
        tma_load_2d(st, &tm_x, &full_bar[stage], kb * kKb, b * T + t0, kEvictFirst);          // frames: read once
        tma_load_2d(st + C::kXBytes, &tm_cw, &full_bar[stage], kb * kKb, 0, kEvictLast);      // centres: every CTA
#pragma unroll
        for (int h = 0; h < KC / 64; ++h) {
          tma_load_2d(st + C::kXBytes + C::kWBytes + h * 8192, &tm_dv_hi, &full_bar[stage], h * 64, b * D + kb * kKb, kEvictNormal);
          tma_load_2d(st + C::kXBytes + 2 * C::kWBytes + h * 8192, &tm_dv_lo, &full_bar[stage], h * 64, b * D + kb * kKb, kEvictNormal);
        }
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    constexpr uint32_t idesc_s = make_idesc_bf16(kTileF, KC, 0, 0);       // X K-major, Cw^T K-major
    constexpr uint32_t idesc_g = make_idesc_bf16(kTileF, KC, 0, 1);       // X K-major, dV MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t s_addr = smem_u32(tiles + stage * C::kStageBytes);
        const uint64_t xdesc = make_sdesc_sw128(s_addr, 16, 1024);
        const uint64_t cdesc = make_sdesc_sw128(s_addr + C::kXBytes, 16, 1024);
        // MN-major SWIZZLE_128B: 8 contraction rows per 1024-byte group (SBO), 64-cluster boxes 8192 B apart (LBO);
        // one UMMA consumes 16 contraction rows = 2048 B
        const uint64_t vhdesc = make_sdesc_sw128(s_addr + C::kXBytes + C::kWBytes, 8192, 1024);
        const uint64_t vldesc = make_sdesc_sw128(s_addr + C::kXBytes + 2 * C::kWBytes, 8192, 1024);
#pragma unroll
        for (int k = 0; k < kKb / 16; ++k) {
          const uint64_t xa = sdesc_advance(xdesc, k * 32);
          const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem_base, xa, sdesc_advance(cdesc, k * 32), idesc_s, accum);
          umma_bf16(tmem_base + KC, xa, sdesc_advance(vhdesc, k * 2048), idesc_g, accum);
          umma_bf16(tmem_base + KC, xa, sdesc_advance(vldesc, k * 2048), idesc_g, 1u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------- epilogue: a thread per frame ----------------
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;
    for (int i = et; i < KC; i += 128) {
      scale_s[i] = scale ? __ldg(scale + i) : 1.0f;
      shift_s[i] = shift ? __ldg(shift + i) : 0.0f;
      dasum_s[i] = __ldg(dasum + static_cast<long long>(b) * KC + i);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool live = t < nf;
    const uint32_t ts = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < KC; c += 32) {
      float z[32];
      tmem_ld32(ts + c, reinterpret_cast<uint32_t*>(z));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, z[j] * scale_s[c + j] + shift_s[c + j]);
    }
    float den = 0.0f, ag = 0.0f;
#pragma unroll 1
    for (int c = 0; c < KC; c += 32) {
      float z[32], g[32];
      tmem_ld32(ts + c, reinterpret_cast<uint32_t*>(z));
      tmem_ld32(ts + KC + c, reinterpret_cast<uint32_t*>(g));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = __expf(z[j] * scale_s[c + j] + shift_s[c + j] - mx);
        den += e;
        ag += e * (g[j] + dasum_s[c + j]);
      }
    }
    const float inv = 1.0f / den;
    const float s = ag * inv;                               // sum_k a_k g_k
    const bool store = t < T;
    __nv_bfloat16* oh = dzs_hi + (static_cast<long long>(b) * T + t) * KC;
    __nv_bfloat16* ol = dzs_lo + (static_cast<long long>(b) * T + t) * KC;
#pragma unroll 1
    for (int c = 0; c < KC; c += 32) {
      float z[32], g[32];
      tmem_ld32(ts + c, reinterpret_cast<uint32_t*>(z));
      tmem_ld32(ts + KC + c, reinterpret_cast<uint32_t*>(g));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = __expf(z[j] * scale_s[c + j] + shift_s[c + j] - mx) * inv;
        // select, not multiply: rows past num_frames may hold anything
        z[j] = live ? a * (g[j] + dasum_s[c + j] - s) : 0.0f;
        g[j] = z[j] * scale_s[c + j];
      }
      if (store) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 hi, lo;
          pack8_hi_lo(g + 8 * j8, hi, lo);
          *reinterpret_cast<uint4*>(oh + c + 8 * j8) = hi;
          *reinterpret_cast<uint4*>(ol + c + 8 * j8) = lo;
        }
      }
      if (dshift) {
        const float col = transpose_reduce32(z, lane);       // lane L: sum over this warp's 32 frames of dz[:, c + L]
        atomicAdd(dshift + c + lane, col);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KC>
int launch_bwd_assign_tc(const CUtensorMap& tm_x, const CUtensorMap& tm_cw, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo,
                         const int* num_frames, const float* scale, const float* shift, const float* dasum, int B, int T, int D,
                         __nv_bfloat16* hi, __nv_bfloat16* lo, float* dshift, cudaStream_t stream) {
  using C = BwdCfg<KC>;
  auto kern = netvlad_bwd_assign_tc_kernel<KC>;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_done = true;
  }
  const dim3 grid((T + kTileF - 1) / kTileF, B);
  kern<<<grid, kThreads, C::kTotal, stream>>>(tm_x, tm_cw, tm_hi, tm_lo, num_frames, scale, shift, dasum, T, D, hi, lo, dshift);
  return check_launch("netvlad_bwd_assign_tc_kernel");
}

}  // namespace

extern "C" {

int yt8m_netvlad_bwd_assign_fused_supported(int T, int D, int K) {
  return (K == 64 || K == 128) && D % 64 == 0 && D >= 64 && T > 0;
}

// x bf16 [B, T, D]; cw_packed bf16 [K, ldcw >= D] (K-major, the forward's operand); scale / shift [K] (nullable: 1 / 0);
// dv_hi / dv_lo bf16 [B, D, K] (yt8m_netvlad_bwd_norm); dasum fp32 [B, K]
// -> dzs_hi / dzs_lo bf16 [B*T, K] (every row written or zeroed), dshift fp32 [K] (nullable).
int yt8m_netvlad_bwd_assign_fused(const yt8m_bf16* x, const int* num_frames, const yt8m_bf16* cw_packed, long long ldcw,
                                  const float* scale, const float* shift, const yt8m_bf16* dv_hi, const yt8m_bf16* dv_lo,
                                  const float* dasum, int B, int T, int D, int K, yt8m_bf16* dzs_hi, yt8m_bf16* dzs_lo,
                                  float* dshift, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && num_frames && cw_packed && dv_hi && dv_lo && dasum && dzs_hi && dzs_lo, YT8M_E_BADPTR,
               "yt8m_netvlad_bwd_assign_fused: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && ldcw >= D && ldcw % 8 == 0, YT8M_E_BADSHAPE, "yt8m_netvlad_bwd_assign_fused: bad shape B=%d T=%d", B, T);
  YT8M_REQUIRE(yt8m_netvlad_bwd_assign_fused_supported(T, D, K), YT8M_E_UNSUPPORTED,
               "yt8m_netvlad_bwd_assign_fused: needs K in {64, 128} and D %% 64 == 0 (D=%d K=%d)", D, K);
  CUtensorMap tm_x, tm_cw, tm_hi, tm_lo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm_x, x, static_cast<uint64_t>(B) * T, D, D, kTileF)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_cw, cw_packed, K, D, ldcw, K)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_hi, dv_hi, static_cast<uint64_t>(B) * D, K, K, 64)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_lo, dv_lo, static_cast<uint64_t>(B) * D, K, K, 64)) != YT8M_OK) return rc;
  const size_t bytes = static_cast<size_t>(B) * T * K * sizeof(__nv_bfloat16);
  YT8M_CUDA(cudaMemsetAsync(dzs_hi, 0, bytes, stream));
  YT8M_CUDA(cudaMemsetAsync(dzs_lo, 0, bytes, stream));
  if (dshift) YT8M_CUDA(cudaMemsetAsync(dshift, 0, K * sizeof(float), stream));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dzs_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(dzs_lo);
  if (K == 64) return launch_bwd_assign_tc<64>(tm_x, tm_cw, tm_hi, tm_lo, num_frames, scale, shift, dasum, B, T, D, hi, lo, dshift, stream);
  return launch_bwd_assign_tc<128>(tm_x, tm_cw, tm_hi, tm_lo, num_frames, scale, shift, dasum, B, T, D, hi, lo, dshift, stream);
}

}  // extern "C"
