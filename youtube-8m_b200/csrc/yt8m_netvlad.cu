// yt8m_b200 -- fused NetVLAD for sm_100a: soft-assignment GEMM (B.T x D by D x K) + masked softmax over
// K + residual-aggregation GEMM (assignment^T . X) + intra-normalisation + final L2 norm, one CTA per
// video.  (NetVLAD is not part of /root/reference; definition: oracle/yt8m_oracle.py:netvlad_pool.)
//
// Data flow per video b (T frames, D features, KC clusters, all tiles SWIZZLE_128B in shared memory):
//   phase 0  S_i[128 frames, KC] = X_i[128, D] . Cw^T     i = 0..NT-1 frame tiles, accumulated in TMEM
//            while the D axis streams through a TMA ring (each Cw k-block is loaded once and shared by the
//            NT frame tiles).  X comes from HBM here -- the only HBM read of X.
//   softmax  4 warps (one TMEM lane = one frame) tcgen05.ld their S rows, apply the folded-BN affine,
//            softmax over KC in registers, zero padded frames, round to bf16 and store the assignment tile
//            to shared memory in the MN-major layout the second GEMM wants; column sums a_sum[k] via a
//            warp transpose-reduce.
//   phase 1  V^T[D, KC] = X^T[D, frames] . a[frames, KC]: X tiles are streamed AGAIN (L2 hits: the video was
//            read microseconds ago) as the MN-major A operand; accumulators for 256 TMEM columns at a
//            time (double buffered) while the epilogue warps subtract a_sum[k]*cw2[d,k], accumulate the
//            per-cluster sum of squares and stash the un-normalised descriptor.
//   rescale  after the last group: per-cluster rsqrt (intra-norm) x global rsqrt (final L2 norm) applied
//            to the stash in place (L2-resident) -> bf16 hi (+lo) / fp32 output, D-major / K-minor.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace yt8m {
int launch_netvlad_v5(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                      cudaStream_t stream);
bool netvlad_v5_supported(int T, int D, int K);
int launch_netvlad_v4(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                      cudaStream_t stream);
}

namespace {

// The stash / output of the descriptor is either bf16 hi (+ lo) or, with out_f16, one IEEE fp16 tensor.
__device__ __forceinline__ void unpack8_stash(const uint4& h, const uint4& l, int f16, float* v) {
  if (f16) {
    unpack8_f16(h, v);
  } else {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
    const uint32_t lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lv[j] << 16);
      v[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lv[j] & 0xFFFF0000u);
    }
  }
}
__device__ __forceinline__ void pack8_stash(const float* v, int f16, uint4& h, uint4& l) {
  if (f16) { h = pack8_f16(v); l = make_uint4(0, 0, 0, 0); }
  else pack8_hi_lo(v, h, l);
}


constexpr int kNvThreads = 224;            // warp 0: X producer, 1: MMA, 2-5: softmax/epilogue, 6: centre producer
constexpr int kNvSms = 148;

// debug-only phase timeline (globaltimer ns) of CTA 0: set with yt8m_debug_set_timeline()
__device__ unsigned long long* g_nv_timeline = nullptr;
__device__ int g_nv_flags = 0;     // debug ablations: 1 = no stash stores, 2 = no rescale pass, 4 = no cw2 loads, 8 = no phase-1 loads
#define NV_T(slot)                                                                         \
  do {                                                                                     \
    if (g_nv_timeline && blockIdx.x == 0 && it < 4) g_nv_timeline[it * 32 + (slot)] = global_timer_ns(); \
  } while (0)
constexpr int kNtMax = 3;                 // up to 384 frames
constexpr int kSlotBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 bf16

template <int KC>
struct NvCfg {
  static constexpr int kNBlk = (KC + 63) / 64;                 // 64-cluster blocks of an assignment tile
  static constexpr int kGM = 128 / KC > 0 ? 128 / KC : 1;      // M-blocks (128 D rows) per TMEM group (128 cols)
  static constexpr int kGroupCols = kGM * KC;                  // 128 (KC <= 128)
  static constexpr int kVBase0 = 512 - 2 * kGroupCols;         // two accumulator groups at the top of TMEM
  // X ring: every TMA op brings 128 frames x 128 features (two 64-wide SWIZZLE_128B sub-tiles, 32 KB) --
  // a thread can only issue one TMA op per ~140 ns, so the ops have to be big (tools/micro/tma_ingest.cu)
  static constexpr int kXSlotBytes = 2 * kSlotBytes;
  static constexpr int kSlots = (KC <= 64) ? 4 : 2;
  static constexpr int kCwKb = (KC <= 64) ? 2 : 1;             // k-blocks per centre stage
  static constexpr int kCwPer = 2 / kCwKb;                     // centre stages per k-block pair
  static constexpr int kCwStages = (KC <= 64) ? 2 : 3;
  static constexpr int kCwBytes = kCwKb * KC * 128;            // kCwKb x (KC rows x 64 bf16)
  static constexpr int kATileBytes = kNBlk * kSlotBytes;       // 128 frames x KC (64-wide blocks)
  static constexpr int kOffX = 0;
  static constexpr int kOffCw = kOffX + kSlots * kXSlotBytes;
  static constexpr int kOffA = kOffCw + ((kCwStages * kCwBytes + 1023) / 1024) * 1024;
  static constexpr int kOffSmall = kOffA + kNtMax * kATileBytes;
  static constexpr int kSmallBytes = 5 * KC * 4 + 512;         // scale, shift, asum, ssq, fscale + barriers
  static constexpr int kTotal = kOffSmall + kSmallBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "NetVLAD shared-memory budget exceeded");
  // S tiles (NT * KC columns from 0) may reach into the accumulator groups (KC = 128): then the next
  // video's assignment GEMM has to wait for the previous video's epilogue.
  static constexpr bool kSOverlapsV = (kNtMax * KC > kVBase0);
};

// 32 values per lane, 32 lanes -> lane L returns sum over lanes of v[L]   (31 shuffles)
__device__ __forceinline__ float warp_transpose_reduce32(float* v, int lane) {
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RingPos {
  int slot;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++slot == n) { slot = 0; phase ^= 1u; }
  }
};

// PERSISTENT: CTA c handles videos c, c + gridDim.x, ...  The TMA producer and the MMA issuer run ahead
// into the next video's assignment GEMM (HBM reads) while the 4 softmax/epilogue warps are still
// subtracting residuals / rescaling the previous video, so the HBM stream never waits for the epilogue.
template <int KC>
__global__ void __launch_bounds__(kNvThreads, 1)
netvlad_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cw,
                     const int* __restrict__ num_frames, int B, int T, int D, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ cw2, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ld_out, int out_f16,
                     float* __restrict__ stats) {
  using C = NvCfg<KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem + C::kOffX;
  uint8_t* cws = smem + C::kOffCw;
  uint8_t* atile = smem + C::kOffA;
  float* scale_s = reinterpret_cast<float*>(smem + C::kOffSmall);
  float* shift_s = scale_s + KC;
  float* asum_s = shift_s + KC;
  float* ssq_s = asum_s + KC;
  float* fscale_s = ssq_s + KC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fscale_s + KC);
  uint64_t* cw_full = bars;                       // [3]
  uint64_t* cw_empty = cw_full + 3;               // [3]
  uint64_t* x_full = cw_empty + 3;                // [8]  one ring for both phases
  uint64_t* x_empty = x_full + 8;                 // [8]
  uint64_t* s_full = x_empty + 8;                 // [1]  assignment accumulators complete (per video)
  uint64_t* a_ready = s_full + 1;                 // [1]  assignment tiles written (per video)
  uint64_t* v_full = a_ready + 1;                 // [2]  accumulator group complete
  uint64_t* v_empty = v_full + 2;                 // [2]  accumulator group drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + 2);
  float* total_s = reinterpret_cast<float*>(tmem_slot + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int NT = (T + 127) / 128;
  const int NKB = D / 64;
  const int NMB = D / 128;
  const int NG = (NMB + C::kGM - 1) / C::kGM;
  const int n_iter = (B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_cw);
    for (int i = 0; i < 3; ++i) { mbar_init(&cw_full[i], 1); mbar_init(&cw_empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(a_ready, 4);
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int k = threadIdx.x; k < KC; k += kNvThreads) {
    scale_s[k] = scale ? scale[k] : 1.0f;
    shift_s[k] = shift ? shift[k] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =================================== X producer ===================================
    RingPos xr{0, 0};
    for (int it = 0; it < n_iter; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      if (lane == 0) NV_T(0);
      // ---- phase 0: two 64-wide k-blocks of 128 frames per op; HBM reads
      for (int kbp = 0; kbp < NKB / 2; ++kbp)
        for (int i = 0; i < NT; ++i) {
          mbar_wait(&x_empty[xr.slot], xr.phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&x_full[xr.slot], C::kXSlotBytes);
            tma_load_4d(xs + xr.slot * C::kXSlotBytes, &tm_x, &x_full[xr.slot], 0, i * 128, 2 * kbp, b, kEvictNormal);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
      if (lane == 0) NV_T(1);
      // ---- phase 1: the same video again (L2 hits), 128 frames x one 128-row M-block of D per op
      for (int g = 0; g < NG; ++g)
        for (int i = 0; i < NT; ++i)
          for (int ml = 0; ml < C::kGM; ++ml) {
            const int m = g * C::kGM + ml;
            if (m >= NMB) break;
            mbar_wait(&x_empty[xr.slot], xr.phase ^ 1u);
            if (elect_one()) {
              mbar_arrive_expect_tx(&x_full[xr.slot], C::kXSlotBytes);
              tma_load_4d(xs + xr.slot * C::kXSlotBytes, &tm_x, &x_full[xr.slot], 0, i * 128, 2 * m, b, kEvictFirst);
            }
            __syncwarp();
            xr.advance(C::kSlots);
          }
      if (lane == 0) NV_T(2);
    }
  } else if (warp == 6) {
    // =================================== centre (Cw) producer ===================================
    RingPos cr{0, 0};
    for (int it = 0; it < n_iter; ++it)
      for (int kc = 0; kc < NKB / C::kCwKb; ++kc) {
        mbar_wait(&cw_empty[cr.slot], cr.phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&cw_full[cr.slot], C::kCwBytes);
          tma_load_3d(cws + cr.slot * C::kCwBytes, &tm_cw, &cw_full[cr.slot], 0, 0, kc * C::kCwKb, kEvictLast);
        }
        __syncwarp();
        cr.advance(C::kCwStages);
      }
  } else if (warp == 1) {
    // =================================== MMA issuer =====================================
    constexpr uint32_t idesc0 = make_idesc_bf16(128, KC, 0, 0);     // S = X . Cw^T      (both K-major)
    constexpr uint32_t idesc1 = make_idesc_bf16(128, KC, 1, 1);     // V^T = X^T . a     (both MN-major)
    RingPos xr{0, 0}, cr{0, 0};
    int gidx = 0;                                                   // accumulator groups issued so far
    for (int it = 0; it < n_iter; ++it) {
      if (C::kSOverlapsV && gidx >= 1) {
        // the S tiles reach into the accumulator groups: wait until the previous video's groups are drained
        const int g1 = gidx - 1;
        mbar_wait(&v_empty[g1 & 1], (g1 >> 1) & 1);
        if (gidx >= 2) { const int g2 = gidx - 2; mbar_wait(&v_empty[g2 & 1], (g2 >> 1) & 1); }
        tc_fence_after();
      }
      if (lane == 0) NV_T(8);
      for (int kbp = 0; kbp < NKB / 2; ++kbp) {
        // centre tiles of this k-block pair: one stage (two k-blocks) or two stages (one each)
        uint32_t cw_addr[2];
        int cw_slot[2];
#pragma unroll
        for (int j = 0; j < C::kCwPer; ++j) {
          mbar_wait(&cw_full[cr.slot], cr.phase);
          cw_slot[j] = cr.slot;
          cw_addr[j] = smem_u32(cws + cr.slot * C::kCwBytes);
          cr.advance(C::kCwStages);
        }
        if (C::kCwPer == 1) { cw_addr[1] = cw_addr[0] + KC * 128; cw_slot[1] = cw_slot[0]; }
        for (int i = 0; i < NT; ++i) {
          mbar_wait(&x_full[xr.slot], xr.phase);
          tc_fence_after();
          if (lane == 0 && kbp == 0 && i == 0) NV_T(9);
          if (elect_one()) {
            const uint32_t x_addr = smem_u32(xs + xr.slot * C::kXSlotBytes);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t adesc0 = make_sdesc_sw128(x_addr + j * kSlotBytes, 16, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(cw_addr[j], 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + i * KC, sdesc_advance(adesc0, k * 32), sdesc_advance(bdesc0, k * 32), idesc0,
                          (kbp > 0 || j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&x_empty[xr.slot]);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
        if (elect_one()) {
          umma_commit(&cw_empty[cw_slot[0]]);
          if (C::kCwPer == 2) umma_commit(&cw_empty[cw_slot[1]]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(s_full);
      if (lane == 0) NV_T(10);
      __syncwarp();
      // phase 1 needs every assignment tile of this video
      mbar_wait(a_ready, it & 1);
      tc_fence_after();
      if (lane == 0) NV_T(11);
      for (int g = 0; g < NG; ++g, ++gidx) {
        const int buf = gidx & 1;
        if (gidx >= 2) {
          mbar_wait(&v_empty[buf], ((gidx >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int i = 0; i < NT; ++i) {
          const int valid = min(128, T - i * 128);
          const int nsteps = (valid + 15) >> 4;
          for (int ml = 0; ml < C::kGM; ++ml) {
            const int m = g * C::kGM + ml;
            if (m >= NMB) break;
            mbar_wait(&x_full[xr.slot], xr.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc0 = make_sdesc_sw128(smem_u32(xs + xr.slot * C::kXSlotBytes), kSlotBytes, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(atile + i * C::kATileBytes), kSlotBytes, 1024);
              const uint32_t dcol = tmem_base + C::kVBase0 + buf * C::kGroupCols + ml * KC;
              for (int s = 0; s < nsteps; ++s)
                umma_bf16(dcol, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1,
                          (i > 0 || s > 0) ? 1u : 0u);
              umma_commit(&x_empty[xr.slot]);
            }
            __syncwarp();
            xr.advance(C::kSlots);
          }
        }
        if (elect_one()) umma_commit(&v_full[buf]);
        if (lane == 0 && g < 5) NV_T(g < 3 ? 13 + g : (g == 3 ? 25 : 28));
        if (lane == 0 && g == NG - 1) NV_T(12);
        __syncwarp();
      }
    }
  } else {
    // ============================ softmax + epilogue warps (128 threads) ============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;                       // 0..127 (warps 2..5)
    const int dbg_flags = g_nv_flags;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int gidx = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const int nf = min(max(num_frames[b], 0), T);
      if (et < KC) { asum_s[et] = 0.0f; ssq_s[et] = 0.0f; }
      if (et == 0) *total_s = 0.0f;
      named_bar_sync(1, 128);
      if (et == 0) NV_T(16);
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      if (et == 0) NV_T(17);
      constexpr bool kRegAcc = (KC <= 64);                // running sums live in registers, reduced once per video
      float acc[kRegAcc ? KC : 1];
      if (kRegAcc) {
#pragma unroll
        for (int k = 0; k < (kRegAcc ? KC : 1); ++k) acc[k] = 0.0f;
      }
      for (int i = 0; i < NT; ++i) {
        float l[KC];
#pragma unroll
        for (int c = 0; c < KC; c += 32) tmem_ld32(taddr + i * KC + c, reinterpret_cast<uint32_t*>(l) + c);
        tmem_ld_wait();
        const bool valid = (i * 128 + row) < nf;
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          l[k] = l[k] * scale_s[k] + shift_s[k];
          mx = fmaxf(mx, l[k]);
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          l[k] = __expf(l[k] - mx);
          sum += l[k];
        }
        const float inv = valid ? 1.0f / sum : 0.0f;
        uint8_t* at = atile + i * C::kATileBytes;
#pragma unroll
        for (int c8 = 0; c8 < KC / 8; ++c8) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(l[c8 * 8 + 2 * j] * inv);
            const __nv_bfloat16 h1 = __float2bfloat16_rn(l[c8 * 8 + 2 * j + 1] * inv);
            l[c8 * 8 + 2 * j] = __bfloat162float(h0);         // a_sum uses the rounded assignment too
            l[c8 * 8 + 2 * j + 1] = __bfloat162float(h1);
            w[j] = pack_bf16x2(h0, h1);
          }
          const int nb = c8 >> 3;
          *reinterpret_cast<uint4*>(at + nb * kSlotBytes + sw128_offset(row, c8 & 7)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (kRegAcc) {
#pragma unroll
          for (int k = 0; k < (kRegAcc ? KC : 1); ++k) acc[k] += l[k];
        } else {
#pragma unroll
          for (int c = 0; c < KC; c += 32) {
            const float tot = warp_transpose_reduce32(l + c, lane);
            atomicAdd(&asum_s[c + lane], tot);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (kRegAcc) {
#pragma unroll
        for (int c = 0; c < (kRegAcc ? KC : 32); c += 32) {
          const float tot = warp_transpose_reduce32(acc + c, lane);
          atomicAdd(&asum_s[c + lane], tot);
        }
      }
      named_bar_sync(1, 128);                               // a_sum complete
      if (et == 0) NV_T(18);

      __nv_bfloat16* ohi = out_hi + static_cast<long long>(b) * ld_out;
      __nv_bfloat16* olo = out_lo ? out_lo + static_cast<long long>(b) * ld_out : nullptr;
      if (kRegAcc) {
#pragma unroll
        for (int k = 0; k < (kRegAcc ? KC : 1); ++k) acc[k] = 0.0f;          // now: running sum of squares
      }
      constexpr int kChunks = KC / 32;
      float4 cc[8];                                            // residual centres of the NEXT chunk (prefetch)
      {
        const float* c2 = cw2 + static_cast<long long>(row) * KC;            // M-block 0, chunk 0
#pragma unroll
        for (int j = 0; j < 8; ++j) cc[j] = __ldg(reinterpret_cast<const float4*>(c2) + j);
      }
      for (int g = 0; g < NG; ++g, ++gidx) {
        const int buf = gidx & 1;
        mbar_wait(&v_full[buf], (gidx >> 1) & 1);
        tc_fence_after();
        if (et == 0 && g < 6) NV_T(19 + g);
        for (int ml = 0; ml < C::kGM; ++ml) {
          const int m = g * C::kGM + ml;
          if (m >= NMB) break;
          const int d = m * 128 + row;
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            const int c = ch * 32;
            float v[32];
            tmem_ld32(taddr + C::kVBase0 + buf * C::kGroupCols + ml * KC + c, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] -= asum_s[c + 4 * j] * cc[j].x;
              v[4 * j + 1] -= asum_s[c + 4 * j + 1] * cc[j].y;
              v[4 * j + 2] -= asum_s[c + 4 * j + 2] * cc[j].z;
              v[4 * j + 3] -= asum_s[c + 4 * j + 3] * cc[j].w;
            }
            {  // prefetch the centres of the next chunk (next chunk of this row, or the next M-block's row)
              const bool last_chunk = (ch == kChunks - 1);
              const int nm = last_chunk ? m + 1 : m;
              const int ncol = last_chunk ? 0 : c + 32;
              if (nm < NMB && !(dbg_flags & 4)) {
                const float* c2n = cw2 + static_cast<long long>(nm * 128 + row) * KC + ncol;
#pragma unroll
                for (int j = 0; j < 8; ++j) cc[j] = __ldg(reinterpret_cast<const float4*>(c2n) + j);
              }
            }
            // stash un-normalised (bf16 hi [+ lo]); keep it in L2 until the rescale pass reads it back
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint4 hi, lo;
              pack8_stash(v + 8 * j8, out_f16, hi, lo);
              const long long o = static_cast<long long>(d) * KC + c + 8 * j8;
              if (!(dbg_flags & 1)) {
                st_global_hint(ohi + o, hi, kEvictLast);
                if (olo) st_global_hint(olo + o, lo, kEvictLast);
              }
            }
            if (kRegAcc) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[(kRegAcc ? c : 0) + (kRegAcc ? j : 0)] += v[j] * v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] * v[j];
              const float tot = warp_transpose_reduce32(v, lane);
              atomicAdd(&ssq_s[c + lane], tot);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_empty[buf]);
        if (et == 0 && g < 5) NV_T(3 + g);
      }
      if (kRegAcc) {
#pragma unroll
        for (int c = 0; c < (kRegAcc ? KC : 32); c += 32) {
          const float tot = warp_transpose_reduce32(acc + c, lane);
          atomicAdd(&ssq_s[c + lane], tot);
        }
      }
      // ------------------------------- rescale in place -------------------------------
      __threadfence_block();
      named_bar_sync(1, 128);
      if (et == 0) NV_T(26);
      if (et < KC) {
        const float ss = ssq_s[et];
        const float rs = rsqrtf(fmaxf(ss, 1e-12f));
        fscale_s[et] = rs;
        atomicAdd(total_s, ss * rs * rs);
        if (stats) {                                           // saved for the backward pass: a_sum, ||V_k||^2
          stats[static_cast<long long>(b) * (2 * KC + 1) + et] = asum_s[et];
          stats[static_cast<long long>(b) * (2 * KC + 1) + KC + et] = ss;
        }
      }
      named_bar_sync(1, 128);
      const float gs = rsqrtf(fmaxf(*total_s, 1e-12f));
      if (stats && et == 0) stats[static_cast<long long>(b) * (2 * KC + 1) + 2 * KC] = *total_s;
      const long long n = static_cast<long long>(D) * KC;
      float* of = out_f32 ? out_f32 + static_cast<long long>(b) * ld_out : nullptr;
      constexpr int kU = 8;                                 // independent 16-byte loads in flight per thread (x2 with lo)
      for (long long e0 = static_cast<long long>(et) * 8; e0 < ((dbg_flags & 2) ? 0 : n); e0 += 128 * 8 * kU) {
        uint4 h[kU], lw[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long long e = e0 + static_cast<long long>(u) * 128 * 8;
          h[u] = make_uint4(0, 0, 0, 0);
          lw[u] = make_uint4(0, 0, 0, 0);
          if (e < n) {
            h[u] = ld_global_hint(ohi + e, kEvictFirst);
            if (olo) lw[u] = ld_global_hint(olo + e, kEvictFirst);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long long e = e0 + static_cast<long long>(u) * 128 * 8;
          if (e >= n) break;
          const int k0 = static_cast<int>(e % KC);
          float v[8];
          unpack8_stash(h[u], lw[u], out_f16, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= fscale_s[k0 + j] * gs;
          uint4 nh, nl;
          pack8_stash(v, out_f16, nh, nl);
          *reinterpret_cast<uint4*>(ohi + e) = nh;
          if (olo) *reinterpret_cast<uint4*>(olo + e) = nl;
          if (of) {
            *reinterpret_cast<float4*>(of + e) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(of + e + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      named_bar_sync(1, 128);                               // fscale_s / total_s are reused by the next video
      if (et == 0) NV_T(27);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================================
// NetVLAD v3 (K = 64): same pipeline, but nothing on the epilogue side touches global memory with
// per-thread strided accesses any more (measured: 7.6 us per accumulator group, 24 us per rescale --
// 16-byte pieces of 32 different lines per instruction):
//   * the residual  - a_sum[k] * cw2[d, k]  is folded into the tensor pipe:  V^T += cw2_tile . (-diag(a_sum)),
//     with cw2 (bf16 hi + lo, [D, K] row-major = K-major A operand) streamed by TMA through the X ring and
//     -diag(a_sum) (bf16 hi + lo, 64 x 64) written to shared memory by the softmax warps;
//   * the un-normalised descriptor is written through a SWIZZLE_128B staging tile with TMA tensor stores
//     and read back / re-written the same way in the rescale pass.
// =================================================================================================
struct Nv3Cfg {
  static constexpr int KC = 64;
  static constexpr int kGM = 2;                                // M-blocks per accumulator group (128 TMEM columns)
  static constexpr int kGroupCols = 128;
  static constexpr int kVBase0 = 256;
  static constexpr int kXSlotBytes = 2 * kSlotBytes;           // 32 KB
  static constexpr int kSlots = 3;
  static constexpr int kCwStages = 2;
  static constexpr int kCwBytes = 2 * KC * 128;                // two k-blocks
  static constexpr int kATileBytes = kSlotBytes;               // 128 frames x 64 clusters
  static constexpr int kDiagBytes = 64 * 128;                  // 64 x 64 bf16, K-major SW128
  static constexpr int kOffX = 0;
  static constexpr int kOffCw = kOffX + kSlots * kXSlotBytes;             //  96 KB
  static constexpr int kOffA = kOffCw + kCwStages * kCwBytes;             // 128 KB
  static constexpr int kOffDiag = kOffA + kNtMax * kATileBytes;           // 176 KB
  static constexpr int kOffStage = kOffDiag + 2 * kDiagBytes;             // 192 KB
  static constexpr int kOffSmall = kOffStage + 2 * kSlotBytes;            // 224 KB
  static constexpr int kSmallBytes = 6 * KC * 4 + 512;         // scale, shift, asum, ssq, fscale[2] + barriers
  static constexpr int kTotal = kOffSmall + kSmallBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "NetVLAD v3 shared-memory budget exceeded");
};

// Warpgroups: WG0 = warps 0-3 (X producer, MMA issuer, centre producer, idle), WG1 = warps 4-7 (softmax +
// accumulator epilogue, one TMEM lane quadrant each), WG2 = warps 8-11 (final rescale of the PREVIOUS video, off
// the critical path).  setmaxnreg hands WG1 the registers the other two do not need.
constexpr int kNv3Threads = 384;

__global__ void __launch_bounds__(kNv3Threads, 1)
netvlad_v3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x64,
                  const __grid_constant__ CUtensorMap tm_cw,
                  const __grid_constant__ CUtensorMap tm_c2_hi, const __grid_constant__ CUtensorMap tm_c2_lo,
                  const __grid_constant__ CUtensorMap tm_out_hi, const __grid_constant__ CUtensorMap tm_out_lo,
                  const int* __restrict__ num_frames, int B, int T, int D, const float* __restrict__ scale,
                  const float* __restrict__ shift, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                  __nv_bfloat16* __restrict__ out_lo, long long ld_out, int out_f16, float* __restrict__ stats) {
  const int want_lo = out_lo != nullptr;
  using C = Nv3Cfg;
  constexpr int KC = C::KC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem + C::kOffX;
  uint8_t* cws = smem + C::kOffCw;
  uint8_t* atile = smem + C::kOffA;
  uint8_t* diag_hi = smem + C::kOffDiag;
  uint8_t* diag_lo = diag_hi + C::kDiagBytes;
  uint8_t* stage_hi = smem + C::kOffStage;
  uint8_t* stage_lo = stage_hi + kSlotBytes;
  float* scale_s = reinterpret_cast<float*>(smem + C::kOffSmall);
  float* shift_s = scale_s + KC;
  float* asum_s = shift_s + KC;
  float* ssq_s = asum_s + KC;
  float* fscale_s = ssq_s + KC;                   // [2][KC]: final per-cluster scale of video parity p (intra-norm x global norm)
  uint64_t* bars = reinterpret_cast<uint64_t*>(fscale_s + 2 * KC);
  uint64_t* cw_full = bars;                       // [2]
  uint64_t* cw_empty = cw_full + 2;               // [2]
  uint64_t* x_full = cw_empty + 2;                // [3]
  uint64_t* x_empty = x_full + 3;                 // [3]
  uint64_t* s_full = x_empty + 3;
  uint64_t* a_ready = s_full + 1;
  uint64_t* v_full = a_ready + 1;                 // [2]
  uint64_t* v_empty = v_full + 2;                 // [2]
  uint64_t* resc_go = v_empty + 2;                // [2] WG1 -> WG2: stash of video parity p complete, fscale_s[p] valid
  uint64_t* resc_done = resc_go + 2;              // [2] WG2 -> WG1: fscale_s[p] may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resc_done + 2);
  float* total_s = reinterpret_cast<float*>(tmem_slot + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int NT = (T + 127) / 128;
  const int NKB = D / 64;
  const int NMB = D / 128;
  const int NG = (NMB + C::kGM - 1) / C::kGM;
  const int n_iter = (B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // the last frame tile is loaded with a 64-row box when it holds <= 64 frames (T = 300: 44): 17 % less X traffic
  const bool short_last = (T - (NT - 1) * 128) <= 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_cw); tma_prefetch_desc(&tm_c2_hi); tma_prefetch_desc(&tm_c2_lo);
    tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    for (int i = 0; i < 2; ++i) { mbar_init(&cw_full[i], 1); mbar_init(&cw_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(a_ready, 4);
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&resc_go[i], 1); mbar_init(&resc_done[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int k = threadIdx.x; k < KC; k += kNv3Threads) {
    scale_s[k] = scale ? scale[k] : 1.0f;
    shift_s[k] = shift ? shift[k] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // =================================== X / cw2 producer ===================================
    RingPos xr{0, 0};
    for (int it = 0; it < n_iter; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      if (lane == 0) NV_T(0);
      for (int kbp = 0; kbp < NKB / 2; ++kbp)
        for (int i = 0; i < NT; ++i) {
          mbar_wait(&x_empty[xr.slot], xr.phase ^ 1u);
          if (elect_one()) {
            const bool sh = short_last && i == NT - 1;
            mbar_arrive_expect_tx(&x_full[xr.slot], sh ? C::kXSlotBytes / 2 : C::kXSlotBytes);
            tma_load_4d(xs + xr.slot * C::kXSlotBytes, sh ? &tm_x64 : &tm_x, &x_full[xr.slot], 0, i * 128, 2 * kbp, b, kEvictNormal);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
      if (lane == 0) NV_T(1);
      for (int g = 0; g < NG; ++g) {
        for (int i = 0; i < NT; ++i)
          for (int ml = 0; ml < C::kGM; ++ml) {
            const int m = g * C::kGM + ml;
            if (m >= NMB) break;
            mbar_wait(&x_empty[xr.slot], xr.phase ^ 1u);
            if (elect_one()) {
              const bool sh = short_last && i == NT - 1;
              mbar_arrive_expect_tx(&x_full[xr.slot], sh ? C::kXSlotBytes / 2 : C::kXSlotBytes);
              tma_load_4d(xs + xr.slot * C::kXSlotBytes, sh ? &tm_x64 : &tm_x, &x_full[xr.slot], 0, i * 128, 2 * m, b, kEvictFirst);
            }
            __syncwarp();
            xr.advance(C::kSlots);
          }
        // residual centres of this group's M-blocks: cw2 rows [m*128, +128) as bf16 hi | lo
        for (int ml = 0; ml < C::kGM; ++ml) {
          const int m = g * C::kGM + ml;
          if (m >= NMB) break;
          mbar_wait(&x_empty[xr.slot], xr.phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&x_full[xr.slot], C::kXSlotBytes);
            tma_load_2d(xs + xr.slot * C::kXSlotBytes, &tm_c2_hi, &x_full[xr.slot], 0, m * 128, kEvictLast);
            tma_load_2d(xs + xr.slot * C::kXSlotBytes + kSlotBytes, &tm_c2_lo, &x_full[xr.slot], 0, m * 128, kEvictLast);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
      }
      if (lane == 0) NV_T(2);
    }
  } else if (warp == 2) {
    // =================================== assignment-centre (Cw) producer ===================================
    RingPos cr{0, 0};
    for (int it = 0; it < n_iter; ++it)
      for (int kc = 0; kc < NKB / 2; ++kc) {
        mbar_wait(&cw_empty[cr.slot], cr.phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&cw_full[cr.slot], C::kCwBytes);
          tma_load_3d(cws + cr.slot * C::kCwBytes, &tm_cw, &cw_full[cr.slot], 0, 0, kc * 2, kEvictLast);
        }
        __syncwarp();
        cr.advance(C::kCwStages);
      }
  } else if (warp == 1) {
    // =================================== MMA issuer =====================================
    constexpr uint32_t idesc0 = make_idesc_bf16(128, KC, 0, 0);     // K-major x K-major
    constexpr uint32_t idesc1 = make_idesc_bf16(128, KC, 1, 1);     // MN-major x MN-major
    RingPos xr{0, 0}, cr{0, 0};
    int gidx = 0;
    for (int it = 0; it < n_iter; ++it) {
      if (lane == 0) NV_T(8);
      for (int kbp = 0; kbp < NKB / 2; ++kbp) {
        mbar_wait(&cw_full[cr.slot], cr.phase);
        const uint32_t cw_addr = smem_u32(cws + cr.slot * C::kCwBytes);
        for (int i = 0; i < NT; ++i) {
          mbar_wait(&x_full[xr.slot], xr.phase);
          tc_fence_after();
          if (lane == 0 && kbp == 0 && i == 0) NV_T(9);
          if (elect_one()) {
            const uint32_t x_addr = smem_u32(xs + xr.slot * C::kXSlotBytes);
            // short tile: the two 64-wide sub-tiles are 64 rows (8 KB) each; MMA rows 64..127 read the neighbouring
            // sub-tile (finite garbage -> S rows of frames >= T, which the softmax forces to zero)
            const uint32_t sub = (short_last && i == NT - 1) ? kSlotBytes / 2 : kSlotBytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t adesc0 = make_sdesc_sw128(x_addr + j * sub, 16, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(cw_addr + j * KC * 128, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + i * KC, sdesc_advance(adesc0, k * 32), sdesc_advance(bdesc0, k * 32), idesc0,
                          (kbp > 0 || j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&x_empty[xr.slot]);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
        if (elect_one()) umma_commit(&cw_empty[cr.slot]);
        __syncwarp();
        cr.advance(C::kCwStages);
      }
      if (elect_one()) umma_commit(s_full);
      if (lane == 0) NV_T(10);
      __syncwarp();
      mbar_wait(a_ready, it & 1);                                   // assignment tiles + -diag(a_sum) are in shared memory
      tc_fence_after();
      if (lane == 0) NV_T(11);
      for (int g = 0; g < NG; ++g, ++gidx) {
        const int buf = gidx & 1;
        if (gidx >= 2) {
          mbar_wait(&v_empty[buf], ((gidx >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int i = 0; i < NT; ++i) {
          const int valid = min(128, T - i * 128);
          const int nsteps = (valid + 15) >> 4;
          for (int ml = 0; ml < C::kGM; ++ml) {
            const int m = g * C::kGM + ml;
            if (m >= NMB) break;
            mbar_wait(&x_full[xr.slot], xr.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sub = (short_last && i == NT - 1) ? kSlotBytes / 2 : kSlotBytes;
              const uint64_t adesc0 = make_sdesc_sw128(smem_u32(xs + xr.slot * C::kXSlotBytes), sub, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(atile + i * C::kATileBytes), kSlotBytes, 1024);
              const uint32_t dcol = tmem_base + C::kVBase0 + buf * C::kGroupCols + ml * KC;
              for (int s = 0; s < nsteps; ++s)
                umma_bf16(dcol, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1, (i > 0 || s > 0) ? 1u : 0u);
              umma_commit(&x_empty[xr.slot]);
            }
            __syncwarp();
            xr.advance(C::kSlots);
          }
        }
        // residual:  V^T[d, k] -= a_sum[k] * cw2[d, k]   as   (cw2_hi + cw2_lo) . (-diag_hi - diag_lo), lo*lo dropped
        for (int ml = 0; ml < C::kGM; ++ml) {
          const int m = g * C::kGM + ml;
          if (m >= NMB) break;
          mbar_wait(&x_full[xr.slot], xr.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t c_addr = smem_u32(xs + xr.slot * C::kXSlotBytes);
            const uint64_t chi = make_sdesc_sw128(c_addr, 16, 1024), clo = make_sdesc_sw128(c_addr + kSlotBytes, 16, 1024);
            const uint64_t dhi = make_sdesc_sw128(smem_u32(diag_hi), 16, 1024), dlo = make_sdesc_sw128(smem_u32(diag_lo), 16, 1024);
            const uint32_t dcol = tmem_base + C::kVBase0 + buf * C::kGroupCols + ml * KC;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(dcol, sdesc_advance(chi, k * 32), sdesc_advance(dhi, k * 32), idesc0, 1u);
              umma_bf16(dcol, sdesc_advance(clo, k * 32), sdesc_advance(dhi, k * 32), idesc0, 1u);
              umma_bf16(dcol, sdesc_advance(chi, k * 32), sdesc_advance(dlo, k * 32), idesc0, 1u);
            }
            umma_commit(&x_empty[xr.slot]);
          }
          __syncwarp();
          xr.advance(C::kSlots);
        }
        if (elect_one()) umma_commit(&v_full[buf]);
        if (lane == 0 && g < 5) NV_T(g < 3 ? 13 + g : (g == 3 ? 25 : 28));
        if (lane == 0 && g == NG - 1) NV_T(12);
        __syncwarp();
      }
    }
  }
  } else if (warp < 8) {
    setmaxnreg_inc<232>();
    // ============================ WG1: softmax + accumulator epilogue (128 threads) ============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;                      // 0..127 (warps 4..7)
    const bool boss = (warp == 4);                          // warp whose elected lane owns the TMA store groups
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int gidx = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const int nf = min(max(num_frames[b], 0), T);
      if (et < KC) { asum_s[et] = 0.0f; ssq_s[et] = 0.0f; }
      if (et == 0) *total_s = 0.0f;
      named_bar_sync(1, 128);
      if (et == 0) NV_T(16);
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      if (et == 0) NV_T(17);
      float acc[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) acc[k] = 0.0f;
      for (int i = 0; i < NT; ++i) {
        float l[KC];
#pragma unroll
        for (int c = 0; c < KC; c += 32) tmem_ld32(taddr + i * KC + c, reinterpret_cast<uint32_t*>(l) + c);
        tmem_ld_wait();
        const bool valid = (i * 128 + row) < nf;
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          l[k] = l[k] * scale_s[k] + shift_s[k];
          mx = fmaxf(mx, l[k]);
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          l[k] = __expf(l[k] - mx);
          sum += l[k];
        }
        const float inv = 1.0f / sum;
        uint8_t* at = atile + i * C::kATileBytes;
#pragma unroll
        for (int c8 = 0; c8 < KC / 8; ++c8) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // select, not multiply: rows of frames >= num_frames may hold non-finite garbage
            const __nv_bfloat16 h0 = __float2bfloat16_rn(valid ? l[c8 * 8 + 2 * j] * inv : 0.0f);
            const __nv_bfloat16 h1 = __float2bfloat16_rn(valid ? l[c8 * 8 + 2 * j + 1] * inv : 0.0f);
            acc[c8 * 8 + 2 * j] += __bfloat162float(h0);      // a_sum uses the rounded assignment too
            acc[c8 * 8 + 2 * j + 1] += __bfloat162float(h1);
            w[j] = pack_bf16x2(h0, h1);
          }
          *reinterpret_cast<uint4*>(at + sw128_offset(row, c8)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
#pragma unroll
      for (int c = 0; c < KC; c += 32) {
        const float tot = warp_transpose_reduce32(acc + c, lane);
        atomicAdd(&asum_s[c + lane], tot);
      }
      named_bar_sync(1, 128);                               // a_sum complete
      if (et < KC) {
        // row n = et of -diag(a_sum): 64 bf16, zero except element n; K-major SWIZZLE_128B tile (hi and lo)
        __nv_bfloat16 h, l2;
        split_bf16(-asum_s[et], h, l2);
        const uint32_t pos = (et & 1) ? 16u : 0u;           // element n sits in chunk n/8, 32-bit word (n%8)/2
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t wh[4] = {0, 0, 0, 0}, wl[4] = {0, 0, 0, 0};
          if (c8 == (et >> 3)) {
            wh[(et & 7) >> 1] = static_cast<uint32_t>(__bfloat16_as_ushort(h)) << pos;
            wl[(et & 7) >> 1] = static_cast<uint32_t>(__bfloat16_as_ushort(l2)) << pos;
          }
          *reinterpret_cast<uint4*>(diag_hi + sw128_offset(et, c8)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          *reinterpret_cast<uint4*>(diag_lo + sw128_offset(et, c8)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (et == 0) NV_T(18);

      // ------------------------------- accumulator groups -------------------------------
#pragma unroll
      for (int k = 0; k < KC; ++k) acc[k] = 0.0f;            // now: running sum of squares per cluster
      for (int g = 0; g < NG; ++g, ++gidx) {
        const int buf = gidx & 1;
        mbar_wait(&v_full[buf], (gidx >> 1) & 1);
        tc_fence_after();
        if (et == 0 && g < 6) NV_T(19 + g);
        for (int ml = 0; ml < C::kGM; ++ml) {
          const int m = g * C::kGM + ml;
          if (m >= NMB) break;
          // staging tile free?  (the previous TMA store has finished reading it)
          if (boss && elect_one()) bulk_wait_group_read0();
          named_bar_sync(1, 128);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tmem_ld32(taddr + C::kVBase0 + buf * C::kGroupCols + ml * KC + ch * 32, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[ch * 32 + j] += v[j] * v[j];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint4 hi, lo;
              pack8_stash(v + 8 * j8, out_f16, hi, lo);
              *reinterpret_cast<uint4*>(stage_hi + sw128_offset(row, ch * 4 + j8)) = hi;
              if (want_lo) *reinterpret_cast<uint4*>(stage_lo + sw128_offset(row, ch * 4 + j8)) = lo;
            }
          }
          if (ml == C::kGM - 1 || m == NMB - 1) {            // last TMEM read of this group: release the accumulators
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&v_empty[buf]);
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (boss && elect_one()) {
            tma_store_2d(&tm_out_hi, stage_hi, 0, b * D + m * 128);
            if (want_lo) tma_store_2d(&tm_out_lo, stage_lo, 0, b * D + m * 128);
            bulk_commit_group();
          }
          __syncwarp();
        }
        if (et == 0 && g < 5) NV_T(3 + g);
      }
#pragma unroll
      for (int c = 0; c < KC; c += 32) {
        const float tot = warp_transpose_reduce32(acc + c, lane);
        atomicAdd(&ssq_s[c + lane], tot);
      }
      // ------------------------------- hand the video over to the rescale warpgroup -------------------------------
      const int p = it & 1;
      if (it >= 2) mbar_wait(&resc_done[p], ((it >> 1) - 1) & 1);     // WG2 finished with fscale_s[p] (video it-2)
      named_bar_sync(1, 128);                                         // ssq_s complete
      if (et == 0) NV_T(26);
      if (et < KC) {
        const float ss = ssq_s[et];
        const float rs = rsqrtf(fmaxf(ss, 1e-12f));
        fscale_s[p * KC + et] = rs;
        atomicAdd(total_s, ss * rs * rs);
        if (stats) {                                           // saved for the backward pass: a_sum, ||V_k||^2
          stats[static_cast<long long>(b) * (2 * KC + 1) + et] = asum_s[et];
          stats[static_cast<long long>(b) * (2 * KC + 1) + KC + et] = ss;
        }
      }
      named_bar_sync(1, 128);
      if (et < KC) fscale_s[p * KC + et] *= rsqrtf(fmaxf(*total_s, 1e-12f));
      if (stats && et == 0) stats[static_cast<long long>(b) * (2 * KC + 1) + 2 * KC] = *total_s;
      if (boss && elect_one()) bulk_wait_group0();                    // every stash store has been performed
      __threadfence();
      named_bar_sync(1, 128);
      if (et == 0) { mbar_arrive(&resc_go[p]); NV_T(27); }
    }
    tc_fence_before();
  } else {
    setmaxnreg_dec<104>();
    // ============================ WG2: final rescale of video it (coalesced, latency-tolerant) ============================
    const int rt = threadIdx.x - 256;                      // 0..127
    for (int it = 0; it < n_iter; ++it) {
      const int b = blockIdx.x + it * gridDim.x;
      const int p = it & 1;
      mbar_wait(&resc_go[p], (it >> 1) & 1);
      const float* fsp = fscale_s + p * KC;
      __nv_bfloat16* ohi = out_hi + static_cast<long long>(b) * ld_out;
      __nv_bfloat16* olo = out_lo ? out_lo + static_cast<long long>(b) * ld_out : nullptr;
      float* of = out_f32 ? out_f32 + static_cast<long long>(b) * ld_out : nullptr;
      const long long n = static_cast<long long>(D) * KC;
      constexpr int kU = 4;
      for (long long e0 = static_cast<long long>(rt) * 8; e0 < n; e0 += 128 * 8 * kU) {
        uint4 h[kU], lw[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long long e = e0 + static_cast<long long>(u) * 128 * 8;
          h[u] = make_uint4(0, 0, 0, 0);
          lw[u] = make_uint4(0, 0, 0, 0);
          if (e < n) {
            h[u] = ld_global_hint(ohi + e, kEvictFirst);
            if (olo) lw[u] = ld_global_hint(olo + e, kEvictFirst);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long long e = e0 + static_cast<long long>(u) * 128 * 8;
          if (e >= n) break;
          const int k0 = static_cast<int>(e % KC);
          float v[8];
          unpack8_stash(h[u], lw[u], out_f16, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= fsp[k0 + j];
          uint4 nh, nl;
          pack8_stash(v, out_f16, nh, nl);
          *reinterpret_cast<uint4*>(ohi + e) = nh;
          if (olo) *reinterpret_cast<uint4*>(olo + e) = nl;
          if (of) {
            *reinterpret_cast<float4*>(of + e) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(of + e + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&resc_done[p]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_netvlad_v3(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const yt8m_bf16* cw2_hi, const yt8m_bf16* cw2_lo, float* out_f32,
                      yt8m_bf16* out_hi, yt8m_bf16* out_lo, long long ld_out, int out_f16, float* stats, cudaStream_t stream) {
  using C = Nv3Cfg;
  CUtensorMap tm_x, tm_x64, tm_cw, tm_c2_hi, tm_c2_lo, tm_out_hi, tm_out_lo;
  int rc;
  {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(D / 64), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, 128, static_cast<uint64_t>(T) * D * 2};
    const uint32_t box[4] = {64, 128, 2, 1};
    const uint32_t box64[4] = {64, 64, 2, 1};
    if ((rc = make_tmap_bf16_nd(&tm_x, x, 4, dims, strides, box)) != YT8M_OK) return rc;
    if ((rc = make_tmap_bf16_nd(&tm_x64, x, 4, dims, strides, box64)) != YT8M_OK) return rc;
  }
  {
    const uint64_t dims[3] = {64, 64, static_cast<uint64_t>(D / 64)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, 128};
    const uint32_t box[3] = {64, 64, 2};
    if ((rc = make_tmap_bf16_nd(&tm_cw, cw_packed, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  if ((rc = make_tmap_bf16_2d(&tm_c2_hi, cw2_hi, D, 64, 64, 128)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_c2_lo, cw2_lo, D, 64, 64, 128)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_out_hi, out_hi, static_cast<uint64_t>(B) * D, 64, 64, 128)) != YT8M_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_out_lo, out_lo ? out_lo : out_hi, static_cast<uint64_t>(B) * D, 64, 64, 128)) != YT8M_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(netvlad_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_done = true;
  }
  const int grid = B < kNvSms ? B : kNvSms;
  netvlad_v3_kernel<<<grid, kNv3Threads, C::kTotal, stream>>>(tm_x, tm_x64, tm_cw, tm_c2_hi, tm_c2_lo, tm_out_hi, tm_out_lo, num_frames, B, T, D,
                                                              scale, shift, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                              reinterpret_cast<__nv_bfloat16*>(out_lo), ld_out, out_f16, stats);
  return check_launch("netvlad_v3_kernel");
}

template <int KC>
int launch_netvlad(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed,
                   const float* scale, const float* shift, const float* cw2, float* out_f32, yt8m_bf16* out_hi,
                   yt8m_bf16* out_lo, long long ld_out, int out_f16, float* stats, cudaStream_t stream) {
  using C = NvCfg<KC>;
  CUtensorMap tm_x, tm_cw;
  int rc;
  // X viewed as [B][D/64][T][64]: one box = 128 frames x two 64-wide k-blocks (32 KB)
  {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(D / 64), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, 128, static_cast<uint64_t>(T) * D * 2};
    const uint32_t box[4] = {64, 128, 2, 1};
    if ((rc = make_tmap_bf16_nd(&tm_x, x, 4, dims, strides, box)) != YT8M_OK) return rc;
  }
  // Cw viewed as [D/64][KC][64]
  {
    const uint64_t dims[3] = {64, static_cast<uint64_t>(KC), static_cast<uint64_t>(D / 64)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, 128};
    const uint32_t box[3] = {64, static_cast<uint32_t>(KC), static_cast<uint32_t>(C::kCwKb)};
    if ((rc = make_tmap_bf16_nd(&tm_cw, cw_packed, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  auto kern = netvlad_fused_kernel<KC>;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_done = true;
  }
  const int grid = B < kNvSms ? B : kNvSms;               // persistent: one CTA per SM
  kern<<<grid, kNvThreads, C::kTotal, stream>>>(tm_x, tm_cw, num_frames, B, T, D, scale, shift, cw2, out_f32,
                                             reinterpret_cast<__nv_bfloat16*>(out_hi),
                                             reinterpret_cast<__nv_bfloat16*>(out_lo), ld_out, out_f16, stats);
  return check_launch("netvlad_fused_kernel");
}

}  // namespace

namespace yt8m {
unsigned long long*& host_debug_timeline() {
  static unsigned long long* p = nullptr;
  return p;
}
}
extern "C" int yt8m_debug_set_timeline(unsigned long long* dev_buf) {
  yt8m::host_debug_timeline() = dev_buf;        // the v4 kernel takes it as an argument
  YT8M_CUDA(cudaMemcpyToSymbol(g_nv_timeline, &dev_buf, sizeof(dev_buf)));
  return YT8M_OK;
}
extern "C" int yt8m_debug_set_flags(int flags) {
  host_debug_flags() = flags;
  YT8M_CUDA(cudaMemcpyToSymbol(g_nv_flags, &flags, sizeof(flags)));
  return YT8M_OK;
}

extern "C" int yt8m_netvlad_fwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K,
                                const yt8m_bf16* cw_packed, const float* scale, const float* shift, const float* cw2,
                                const yt8m_bf16* cw2_hi, const yt8m_bf16* cw2_lo, float* out_f32, yt8m_bf16* out_hi,
                                yt8m_bf16* out_lo, long long ld_out, int out_fmt, float* stats, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && num_frames && cw_packed && cw2 && out_hi, YT8M_E_BADPTR,
               "yt8m_netvlad_fwd: null pointer (out_hi is required: it doubles as the stash)");
  YT8M_REQUIRE(B > 0 && T > 0 && T <= 128 * kNtMax && D > 0 && D % 128 == 0, YT8M_E_BADSHAPE,
               "yt8m_netvlad_fwd: need 0 < T <= %d and D %% 128 == 0 (T=%d D=%d)", 128 * kNtMax, T, D);
  YT8M_REQUIRE(ld_out >= static_cast<long long>(D) * K && ld_out % 8 == 0, YT8M_E_BADSHAPE, "yt8m_netvlad_fwd: ld_out");
  YT8M_REQUIRE(aligned16(out_hi) && (!out_lo || aligned16(out_lo)) && (!out_f32 || aligned16(out_f32)) && aligned16(cw2),
               YT8M_E_BADPTR, "yt8m_netvlad_fwd: outputs / cw2 must be 16-byte aligned");
  YT8M_REQUIRE(out_fmt == YT8M_FMT_BF16 || (out_fmt == YT8M_FMT_F16 && !out_lo), YT8M_E_UNSUPPORTED,
               "yt8m_netvlad_fwd: out_fmt must be YT8M_FMT_BF16, or YT8M_FMT_F16 without a lo tensor");
  const int out_f16 = out_fmt == YT8M_FMT_F16;
  // (debug flag 1 << 20: the four-CTA-cluster kernel of yt8m_netvlad_fwd_tiled with ROW-MAJOR epilogue accesses -- slower than
  //  its tiled form, kept to check the kernel against this entry point's layout)
  if ((host_debug_flags() & (1 << 20)) && !out_lo && !out_f32 && ld_out == static_cast<long long>(D) * K && netvlad_v5_supported(T, D, K))
    return launch_netvlad_v5(x, num_frames, B, T, D, K, cw_packed, scale, shift, cw2, out_hi, out_f16, stats, stream);
  // K = 64, a single 16-bit output tensor, at least three 32-frame tiles: the one-pass two-CTA-cluster kernel (yt8m_netvlad_v4.cu)
  if (K == 64 && !out_lo && !out_f32 && ld_out == static_cast<long long>(D) * K && D / 128 <= 9 && T > 64 &&
      !(host_debug_flags() & 4096))
    return launch_netvlad_v4(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out_hi, out_f16, stats, stream);
  // K = 64 with the bf16 hi/lo copy of cw2 and a dense output: the TMA-staged kernel
  if (K == 64 && cw2_hi && cw2_lo && ld_out == static_cast<long long>(D) * K)
    return launch_netvlad_v3(x, num_frames, B, T, D, cw_packed, scale, shift, cw2_hi, cw2_lo, out_f32, out_hi, out_lo, ld_out, out_f16, stats, stream);
  switch (K) {
    case 32: return launch_netvlad<32>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out_f32, out_hi, out_lo, ld_out, out_f16, stats, stream);
    case 64: return launch_netvlad<64>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out_f32, out_hi, out_lo, ld_out, out_f16, stats, stream);
    case 128: return launch_netvlad<128>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out_f32, out_hi, out_lo, ld_out, out_f16, stats, stream);
    default:
      set_error("yt8m_netvlad_fwd: cluster count K=%d unsupported (32, 64, 128)", K);
      return YT8M_E_UNSUPPORTED;
  }
}
