// yt8m_b200 -- NetVLAD as TWO streaming kernels for sm_100a (K = 64; yt8m_netvlad_fwd_tiled, include/yt8m_b200.h).
//
// (NetVLAD is not part of /root/reference; definition: oracle/yt8m_oracle.py:netvlad_pool.)
//
// Why two kernels.  The one-pass kernels (yt8m_netvlad_v4.cu, yt8m_netvlad_v5.cu) keep a frame tile in shared memory from the
// assignment GEMM, through the cross-CTA exchange of partial logits (the feature axis has to be split: the centres alone are
// 147 KB), the softmax and the broadcast of the assignment, to the aggregation GEMM: a DEPENDENT chain of ~5 us per tile with
// room for three tiles per CTA -- 1.9 us per 64-frame tile whatever the arithmetic (measured: tools/netvlad_v5_scan.py,
// profiles/r02b_netvlad_v5_scan.txt).  Splitting at the assignment removes the chain:
//   K1  netvlad_assign_kernel     a[b, t, :] = softmax_K(scale * (x[b, t, :] . Cw) + shift), masked by num_frames, bf16.
//       One CTA per frame tile over the FULL feature axis: the centres (K x D bf16, 147 KB) stay resident in shared memory, the
//       64-frame tile streams through a ring of 2-k-block stages, logits land frame-major in TMEM (lane = frame), the whole
//       softmax of a frame runs in the registers of one thread, the assignment tile leaves through a TMA store.  No cluster, no
//       exchange.  It writes 128 B per frame (0.06 of what it reads).
//   K2  netvlad_aggregate_kernel  V^T[d, k] += x^T . a per video, residual, intra-norm, final L2 norm, tiled descriptor.
//       A cluster of four CTAs per video splits the feature axis (no reduction needed for this GEMM); x tiles and the matching
//       assignment tiles stream through a four-slot TMA ring; the only cross-CTA traffic is 64 partial norms per video.
// x is read twice; the second read comes out of L2 when the batch fits (the bench batch: 97 MB of real frames, 126 MB of L2).
// Both kernels stream only the ceil(num_frames / 64) tiles that hold real frames and schedule videos longest first.
#include <cstdio>
#include <cstdlib>
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace yt8m {
int launch_netvlad_v6(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2_tiled, yt8m_bf16* out_tiled, int out_f16, float* stats,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
bool netvlad_v6_supported(int T, int D, int K);
size_t netvlad_v6_workspace_bytes(int B, int T, int K);
}

namespace {

constexpr int KC = 64;                       // clusters
constexpr int kF = 64;                       // frames per tile
constexpr int kSubBytes = kF * 128;          // one 64-feature sub-tile of a frame tile: 64 rows x 128 B
constexpr int kMaxIter = 40;                 // videos per CTA / cluster that get the longest-first schedule
constexpr int kMaxTileBuckets = 8;           // tiles per video <= 8 (T <= 512)

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 25)) __trap(); }     // ~100 cycles per probe: traps after seconds
}
__device__ __forceinline__ void wait_bar_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) { if (++spins > (1u << 25)) __trap(); }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t taddr_of(uint32_t tmem_base, int quadrant) {
  return tmem_base + (static_cast<uint32_t>(quadrant * 32) << 16);
}
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// Longest-first serpentine schedule of the B videos over `n_units` CTAs / clusters (unit `uid`): a counting sort over the tile
// count; every CTA derives the same assignment from num_frames on its own.  sched[0] = videos of this unit, sched[1 + w] = video
// of wave w, sched[1 + kMaxIter + w] = its tiles.  Returns false (round-robin instead) when a unit would get more than kMaxIter
// videos.  All threads of the CTA must call it; `sched` needs 1 + 2 * kMaxIter + (2 + nwarps) * kMaxTileBuckets ints.
template <int kThreads>
__device__ bool build_schedule(int* sched, const int* __restrict__ num_frames, int B, int T, int n_units, int uid, int warp, int lane) {
  const int NT = (T + kF - 1) / kF;
  auto tiles_of = [&](int nfv) { return min(max((min(nfv, T) + kF - 1) / kF, 1), NT); };
  if ((B + n_units - 1) / n_units > kMaxIter || NT > kMaxTileBuckets) return false;
  int* bucket_cnt = sched + 1 + 2 * kMaxIter;
  int* bucket_base = bucket_cnt + kMaxTileBuckets;
  int* warp_cnt = bucket_base + kMaxTileBuckets;
  constexpr int kWarps = kThreads / 32;
  if (threadIdx.x == 0) sched[0] = 0;
  const int per_warp = (B + kWarps - 1) / kWarps;
  const int wb0 = min(warp * per_warp, B), wb1 = min(wb0 + per_warp, B);
  {
    int cnt[kMaxTileBuckets];
#pragma unroll
    for (int t = 0; t < kMaxTileBuckets; ++t) cnt[t] = 0;
    for (int bb = wb0; bb < wb1; bb += 32) {
      const int b = bb + lane;
      const int nt = b < wb1 ? tiles_of(__ldg(num_frames + b)) : 0;
#pragma unroll
      for (int t = 1; t <= kMaxTileBuckets; ++t) cnt[t - 1] += __popc(__ballot_sync(0xffffffffu, nt == t));
    }
#pragma unroll
    for (int t = 0; t < kMaxTileBuckets; ++t)
      if (lane == t) warp_cnt[warp * kMaxTileBuckets + t] = cnt[t];
  }
  __syncthreads();
  if (threadIdx.x < kMaxTileBuckets) {
    int tot = 0;
    for (int w = 0; w < kWarps; ++w) {
      const int c = warp_cnt[w * kMaxTileBuckets + threadIdx.x];
      warp_cnt[w * kMaxTileBuckets + threadIdx.x] = tot;
      tot += c;
    }
    bucket_cnt[threadIdx.x] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int base = 0;
    for (int t = kMaxTileBuckets - 1; t >= 0; --t) { bucket_base[t] = base; base += bucket_cnt[t]; }   // longest first
  }
  __syncthreads();
  {
    int run[kMaxTileBuckets];
#pragma unroll
    for (int t = 0; t < kMaxTileBuckets; ++t) run[t] = 0;
    for (int bb = wb0; bb < wb1; bb += 32) {
      const int b = bb + lane;
      const int nt = b < wb1 ? tiles_of(__ldg(num_frames + b)) : 0;
      int r = -1;
#pragma unroll
      for (int t = 1; t <= kMaxTileBuckets; ++t) {
        const unsigned m = __ballot_sync(0xffffffffu, nt == t);
        if (nt == t) r = bucket_base[t - 1] + warp_cnt[warp * kMaxTileBuckets + t - 1] + run[t - 1] + __popc(m & ((1u << lane) - 1u));
        run[t - 1] += __popc(m);
      }
      if (r >= 0) {
        const int w = r / n_units, pos = r - w * n_units;
        if (((w & 1) ? n_units - 1 - pos : pos) == uid) {
          sched[1 + w] = b;
          sched[1 + kMaxIter + w] = nt;
          atomicMax(&sched[0], w + 1);
        }
      }
    }
  }
  __syncthreads();
  return true;
}
template <int kThreads>
constexpr int sched_ints() { return 1 + 2 * kMaxIter + (2 + kThreads / 32) * kMaxTileBuckets + 7; }

// =====================================================================================================================
// K1: assignment
// =====================================================================================================================
constexpr int kAThreads = 192;               // warp 0: TMA producer | 1: MMA issuer | 2-5: softmax epilogue (one per TMEM lane quadrant)
constexpr int kAStageKb = 2;                 // 64-feature blocks per ring stage
constexpr int kAStageBytes = kAStageKb * kSubBytes;      // 16 KB
constexpr int kAStages = 4;
constexpr int kAMaxKb = 18;                  // resident centres: D <= 1152
constexpr int kACwSub = KC * 128;            // one 64-feature block of the centres: 64 rows x 128 B
constexpr int kAOffCw = 0;
constexpr int kAOffRing = kAOffCw + kAMaxKb * kACwSub;                 // 144 KB
constexpr int kAOffStage = kAOffRing + kAStages * kAStageBytes;        // 208 KB
constexpr int kAOffSmall = kAOffStage + kSubBytes;                     // 216 KB: one 64 x 128 B output staging tile
constexpr int kASmallBytes = 2 * KC * 4 + 32 * 8 + sched_ints<kAThreads>() * 4 + 64;
constexpr int kASmemTotal = kAOffSmall + kASmallBytes;
static_assert(kASmemTotal <= 227 * 1024, "assign kernel shared-memory budget exceeded");

__global__ void __launch_bounds__(kAThreads, 1)
netvlad_assign_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cw,
                      const __grid_constant__ CUtensorMap tm_a, const int* __restrict__ num_frames, int B, int T, int D,
                      const float* __restrict__ scale, const float* __restrict__ shift) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* cws = smem + kAOffCw;
  uint8_t* ring = smem + kAOffRing;
  uint8_t* ostage = smem + kAOffStage;
  float* scale_s = reinterpret_cast<float*>(smem + kAOffSmall);
  float* shift_s = scale_s + KC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(shift_s + KC);
  uint64_t* cw_full = bars;                      // [1]
  uint64_t* full = cw_full + 1;                  // [kAStages] TMA -> MMA
  uint64_t* empty = full + kAStages;             // [kAStages] MMA commit -> producer
  uint64_t* s_full = empty + kAStages;           // [2] MMA commit -> epilogue
  uint64_t* s_free = s_full + 2;                 // [2] epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  int* sched = reinterpret_cast<int*>(smem + kAOffSmall + 2 * KC * 4 + 32 * 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_units = static_cast<int>(gridDim.x), uid = static_cast<int>(blockIdx.x);
  const int NT = (T + kF - 1) / kF;
  const int nkb = D / 64;
  const int nst = (nkb + kAStageKb - 1) / kAStageKb;
  auto tiles_of = [&](int nfv) { return min(max((min(nfv, T) + kF - 1) / kF, 1), NT); };
  const bool use_list = build_schedule<kAThreads>(sched, num_frames, B, T, n_units, uid, warp, lane);
  const int n_iter = use_list ? sched[0] : (B - uid + n_units - 1) / n_units;
  auto vid = [&](int it) { return use_list ? sched[1 + it] : uid + it * n_units; };
  auto vnt = [&](int it) { return use_list ? sched[1 + kMaxIter + it] : tiles_of(__ldg(num_frames + uid + it * n_units)); };
  int total_tiles = 0;
  for (int it = 0; it < n_iter; ++it) total_tiles += vnt(it);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_cw); tma_prefetch_desc(&tm_a);
    mbar_init(cw_full, 1);
    for (int i = 0; i < kAStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * KC);
  for (int k = threadIdx.x; k < KC; k += kAThreads) {
    scale_s[k] = scale ? scale[k] : 1.0f;
    shift_s[k] = shift ? shift[k] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (elect_one()) {
      mbar_arrive_expect_tx(cw_full, nkb * kACwSub);
      for (int kb = 0; kb < nkb; ++kb) tma_load_3d(cws + kb * kACwSub, &tm_cw, cw_full, 0, 0, kb, kEvictLast);
    }
    __syncwarp();
    int S = 0;                                       // stage counter across tiles
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it), ntv = vnt(it);
      for (int i = 0; i < ntv; ++i) {
        for (int st = 0; st < nst; ++st, ++S) {
          const int stage = S % kAStages, u = S / kAStages;
          wait_bar(&empty[stage], (u & 1) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], kAStageBytes);      // (blocks past D / 64 are zero-filled and still counted)
            // x stays in L2 for the aggregation kernel that follows (the batch's real frames mostly fit)
            tma_load_4d(ring + stage * kAStageBytes, &tm_x, &full[stage], 0, i * kF, st * kAStageKb, b, kEvictNormal);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer ===================================
    // S[f, k] = X . Cw^T (K-major x K-major), M = 64 frames: frame 16 j + i lands in TMEM lane 32 j + i
    constexpr uint32_t idesc = make_idesc_bf16(64, KC, 0, 0);
    wait_bar(cw_full, 0);
    int S = 0;
    for (int G = 0; G < total_tiles; ++G) {
      const int sb = G & 1, us = G >> 1;
      wait_bar(&s_free[sb], (us & 1) ^ 1u);
      tc_fence_after();
      for (int st = 0; st < nst; ++st, ++S) {
        const int stage = S % kAStages, u = S / kAStages;
        wait_bar(&full[stage], u & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(ring + stage * kAStageBytes);
          const uint32_t c_addr = smem_u32(cws) + st * kAStageKb * kACwSub;
          const uint32_t d_tmem = tmem_base + sb * KC;
          const int kbs = min(kAStageKb, nkb - st * kAStageKb);
          for (int kb = 0; kb < kbs; ++kb) {
            const uint64_t adesc0 = make_sdesc_sw128(a_addr + kb * kSubBytes, 16, 1024);
            const uint64_t bdesc0 = make_sdesc_sw128(c_addr + kb * kACwSub, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, sdesc_advance(adesc0, k * 32), sdesc_advance(bdesc0, k * 32), idesc, (st > 0 || kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (st == nst - 1) umma_commit(&s_full[sb]);
        }
        __syncwarp();
      }
    }
  } else {
    // =================================== softmax epilogue ===================================
    // lane quadrant q holds frames 16 q .. 16 q + 15 of the tile in its lanes 0..15: one thread, one frame, 64 logits
    const int q = warp & 3;
    const bool act = lane < 16;
    const int f = q * 16 + lane;
    const int et = (warp - 2) * 32 + lane;                             // 0..127
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it), ntv = vnt(it);
      const int nf = min(max(__ldg(num_frames + b), 0), T);
      for (int i = 0; i < ntv; ++i, ++G) {
        const int sb = G & 1, us = G >> 1;
        wait_bar(&s_full[sb], us & 1);
        tc_fence_after();
        float r[KC];
        tmem_ld32(taddr_of(tmem_base, q) + sb * KC, reinterpret_cast<uint32_t*>(r));
        tmem_ld32(taddr_of(tmem_base, q) + sb * KC + 32, reinterpret_cast<uint32_t*>(r) + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[sb]);
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < KC / 4; ++c) {
          const float4 sc = *reinterpret_cast<const float4*>(scale_s + 4 * c);
          const float4 sh = *reinterpret_cast<const float4*>(shift_s + 4 * c);
          r[4 * c] = r[4 * c] * sc.x + sh.x; r[4 * c + 1] = r[4 * c + 1] * sc.y + sh.y;
          r[4 * c + 2] = r[4 * c + 2] * sc.z + sh.z; r[4 * c + 3] = r[4 * c + 3] * sc.w + sh.w;
          mx = fmaxf(fmaxf(mx, fmaxf(r[4 * c], r[4 * c + 1])), fmaxf(r[4 * c + 2], r[4 * c + 3]));
        }
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < KC; ++j) {
          r[j] = __expf(r[j] - mx);
          sum += r[j];
        }
        const float inv = 1.0f / sum;
        const bool valid = (i * kF + f) < nf;
        // the previous tile's TMA store has finished reading the staging tile (thread 0 waited before arriving here)
        named_bar_sync(1, 128);
        if (act) {
          // select, not multiply: rows of frames >= num_frames may hold non-finite garbage
#pragma unroll
          for (int c = 0; c < KC / 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              pk[j] = pack_bf16x2(__float2bfloat16_rn(valid ? r[8 * c + 2 * j] * inv : 0.0f),
                                  __float2bfloat16_rn(valid ? r[8 * c + 2 * j + 1] * inv : 0.0f));
            *reinterpret_cast<uint4*>(ostage + sw128_offset(f, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async();
        named_bar_sync(2, 128);
        if (et == 0) {
          tma_store_3d(&tm_a, ostage, 0, i * kF, b);          // rows at or beyond T are clipped by the tensor map
          bulk_commit_group();
          bulk_wait_group_read0();
        }
      }
    }
    if (et == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * KC);
  }
}

// =====================================================================================================================
// K2: aggregation + normalisation
// =====================================================================================================================
constexpr int kC = 4;                        // CTAs per cluster
constexpr int kMaxKb = 5;                    // 64-wide feature blocks per CTA (D <= 1280)
constexpr int kMaxMb = 3;                    // 128-row accumulator blocks per CTA
constexpr int kXBytes = kMaxKb * kSubBytes;  // 40 KB
constexpr int kSlotBytes = kXBytes + kSubBytes;          // x tile + its assignment tile (64 frames x 128 B)
constexpr int kSlots = 4;
constexpr int kGThreads = 384;               // warp 0: TMA producer | 1: MMA issuer | 2: a_sum | 3: - | 4-11: epilogue
constexpr int kVCol = 0;                     // TMEM: V^T, kMaxMb x 64 columns
constexpr int kC2Col = kMaxMb * KC;          // TMEM: this CTA's cw2 slice, kMaxMb x 64 columns, resident
constexpr int kGOffSmall = kSlots * kSlotBytes;                         // 192 KB
// floats: asum[2] | ssq_part[2][4] | fscale | contrib | ssq_w[8][32]
constexpr int kGSmallFloats = KC * (2 + 2 * kC + 1 + 1 + 4);
constexpr int kGSmallBytes = kGSmallFloats * 4 + 32 * 8 + sched_ints<kGThreads>() * 4 + 64;
constexpr int kGSmemTotal = kGOffSmall + kGSmallBytes;
static_assert(kGSmemTotal <= 227 * 1024, "aggregate kernel shared-memory budget exceeded");

__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kGThreads, 1)
netvlad_aggregate_kernel(const __grid_constant__ CUtensorMap tm_xa, const __grid_constant__ CUtensorMap tm_xb,
                         const __grid_constant__ CUtensorMap tm_a, uint16_t* __restrict__ out, const int* __restrict__ num_frames,
                         int B, int T, int D, const float* __restrict__ cw2, int out_f16, float* __restrict__ stats, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* asum_s = reinterpret_cast<float*>(smem + kGOffSmall);       // [2][KC] by video parity
  float* ssq_part = asum_s + 2 * KC;             // [2][kC][KC] by video parity and source CTA (own slot written locally)
  float* fscale_s = ssq_part + 2 * kC * KC;      // [KC]
  float* contrib_s = fscale_s + KC;              // [KC]
  float* ssq_w = contrib_s + KC;                 // [8][32] per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssq_w + 4 * KC);
  uint64_t* x_full = bars;                       // [kSlots] TMA (x tile + assignment tile) -> MMA, a_sum
  uint64_t* x_empty = x_full + kSlots;           // [kSlots] MMA commit -> producer
  uint64_t* sum_done = x_empty + kSlots;         // [kSlots] a_sum warp -> MMA: the slot's assignment tile has been read
  uint64_t* asum_ready = sum_done + kSlots;      // [2] a_sum warp -> epilogue, by video parity
  uint64_t* asum_free = asum_ready + 2;          // [2] epilogue -> a_sum warp
  uint64_t* v_full = asum_free + 2;              // [1] MMA commit -> epilogue, per video
  uint64_t* v_free = v_full + 1;                 // [1] epilogue (8 warps) -> MMA
  uint64_t* ssq_full = v_free + 1;               // [2] the three peers' partial sums of squares have landed (st.async bytes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ssq_full + 2);
  int* sched = reinterpret_cast<int*>(smem + kGOffSmall + kGSmallFloats * 4 + 32 * 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_units = static_cast<int>(gridDim.x) / kC, uid = static_cast<int>(blockIdx.x) / kC;
  const int NT = (T + kF - 1) / kF;
  const int nkb_total = D / 64;
  const int kb_base = nkb_total / kC, kb_extra = nkb_total % kC;
  const int nkb = kb_base + (static_cast<int>(rank) < kb_extra ? 1 : 0);
  const int kb0 = static_cast<int>(rank) * kb_base + min(static_cast<int>(rank), kb_extra);
  const int DH = nkb * 64, d0 = kb0 * 64;
  const int nmb = (nkb + 1) >> 1;               // 128-row accumulator blocks (the last one may be half valid)
  auto tiles_of = [&](int nfv) { return min(max((min(nfv, T) + kF - 1) / kF, 1), NT); };
  const bool use_list = build_schedule<kGThreads>(sched, num_frames, B, T, n_units, uid, warp, lane);
  const int n_iter = use_list ? sched[0] : (B - uid + n_units - 1) / n_units;
  auto vid = [&](int it) { return use_list ? sched[1 + it] : uid + it * n_units; };
  auto vnt = [&](int it) { return use_list ? sched[1 + kMaxIter + it] : tiles_of(__ldg(num_frames + uid + it * n_units)); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xa); tma_prefetch_desc(&tm_xb); tma_prefetch_desc(&tm_a);
    for (int i = 0; i < kSlots; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&sum_done[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&asum_ready[i], 1); mbar_init(&asum_free[i], 1); mbar_init(&ssq_full[i], 1); }
    mbar_init(v_full, 1); mbar_init(v_free, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peers' barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // =================================== TMA producer ===================================
    const CUtensorMap* tmx = (static_cast<int>(rank) < kb_extra) ? &tm_xa : &tm_xb;     // box = this CTA's nkb feature blocks
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it), ntv = vnt(it);
      for (int i = 0; i < ntv; ++i, ++G) {
        const int slot = G % kSlots, u = G / kSlots;
        wait_bar(&x_empty[slot], (u & 1) ^ 1u);
        if (elect_one()) {
          uint8_t* s = smem + slot * kSlotBytes;
          mbar_arrive_expect_tx(&x_full[slot], nkb * kSubBytes + kSubBytes);
          tma_load_4d(s, tmx, &x_full[slot], 0, i * kF, kb0, b, kEvictFirst);
          tma_load_3d(s + kXBytes, &tm_a, &x_full[slot], 0, i * kF, b, kEvictFirst);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer ===================================
    constexpr uint32_t idesc1 = make_idesc_bf16(128, KC, 1, 1);     // V^T += X^T . a    (MN-major x MN-major)
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int ntv = vnt(it);
      for (int i = 0; i < ntv; ++i, ++G) {
        const int slot = G % kSlots, u = G / kSlots;
        wait_bar(&x_full[slot], u & 1);
        tc_fence_after();
        if (i == 0) {                                                // the epilogue has taken the previous video out of TMEM
          wait_bar(v_free, (it & 1) ^ 1u);
          tc_fence_after();
        }
        if (elect_one()) {
          const uint32_t x_addr = smem_u32(smem + slot * kSlotBytes);
          const uint64_t bdesc0 = make_sdesc_sw128(x_addr + kXBytes, kSubBytes, 1024);
          for (int m = 0; m < nmb; ++m) {
            // rows m*128 .. +127 of this CTA's features = sub-tiles 2m and 2m+1, one box apart (LBO); the second sub-tile of a
            // half-valid last block is whatever follows in shared memory (its 64 accumulator rows are never read)
            const uint64_t adesc0 = make_sdesc_sw128(x_addr + m * 2 * kSubBytes, kSubBytes, 1024);
            const uint32_t d_tmem = tmem_base + kVCol + m * KC;
#pragma unroll
            for (int s = 0; s < kF / 16; ++s)
              umma_bf16(d_tmem, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1, (i > 0 || s > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
        wait_bar(&sum_done[slot], u & 1);                            // the a_sum warp has read the slot's assignment tile too
        if (elect_one()) {
          umma_commit(&x_empty[slot]);
          if (i == ntv - 1) umma_commit(v_full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // =================================== a_sum: column sums of the assignment tiles =============================
    // lane L owns clusters 2L, 2L+1: one 4-byte word per row, conflict free
    float acc0 = 0.0f, acc1 = 0.0f;
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int ntv = vnt(it);
      const int p = it & 1;
      for (int i = 0; i < ntv; ++i, ++G) {
        const int slot = G % kSlots, u = G / kSlots;
        wait_bar(&x_full[slot], u & 1);
        const uint8_t* at = smem + slot * kSlotBytes + kXBytes;
#pragma unroll 4
        for (int f = 0; f < kF; ++f) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(at + f * 128 + ((((lane >> 2) ^ (f & 7))) << 4) + (lane & 3) * 4);
          acc0 += __uint_as_float(w << 16);
          acc1 += __uint_as_float(w & 0xFFFF0000u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sum_done[slot]);
      }
      wait_bar(&asum_free[p], ((it >> 1) & 1) ^ 1u);                 // the epilogue has consumed a_sum of video it-2
      asum_s[p * KC + 2 * lane] = acc0;
      asum_s[p * KC + 2 * lane + 1] = acc1;
      acc0 = 0.0f; acc1 = 0.0f;
      __syncwarp();
      if (lane == 0) mbar_arrive(&asum_ready[p]);
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ============================ epilogue: residual, norms, output (per video) ============================
    // warp e = (lane quadrant q, cluster half h): rows q*32 .. +31 of every 128-row accumulator block, clusters 32 h .. + 31.
    // The accumulator is read from TMEM once (and handed back to the MMA warp right away); the corrected values stay in
    // registers across the exchange of the per-cluster norms.
    const int e = warp - 4;
    const int q = e & 3, h = e >> 2;
    const int et = e * 32 + lane;                             // 0..255
    // accumulator blocks in which this warp's 32 rows exist (a prefix: only the last block can be half valid)
    const int nmb_w = (DH - q * 32 + 127) / 128 > 0 ? (DH - q * 32 + 127) / 128 : 0;
    const uint32_t tv = taddr_of(tmem_base, q) + kVCol + h * 32;         // this warp's V columns: + m * KC per block
    const uint32_t tc2 = taddr_of(tmem_base, q) + kC2Col + h * 32;       // and the matching block of cw2
    // ---- once: this warp's share of cw2 into TMEM (tiled layout [D/32][K/4 chunks][32 rows][4 floats]: coalesced) ----
#pragma unroll 1
    for (int m = 0; m < nmb_w; ++m) {
      const long long g32 = (d0 + m * 128 + q * 32) >> 5;
      const float4* c2 = reinterpret_cast<const float4*>(cw2) + (g32 * (KC / 4) + h * 8) * 32 + lane;
      float c[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t4 = __ldg(c2 + j * 32);
        c[4 * j] = t4.x; c[4 * j + 1] = t4.y; c[4 * j + 2] = t4.z; c[4 * j + 3] = t4.w;
      }
      tmem_st32(tc2 + m * KC, reinterpret_cast<const uint32_t*>(c));
    }
    tmem_st_wait();
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int p = it & 1;
      const bool dbg_noexch = dbg & ((1 << 27) | (1 << 28));
      if (et == 0 && !dbg_noexch) mbar_arrive_expect_tx(&ssq_full[p], (kC - 1) * KC * 4);   // this video's partial sums from the three peers
      wait_bar(v_full, it & 1);
      tc_fence_after();
      float v[kMaxMb][32];
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m)
        if (m < nmb_w) tmem_ld32(tv + m * KC, reinterpret_cast<uint32_t*>(v[m]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(v_free);                     // the next video's aggregation may overwrite the accumulator
      wait_bar(&asum_ready[p], (it >> 1) & 1);
      if (dbg & (1 << 28)) {                                  // debug: nothing but the hand-offs
        named_bar_sync(1, 256);
        if (et == 0) mbar_arrive(&asum_free[p]);
        continue;
      }
      const float* asum = asum_s + p * KC + h * 32;
      float ssq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ssq[j] = 0.0f;
      // ---- pass 1 (registers): V -= a_sum * cw2 in fp32 (cw2 from TMEM), per-cluster sum of squares ----
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m) {
        if (m < nmb_w && !(dbg & (1 << 26))) {
          float c[32];
          tmem_ld32(tc2 + m * KC, reinterpret_cast<uint32_t*>(c));
          tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 as4 = *reinterpret_cast<const float4*>(asum + j4 * 4);               // broadcast read
            const int j = j4 * 4;
            v[m][j] -= as4.x * c[j]; v[m][j + 1] -= as4.y * c[j + 1]; v[m][j + 2] -= as4.z * c[j + 2]; v[m][j + 3] -= as4.w * c[j + 3];
            ssq[j] += v[m][j] * v[m][j]; ssq[j + 1] += v[m][j + 1] * v[m][j + 1];
            ssq[j + 2] += v[m][j + 2] * v[m][j + 2]; ssq[j + 3] += v[m][j + 3] * v[m][j + 3];
          }
        }
      }
      ssq_w[e * 32 + lane] = warp_transpose_reduce32(ssq, lane);      // lane L: sum over this warp's rows of cluster 32 h + L
      named_bar_sync(1, 256);                                 // this CTA's partial sums are complete
      // ---- all-to-all of the KC partial sums between the four CTAs ----
      if (et < KC) {
        const int hh = et >> 5, kk = et & 31;
        const float mine = (ssq_w[(hh * 4 + 0) * 32 + kk] + ssq_w[(hh * 4 + 1) * 32 + kk]) + (ssq_w[(hh * 4 + 2) * 32 + kk] + ssq_w[(hh * 4 + 3) * 32 + kk]);
        ssq_part[(p * kC + rank) * KC + et] = mine;
        if (!dbg_noexch) {
#pragma unroll
          for (int s = 1; s < kC; ++s) {
            const uint32_t dst = (rank + s) & 3u;
            st_async_f32(mapa_u32(smem_u32(&ssq_part[(p * kC + rank) * KC + et]), dst), mapa_u32(smem_u32(&ssq_full[p]), dst), mine);
          }
        }
      }
      if (!dbg_noexch) wait_bar_cluster(&ssq_full[p], (it >> 1) & 1);
      named_bar_sync(1, 256);                                 // (the own slot was written by threads of other warps)
      if (et < KC) {
        const float* sp = ssq_part + p * kC * KC + et;
        const float ss = (sp[0] + sp[KC]) + (sp[2 * KC] + sp[3 * KC]);        // same order on every CTA
        const float rs = rsqrtf(fmaxf(ss, 1e-12f));
        fscale_s[et] = rs;
        contrib_s[et] = ss * rs * rs;
        if (stats && rank == 0) {                             // saved for the backward pass: a_sum, ||V_k||^2
          stats[static_cast<long long>(b) * (2 * KC + 1) + et] = asum_s[p * KC + et];
          stats[static_cast<long long>(b) * (2 * KC + 1) + KC + et] = ss;
        }
      }
      named_bar_sync(1, 256);
      // every epilogue thread is done with asum_s[p]: the a_sum warp may publish video it+2 into it
      if (et == 0) mbar_arrive(&asum_free[p]);
      const float total = warp_sum(contrib_s[lane] + contrib_s[lane + 32]);     // same tree on every warp of all four CTAs
      const float gs = rsqrtf(fmaxf(total, 1e-12f));
      if (stats && rank == 0 && et == 0) stats[static_cast<long long>(b) * (2 * KC + 1) + 2 * KC] = total;
      // ---- pass 2 (registers): rescale (intra-norm x final L2 norm), convert, store: tiled [D/32][K/8 chunks][32 rows][8 values],
      //      every store instruction of the warp writes 512 contiguous bytes ----
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m) {
        if (m < nmb_w && !(dbg & (1 << 25))) {
          const long long g32 = (d0 + m * 128 + q * 32) >> 5;
          uint16_t* dst = out + static_cast<long long>(b) * D * KC + ((g32 * (KC / 8) + h * 4) * 32 + lane) * 8;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float w8[8];
            const float4 f0 = *reinterpret_cast<const float4*>(fscale_s + h * 32 + j8 * 8), f1 = *reinterpret_cast<const float4*>(fscale_s + h * 32 + j8 * 8 + 4);
            const float fs8[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};          // broadcast reads
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = v[m][j8 * 8 + j] * (fs8[j] * gs);
            uint4 hi, lo;
            if (out_f16) hi = pack8_f16(w8);
            else pack8_hi_lo(w8, hi, lo);
            if (!(dbg & (1 << 24)) || (hi.x == 0x12345678u)) *reinterpret_cast<uint4*>(dst + j8 * 256) = hi;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody exits while a peer may still touch its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool yt8m::netvlad_v6_supported(int T, int D, int K) {
  return K == 64 && D % 64 == 0 && D / 64 >= kC && D / 64 <= kAMaxKb && (D / 64 + kC - 1) / kC <= kMaxKb && T >= 1;
}
size_t yt8m::netvlad_v6_workspace_bytes(int B, int T, int K) { return static_cast<size_t>(B) * T * K * 2 + 256; }

int yt8m::launch_netvlad_v6(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                            const float* scale, const float* shift, const float* cw2_tiled, yt8m_bf16* out_tiled, int out_f16,
                            float* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  YT8M_REQUIRE(netvlad_v6_supported(T, D, K), YT8M_E_BADSHAPE, "netvlad v6: T=%d D=%d K=%d", T, D, K);
  YT8M_REQUIRE(workspace && workspace_bytes >= netvlad_v6_workspace_bytes(B, T, K), YT8M_E_BADSHAPE, "netvlad v6: workspace too small");
  yt8m_bf16* a = reinterpret_cast<yt8m_bf16*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const int nkb_total = D / 64;
  CUtensorMap tm_x1, tm_cw, tm_a, tm_xa, tm_xb;
  int rc;
  {
    // X viewed as [B][D/64][T][64]: K1's box = 64 frames x 2 feature blocks
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(nkb_total), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, 128, static_cast<uint64_t>(T) * D * 2};
    const uint32_t box1[4] = {64, kF, kAStageKb, 1};
    if ((rc = make_tmap_bf16_nd(&tm_x1, x, 4, dims, strides, box1)) != YT8M_OK) return rc;
    // K2's boxes: 64 frames x the CTA's feature blocks (ceil and floor of D / 256)
    const uint32_t boxa[4] = {64, kF, static_cast<uint32_t>((nkb_total + kC - 1) / kC), 1};
    const uint32_t boxb[4] = {64, kF, static_cast<uint32_t>(nkb_total / kC), 1};
    if ((rc = make_tmap_bf16_nd(&tm_xa, x, 4, dims, strides, boxa)) != YT8M_OK) return rc;
    if ((rc = make_tmap_bf16_nd(&tm_xb, x, 4, dims, strides, boxb)) != YT8M_OK) return rc;
  }
  {
    // Cw viewed as [D/64][64 clusters][64]
    const uint64_t dims[3] = {64, KC, static_cast<uint64_t>(nkb_total)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, 128};
    const uint32_t box[3] = {64, KC, 1};
    if ((rc = make_tmap_bf16_nd(&tm_cw, cw_packed, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  {
    // the assignment [B][T][64 clusters] bf16: one box = 64 frames (written by K1's TMA store, read by K2's TMA load)
    const uint64_t dims[3] = {KC, static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {KC * 2, static_cast<uint64_t>(T) * KC * 2};
    const uint32_t box[3] = {KC, kF, 1};
    if ((rc = make_tmap_bf16_nd(&tm_a, a, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  static int max_clusters = -1, sms = 0;
  if (max_clusters < 0) {
    YT8M_CUDA(cudaFuncSetAttribute(netvlad_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kASmemTotal));
    YT8M_CUDA(cudaFuncSetAttribute(netvlad_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    int dev = 0;
    YT8M_CUDA(cudaGetDevice(&dev));
    YT8M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(kC * 37, 1, 1);
    qc.blockDim = dim3(kGThreads, 1, 1);
    qc.dynamicSmemBytes = kGSmemTotal;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, netvlad_aggregate_kernel, &qc) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    max_clusters = n > 0 ? n : 32;              // (the query can fail under a profiler: 32 clusters always fit 148 SMs)
    if (getenv("YT8M_VERBOSE")) fprintf(stderr, "netvlad v6: occupancy query -> %d clusters of %d CTAs (%d SMs)\n", n, kC, sms);
    if (max_clusters > sms / kC) max_clusters = sms / kC;
  }
  const int ctas1 = B < sms ? B : sms;
  if (!(host_debug_flags() & (1 << 23))) {       // (debug: 1 << 23 skips the assignment kernel, 1 << 22 the aggregation kernel)
    netvlad_assign_kernel<<<ctas1, kAThreads, kASmemTotal, stream>>>(tm_x1, tm_cw, tm_a, num_frames, B, T, D, scale, shift);
    if ((rc = check_launch("netvlad_assign_kernel")) != YT8M_OK) return rc;
  }
  if (host_debug_flags() & (1 << 22)) return YT8M_OK;
  const int clusters = B < max_clusters ? B : max_clusters;
  netvlad_aggregate_kernel<<<kC * clusters, kGThreads, kGSmemTotal, stream>>>(tm_xa, tm_xb, tm_a, reinterpret_cast<uint16_t*>(out_tiled),
                                                                             num_frames, B, T, D, cw2_tiled, out_f16, stats, host_debug_flags());
  return check_launch("netvlad_aggregate_kernel");
}
