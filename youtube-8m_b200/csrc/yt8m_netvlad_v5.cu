// yt8m_b200 -- fused NetVLAD v5 for sm_100a: ONE pass over the frames, a CLUSTER OF FOUR CTAs per video.
//
// (NetVLAD is not part of /root/reference; definition: oracle/yt8m_oracle.py:netvlad_pool.  Same arithmetic as
//  yt8m_netvlad_v4.cu.)
//
// What bounded v4 (two CTAs per video, 32-frame tiles; profiles/r01e_netvlad_v4_timeline.txt, r02a_bench_c2.json: 77 us,
// 0.26 of HBM): (1) 46 small UMMAs per 32-frame tile -- the tensor pipe is ISSUE bound at ~21 ns per instruction whatever
// the shape; (2) V^T took 320 of the 512 TMEM columns, so the per-video epilogue (7.5 us) could not overlap the next video's
// aggregation; (3) a three-slot ring of 36 KB tiles beside 72 KB of resident centres left nothing to deepen the pipeline.
// v5 splits the feature axis over FOUR CTAs (CTA r owns k-blocks [kb0_r, kb0_r + nkb_r) of 64 features: 5,5,4,4 for D = 1152):
//   * resident centres are 40 KB, a 64-frame tile is 40 KB: three tiles in flight and UMMAs twice as large
//     (phase 0: M = 64 frames x N = K clusters; phase 1: M = 128 features x N = K, 64 frames per tile) -- 32 instructions per
//     64 frames instead of 92;
//   * V^T is 3 x K columns: TWO accumulators (K = 64), so pass 1 / pass 2 of video i run beside the tiles of video i+1;
//   * the logits are laid out FRAME-major in TMEM (lane = frame): the partial logits over a CTA's features go to the
//     frame's OWNER CTA (frames 16q..16q+15 of a tile belong to CTA q: a reduce-scatter through st.async DSMEM stores that
//     complete a transaction count on the owner's mbarrier), the owner thread adds the four partials, does the whole
//     masked softmax in registers and broadcasts the bf16 assignment row straight into the MN-major operand tile of all four
//     CTAs (st.async again).  18 KB of DSMEM traffic per CTA and 64-frame tile.
//   * a_sum is accumulated by a spare warp from the assignment tile every CTA holds anyway (no cross-CTA exchange).
//
// Warp roles (512 threads, four warpgroups with their own register budgets):
//   0 TMA producer | 1 MMA issuer phase 0 | 2 MMA issuer phase 1 | 3 a_sum | 4-7 exchange + softmax (one per TMEM lane
//   quadrant; each warp OWNS 4 frames of every tile, 8 threads per frame) | 8-15 epilogue (two per lane quadrant, one per
//   32-cluster half: the video's accumulator is read from TMEM ONCE and stays in registers across the norm exchange).
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace yt8m {
int launch_netvlad_v5(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                      cudaStream_t stream);
bool netvlad_v5_supported(int T, int D, int K);
}

namespace {

constexpr int kC = 4;                        // CTAs per cluster
constexpr int kF = 64;                       // frames per tile
constexpr int kOwn = kF / kC;                // frames of a tile owned by one CTA: 4 from each TMEM lane quadrant (16 active lanes)
constexpr int kMaxKb = 5;                    // 64-wide feature blocks per CTA (D <= 1280)
constexpr int kMaxMb = 3;                    // 128-row accumulator blocks per CTA
constexpr int kSubBytes = kF * 128;          // one 64-feature sub-tile of a frame tile: 64 rows x 128 B
constexpr int kXSlotBytes = kMaxKb * kSubBytes;     // 40 KB
constexpr int kThreads = 512;
constexpr int kMaxIter = 40;                 // videos per cluster that get the longest-first schedule
constexpr int kMaxTileBuckets = 8;           // tiles per video <= 8 (T <= 512)

template <int KC>
struct Cfg {
  static constexpr int kSlots = KC == 64 ? 3 : 2;        // X tiles in flight
  static constexpr int kSB = KC == 64 ? 2 : 1;           // S (logits) accumulators in TMEM
  static constexpr int kVB = 1;                          // V accumulators in TMEM (the other 3 x KC columns hold this CTA's cw2 slice)
  static constexpr int kAT = KC == 64 ? 2 : 1;           // assignment tiles in shared memory
  static constexpr int kRB = KC == 64 ? 2 : 1;           // receive buffers for partial logits
  static constexpr int kCwSubBytes = KC * 128;           // one 64-feature block of the centres: KC rows x 128 B
  static constexpr int kATileBytes = kF * 2 * KC;        // 64 frames x KC clusters bf16: KC / 64 MN-major atoms of 8 KB
  static constexpr int kRecvRow = KC * 4;                // partial logits of one frame, fp32
  static constexpr int kRecvWarp = kC * 4 * kRecvRow;    // one exchange warp's buffer: [4 sources (3 peers + own)][4 frames] rows
  static constexpr int kRecvBytes = 4 * kRecvWarp;
  static constexpr int kSCol = 0;
  static constexpr int kVCol = kSB * KC;                 // 128
  static constexpr int kVStride = kMaxMb * KC;           // columns of one V accumulator
  static constexpr int kC2Col = kVCol + kVB * kVStride;  // fp32 cw2[d0 .. d0 + DH, :] resident for the whole kernel (KC = 64)
  static_assert(kC2Col + (KC == 64 ? kVStride : 0) <= 512, "TMEM budget");
  static constexpr int kOffCw = 0;
  static constexpr int kOffX = kOffCw + kMaxKb * kCwSubBytes;
  static constexpr int kOffA = kOffX + kSlots * kXSlotBytes;
  static constexpr int kOffRecv = kOffA + kAT * kATileBytes;
  static constexpr int kOffSmall = kOffRecv + kRB * kRecvBytes;
  // floats: scale, shift | asum[2] | ssq | ssq_part[2][4] | fscale | contrib | ssq_w[8][32]
  static constexpr int kSmallFloats = KC * (2 + 2 + 1 + 2 * kC + 1 + 1 + 4);
  static constexpr int kBarBytes = 80 * 8;
  static constexpr int kSchedInts = 1 + 2 * kMaxIter + 2 * kMaxTileBuckets + (kThreads / 32) * kMaxTileBuckets + 7;
  static constexpr int kSmallBytes = kSmallFloats * 4 + kBarBytes + kSchedInts * 4 + 64;
  static constexpr int kSmemTotal = kOffSmall + kSmallBytes;
  static_assert(kSmemTotal <= 227 * 1024, "NetVLAD v5 shared-memory budget exceeded");
};

// debug-only phase timeline (globaltimer ns) of cluster 0 / CTA rank 0, first 3 videos, 128 stamps per video:
// [0,8) mma0: tile landed   [8,16) exch: S ready   [16,24) owner: partials landed   [24,32) owner: assignment sent
// [32,40) mma1: assignment landed   [40,48) mma1: issued   48 epi: video complete   49 pass 1 done   50 norms exchanged
// 51 pass 2 done   [56,64) producer: slot free
#ifdef YT8M_V5_TIMELINE
#define NV5_T(itv, slot)                                                                                          \
  do {                                                                                                            \
    if (timeline && blockIdx.x == 0 && (itv) < 3) timeline[(itv) * 128 + (slot)] = global_timer_ns();              \
  } while (0)
#else
#define NV5_T(itv, slot) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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

// Code size matters here: sixteen warps in six roles run concurrently out of one instruction cache (the first version's
// straight-line bodies -- ~3100 hot SASS instructions -- made pure-ALU stretches run at ~15 cycles per instruction).  The
// bounded-wait loops are therefore a few instructions each, loops stay rolled where registers allow, and the debug timeline
// is compiled in only with -DYT8M_V5_TIMELINE.
// compact bounded waits (ptxas cannot allocate registers for real calls in a kernel that uses setmaxnreg, so they are inline;
// ~100 cycles per probe: a wedged pipeline traps after a few seconds instead of hanging the GPU)
__device__ __forceinline__ void wait_bar_inl(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 25)) __trap(); }
}
__device__ __forceinline__ void wait_bar_cluster_inl(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) { if (++spins > (1u << 25)) __trap(); }
}

__device__ __forceinline__ uint32_t taddr_of(uint32_t tmem_base, int quadrant) {
  return tmem_base + (static_cast<uint32_t>(quadrant * 32) << 16);
}
__device__ __forceinline__ void st_async_v4_b32(uint32_t cluster_addr, uint32_t cluster_bar, uint32_t a, uint32_t b, uint32_t c,
                                                uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in every CTA of `mask` when all previously issued tcgen05.mma of
// this thread have completed
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// TILED = the blocked descriptor / centre layouts of yt8m_netvlad_fwd_tiled (include/yt8m_b200.h): every global access of the
// epilogue is then a 512-byte contiguous warp access straight from / to the registers that tcgen05.ld filled (lane = feature
// row), with no shared-memory transposition.
template <int KC, bool TILED>
__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1)
netvlad_v5_kernel(const __grid_constant__ CUtensorMap tm_xa, const __grid_constant__ CUtensorMap tm_xb,
                  const __grid_constant__ CUtensorMap tm_cw, uint16_t* __restrict__ out, const int* __restrict__ num_frames, int B,
                  int T, int D, const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ cw2,
                  int out_f16, float* __restrict__ stats, unsigned long long* __restrict__ timeline, int dbg_flags) {
  using C = Cfg<KC>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* cws = smem + C::kOffCw;
  uint8_t* xs = smem + C::kOffX;
  uint8_t* atile = smem + C::kOffA;
  uint8_t* recvbuf = smem + C::kOffRecv;
  float* scale_s = reinterpret_cast<float*>(smem + C::kOffSmall);
  float* shift_s = scale_s + KC;
  float* asum_s = shift_s + KC;                  // [2][KC] by video parity
  float* ssq_s = asum_s + 2 * KC;                // [KC] this CTA's partial sums of squares
  float* ssq_part = ssq_s + KC;                  // [2][kC][KC] by video parity and source CTA (own slot written locally)
  float* fscale_s = ssq_part + 2 * kC * KC;      // [KC]
  float* contrib_s = fscale_s + KC;              // [KC]
  float* ssq_w = contrib_s + KC;                 // [4][KC] per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssq_w + 4 * KC);
  uint64_t* cw_full = bars;                      // [1]
  uint64_t* x_full = cw_full + 1;                // [3]  TMA -> MMA
  uint64_t* x_empty = x_full + 3;                // [3]  phase-1 commit -> producer
  uint64_t* s_full = x_empty + 3;                // [2]  phase-0 commit -> exchange warps
  uint64_t* s_free = s_full + 2;                 // [2]  exchange warps (4) -> phase 0
  uint64_t* recv_full = s_free + 2;              // [4][2] exchange warp q: the three peers' partial logits of its 4 frames have landed (st.async bytes)
  uint64_t* send_credit = recv_full + 8;         // [4][2] exchange warp q: the three owners have consumed what it pushed into their buffers rb (remote arrives, 3)
  uint64_t* a_full = send_credit + 8;            // [2]  the sixteen owners' assignment rows have landed in my tile (st.async bytes)
  uint64_t* a_credit = a_full + 2;               // [2]  all four CTAs' phase 1 are done with assignment tile ab (multicast commits, 4)
  uint64_t* a_sumdone = a_credit + 2;            // [2]  a_sum warp -> phase 1: tile ab has been read
  uint64_t* asum_ready = a_sumdone + 2;          // [2]  a_sum warp -> epilogue, by video parity
  uint64_t* asum_free = asum_ready + 2;          // [2]  epilogue -> a_sum warp
  uint64_t* v_full = asum_free + 2;              // [2]  phase-1 commit -> epilogue, per accumulator
  uint64_t* v_free = v_full + 2;                 // [2]  epilogue (4 warps) -> phase 1
  uint64_t* ssq_full = v_free + 2;               // [2]  the three peers' partial sums of squares have landed (st.async bytes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ssq_full + 2);
  int* sched = reinterpret_cast<int*>(smem + C::kOffSmall + C::kSmallFloats * 4 + C::kBarBytes);
  // sched[0] = videos of this cluster, [1 + w] = video of wave w, [1 + kMaxIter + w] = its tiles, then scratch

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_clusters = static_cast<int>(gridDim.x) / kC;
  const int cid = static_cast<int>(blockIdx.x) / kC;
  const int NT = (T + kF - 1) / kF;
  // feature blocks of this CTA: 64-wide blocks [kb0, kb0 + nkb)
  const int nkb_total = D / 64;
  const int kb_base = nkb_total / kC, kb_extra = nkb_total % kC;
  const int nkb = kb_base + (static_cast<int>(rank) < kb_extra ? 1 : 0);
  const int kb0 = static_cast<int>(rank) * kb_base + min(static_cast<int>(rank), kb_extra);
  const int DH = nkb * 64;                      // features of this CTA
  const int d0 = kb0 * 64;                      // its first feature
  const int nmb = (nkb + 1) >> 1;               // 128-row accumulator blocks (the last one may be half valid)

  // PADDED FRAMES ARE NOT READ: a video is streamed for ceil(num_frames / 64) tiles only (at least one, which also
  // zero-initialises the accumulators of an empty video).  Videos are handed to the clusters longest first in serpentine
  // order (wave w: cluster c takes rank w*C + c, or w*C + C-1-c when w is odd); every CTA derives the same schedule from
  // num_frames on its own with a counting sort over the tile count.
  auto tiles_of = [&](int nfv) { return min(max((min(nfv, T) + kF - 1) / kF, 1), NT); };
  const bool use_list = (B + n_clusters - 1) / n_clusters <= kMaxIter && NT <= kMaxTileBuckets;
  if (use_list) {
    int* bucket_cnt = sched + 1 + 2 * kMaxIter;            // [kMaxTileBuckets] videos per tile count
    int* bucket_base = bucket_cnt + kMaxTileBuckets;       // [kMaxTileBuckets] first rank of the bucket (descending tile count)
    int* warp_cnt = bucket_base + kMaxTileBuckets;         // [warps][kMaxTileBuckets] -> exclusive prefix over the warps
    constexpr int kWarps = kThreads / 32;
    if (threadIdx.x == 0) sched[0] = 0;
    // the rank inside a bucket follows the video index: each warp takes a contiguous range of videos, ballots give the
    // in-warp prefix, a 12-step scan gives the prefix over the warps
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
          const int w = r / n_clusters, pos = r - w * n_clusters;
          if (((w & 1) ? n_clusters - 1 - pos : pos) == cid) {
            sched[1 + w] = b;
            sched[1 + kMaxIter + w] = nt;
            atomicMax(&sched[0], w + 1);
          }
        }
      }
    }
    __syncthreads();
  }
  const int n_iter = use_list ? sched[0] : (B - cid + n_clusters - 1) / n_clusters;
  auto vid = [&](int it) { return use_list ? sched[1 + it] : cid + it * n_clusters; };
  auto vnt = [&](int it) { return use_list ? sched[1 + kMaxIter + it] : tiles_of(__ldg(num_frames + cid + it * n_clusters)); };
  int total_tiles = 0;
  for (int it = 0; it < n_iter; ++it) total_tiles += vnt(it);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xa); tma_prefetch_desc(&tm_xb); tma_prefetch_desc(&tm_cw);
    mbar_init(cw_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4);
      for (int q = 0; q < 4; ++q) { mbar_init(&recv_full[q * 2 + i], 1); mbar_init(&send_credit[q * 2 + i], kC - 1); }
      mbar_init(&a_full[i], 1); mbar_init(&a_credit[i], kC); mbar_init(&a_sumdone[i], 1);
      mbar_init(&asum_ready[i], 1); mbar_init(&asum_free[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_free[i], 8);
      mbar_init(&ssq_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int k = threadIdx.x; k < KC; k += kThreads) {
    scale_s[k] = scale ? scale[k] : 1.0f;
    shift_s[k] = shift ? shift[k] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peers' barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  setmaxnreg_dec<56>();                          // one instruction for the whole warpgroup (.sync.aligned)
  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (elect_one()) {
      mbar_arrive_expect_tx(cw_full, nkb * C::kCwSubBytes);
      for (int kb = 0; kb < nkb; ++kb) tma_load_3d(cws + kb * C::kCwSubBytes, &tm_cw, cw_full, 0, 0, kb0 + kb, kEvictLast);
    }
    __syncwarp();
    const CUtensorMap* tmx = (static_cast<int>(rank) < kb_extra) ? &tm_xa : &tm_xb;     // box = this CTA's nkb feature blocks
    // L2 prefetch cursor: a load issued when its slot frees pays the full HBM latency (~0.9 us) inside a three-slot ring whose
    // slots live ~3.5 us; every load therefore also asks L2 for the tile kAhead tiles further down this CTA's stream.
    constexpr int kAhead = 3;
    const bool do_pf = !(dbg_flags & 65536);
    int itp = 0, ip = 0, ntp = n_iter > 0 ? vnt(0) : 0;
    auto prefetch_next = [&]() {
      if (itp >= n_iter) return;
      if (do_pf && elect_one()) tma_prefetch_4d(tmx, 0, ip * kF, kb0, vid(itp));
      __syncwarp();
      if (++ip == ntp) { ++itp; ip = 0; ntp = itp < n_iter ? vnt(itp) : 0; }
    };
    for (int j = 0; j < C::kSlots + kAhead; ++j) prefetch_next();
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int ntv = vnt(it);
      for (int i = 0; i < ntv; ++i, ++G) {
        const int slot = G % C::kSlots, u = G / C::kSlots;
        wait_bar_inl(&x_empty[slot], (u & 1) ^ 1u);
        if (lane == 0 && i < 8) NV5_T(it, 56 + i);
        if (elect_one()) {
          mbar_arrive_expect_tx(&x_full[slot], nkb * kSubBytes);
          tma_load_4d(xs + slot * kXSlotBytes, tmx, &x_full[slot], 0, i * kF, kb0, b, kEvictFirst);
        }
        __syncwarp();
        prefetch_next();
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer, phase 0 =============================
    // S[f, k] = X . Cw^T (K-major x K-major), M = 64 frames: frame 16 j + i lands in TMEM lane 32 j + i
    constexpr uint32_t idesc0 = make_idesc_bf16(64, KC, 0, 0);
    wait_bar_inl(cw_full, 0);
    int it0 = 0, i0 = 0, nt0 = n_iter > 0 ? vnt(0) : 0;       // (video, tile) of G, for the debug timeline only
    for (int G = 0; G < total_tiles; ++G) {
      const int sb = G % C::kSB, us = G / C::kSB, slot = G % C::kSlots;
      wait_bar_inl(&s_free[sb], (us & 1) ^ 1u);
      wait_bar_inl(&x_full[slot], (G / C::kSlots) & 1);
      tc_fence_after();
      if (lane == 0 && i0 < 8) NV5_T(it0, i0);
      if (elect_one()) {
        const uint32_t x_addr = smem_u32(xs + slot * kXSlotBytes);
        const uint32_t c_addr = smem_u32(cws);
        const uint32_t d_tmem = tmem_base + C::kSCol + sb * KC;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t adesc0 = make_sdesc_sw128(x_addr + kb * kSubBytes, 16, 1024);
          const uint64_t bdesc0 = make_sdesc_sw128(c_addr + kb * C::kCwSubBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, sdesc_advance(adesc0, k * 32), sdesc_advance(bdesc0, k * 32), idesc0, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
      if (++i0 == nt0) { ++it0; i0 = 0; nt0 = it0 < n_iter ? vnt(it0) : 0; }
    }
  } else if (warp == 2) {
    // =================================== MMA issuer, phase 1 =============================
    constexpr uint32_t idesc1 = make_idesc_bf16(128, KC, 1, 1);     // V^T += X^T . a    (MN-major x MN-major)
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int ntv = vnt(it);
      const int vb = it % C::kVB, uv = it / C::kVB;
      for (int i = 0; i < ntv; ++i, ++G) {
        const int ab = G % C::kAT, ua = G / C::kAT, slot = G % C::kSlots;
        if (elect_one()) mbar_arrive_expect_tx(&a_full[ab], C::kATileBytes);      // 64 rows from the four owners
        __syncwarp();
        wait_bar_cluster_inl(&a_full[ab], ua & 1);
        wait_bar_inl(&x_full[slot], (G / C::kSlots) & 1);               // (long complete: phase 0 of this tile ran on it)
        fence_proxy_async();                                         // st.async rows -> tcgen05 operand reads
        tc_fence_after();
        if (lane == 0 && i < 8) NV5_T(it, 32 + i);
        if (i == 0) {                                                // the epilogue has drained this accumulator (video it - kVB)
          wait_bar_inl(&v_free[vb], (uv & 1) ^ 1u);
          tc_fence_after();
        }
        if (elect_one()) {
          const uint32_t x_addr = smem_u32(xs + slot * kXSlotBytes);
          const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(atile + ab * C::kATileBytes), kSubBytes, 1024);
          for (int m = 0; m < nmb; ++m) {
            // rows m*128 .. +127 of this CTA's features = sub-tiles 2m and 2m+1, one box apart (LBO); the second sub-tile of a
            // half-valid last block is whatever follows in shared memory (its 64 accumulator rows are never read)
            const uint64_t adesc0 = make_sdesc_sw128(x_addr + m * 2 * kSubBytes, kSubBytes, 1024);
            const uint32_t d_tmem = tmem_base + C::kVCol + vb * C::kVStride + m * KC;
#pragma unroll
            for (int s = 0; s < kF / 16; ++s)
              umma_bf16(d_tmem, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1, (i > 0 || s > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
        if (lane == 0 && i < 8) NV5_T(it, 40 + i);
        wait_bar_inl(&a_sumdone[ab], ua & 1);                           // the a_sum warp has read the tile too
        if (elect_one()) {
          umma_commit(&x_empty[slot]);
          umma_commit_multicast(&a_credit[ab], 0xF);                 // every owner may overwrite tile ab in this CTA
          if (i == ntv - 1) umma_commit(&v_full[vb]);
        }
        __syncwarp();
      }
    }
  } else {
    // =================================== a_sum: column sums of the assignment tiles =============================
    // lane L owns clusters 2L, 2L+1 (+64 for the second atom): one 4-byte word per row, conflict free
    float acc[KC / 32];
#pragma unroll
    for (int j = 0; j < KC / 32; ++j) acc[j] = 0.0f;
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int ntv = vnt(it);
      const int p = it & 1;
      for (int i = 0; i < ntv; ++i, ++G) {
        const int ab = G % C::kAT, ua = G / C::kAT;
        wait_bar_cluster_inl(&a_full[ab], ua & 1);
        const uint8_t* at = atile + ab * C::kATileBytes;
#pragma unroll 4
        for (int f = 0; f < kF; ++f) {
#pragma unroll
          for (int a = 0; a < KC / 64; ++a) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(at + a * kSubBytes + f * 128 + ((((lane >> 2) ^ (f & 7))) << 4) + (lane & 3) * 4);
            acc[2 * a] += __uint_as_float(w << 16);
            acc[2 * a + 1] += __uint_as_float(w & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_sumdone[ab]);
      }
      wait_bar_inl(&asum_free[p], ((it >> 1) & 1) ^ 1u);                // the epilogue has consumed a_sum of video it-2
#pragma unroll
      for (int a = 0; a < KC / 64; ++a) {
        asum_s[p * KC + a * 64 + 2 * lane] = acc[2 * a];
        asum_s[p * KC + a * 64 + 2 * lane + 1] = acc[2 * a + 1];
        acc[2 * a] = 0.0f; acc[2 * a + 1] = 0.0f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&asum_ready[p]);
    }
  }
  } else if (warp < 8) {
    setmaxnreg_inc<136>();
    // ============================ exchange + softmax ============================
    // TMEM lane quadrant q holds frames 16 q .. 16 q + 15 of a tile in its lanes 0..15.  Frame 16 q + j belongs to CTA j / 4:
    // lane j pushes its 64 partial logits to that CTA's warp q (st.async), and THIS warp owns frames 16 q + 4 rank .. + 3 --
    // four frames per tile, softmaxed by 8 threads each (thread t: frame t / 8, clusters 8 (t % 8) .. + 7).
    const int q = warp & 3;
    const int jj = lane >> 3, part = lane & 7;                        // owner role: frame 4 rank + jj of the quadrant, 16-byte chunk `part`
    uint8_t* myrecv = recvbuf + q * C::kRecvWarp;                     // [rb] x [4 sources][4 frames][KC floats]
    const int dst_cta = lane >> 2;                                    // sender role (lanes 0..15): the owner of my frame
    const bool sender = lane < 16 && static_cast<uint32_t>(dst_cta) != rank;
    const int src_slot = static_cast<int>((rank - dst_cta - 1) & 3u); // my slot in the owner's buffer (0..2)
    const uint32_t recv_remote = mapa_u32(smem_u32(myrecv), dst_cta & 3) + (src_slot * 4 + (lane & 3)) * C::kRecvRow;
    const uint32_t recv_full_remote = mapa_u32(smem_u32(&recv_full[q * 2]), dst_cta & 3);
    constexpr int kPer = KC / 8;                                      // clusters per softmax thread (8 threads per frame)
    float sc[kPer], sh[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) { sc[j] = scale_s[kPer * part + j]; sh[j] = shift_s[kPer * part + j]; }
    uint32_t a_tile_remote[kC], a_full_remote[kC];
#pragma unroll
    for (int d = 0; d < kC; ++d) {
      a_tile_remote[d] = mapa_u32(smem_u32(atile), d);
      a_full_remote[d] = mapa_u32(smem_u32(a_full), d);
    }
    const int f = q * 16 + static_cast<int>(rank) * 4 + jj;           // the frame (inside a tile) this thread helps to softmax
    int it = 0, i = 0, ntv = n_iter > 0 ? vnt(0) : 0;
    int nf = n_iter > 0 ? min(max(__ldg(num_frames + vid(0)), 0), T) : 0;
    for (int G = 0; G < total_tiles; ++G) {
      const int sb = G % C::kSB, us = G / C::kSB, rb = G % C::kRB, ur = G / C::kRB, ab = G % C::kAT, ua = G / C::kAT;
      if (lane == 0) mbar_arrive_expect_tx(&recv_full[q * 2 + rb], (kC - 1) * 4 * C::kRecvRow);
      wait_bar_inl(&s_full[sb], us & 1);
      tc_fence_after();
      if (warp == 4 && lane == 0 && i < 8) NV5_T(it, 8 + i);
      // the three owners are done with what this warp pushed NRB tiles ago
      wait_bar_inl(&send_credit[q * 2 + rb], (ur & 1) ^ 1u);
      // 64 logit columns at a time: TMEM -> registers -> the frame's owner (own frames: source slot 3 of my buffer)
#pragma unroll 1
      for (int hc = 0; hc < KC / 64; ++hc) {
        float r[64];
        tmem_ld32(taddr_of(tmem_base, q) + C::kSCol + sb * KC + hc * 64, reinterpret_cast<uint32_t*>(r));
        tmem_ld32(taddr_of(tmem_base, q) + C::kSCol + sb * KC + hc * 64 + 32, reinterpret_cast<uint32_t*>(r) + 32);
        tmem_ld_wait();
        if (sender) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            st_async_v4(recv_remote + rb * C::kRecvBytes + hc * 256 + c * 16, recv_full_remote + rb * 8, r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        } else if (lane < 16) {
          float4* row = reinterpret_cast<float4*>(myrecv + rb * C::kRecvBytes + (3 * 4 + (lane & 3)) * C::kRecvRow + hc * 256);
#pragma unroll
          for (int c = 0; c < 16; ++c) row[c] = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);
      wait_bar_cluster_inl(&recv_full[q * 2 + rb], ur & 1);
      if (warp == 4 && lane == 0 && i < 8) NV5_T(it, 16 + i);
      // ---- owner role: sum the four partials of my kPer clusters, masked softmax over the frame's 8 threads ----
      float l[kPer];
      {
        const uint8_t* base = myrecv + rb * C::kRecvBytes + jj * C::kRecvRow + part * kPer * 4;
#pragma unroll
        for (int g4 = 0; g4 < kPer / 4; ++g4) {
          const float4 a0 = *reinterpret_cast<const float4*>(base + g4 * 16);
          l[4 * g4] = a0.x; l[4 * g4 + 1] = a0.y; l[4 * g4 + 2] = a0.z; l[4 * g4 + 3] = a0.w;
        }
#pragma unroll
        for (int sidx = 1; sidx < kC; ++sidx) {
#pragma unroll
          for (int g4 = 0; g4 < kPer / 4; ++g4) {
            const float4 b0 = *reinterpret_cast<const float4*>(base + sidx * 4 * C::kRecvRow + g4 * 16);
            l[4 * g4] += b0.x; l[4 * g4 + 1] += b0.y; l[4 * g4 + 2] += b0.z; l[4 * g4 + 3] += b0.w;
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        l[j] = l[j] * sc[j] + sh[j];
        mx = fmaxf(mx, l[j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      // every lane has consumed its loads of the buffer: the three senders may push their next partials into it.  Slot s
      // belongs to CTA (rank + 1 + s) & 3, whose warp q waits on send_credit[q][rb] in ITS shared memory.
      __syncwarp();
      if (lane < kC - 1) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&send_credit[q * 2 + rb]), (rank + 1 + lane) & 3u));
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        l[j] = __expf(l[j] - mx);
        sum += l[j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const float inv = 1.0f / sum;
      const bool valid = (i * kF + f) < nf;
      // select, not multiply: rows of frames >= num_frames may hold non-finite garbage
      uint32_t pk[kPer / 2];
#pragma unroll
      for (int j = 0; j < kPer / 2; ++j)
        pk[j] = pack_bf16x2(__float2bfloat16_rn(valid ? l[2 * j] * inv : 0.0f), __float2bfloat16_rn(valid ? l[2 * j + 1] * inv : 0.0f));
      // every CTA's phase 1 is done with assignment tile ab of NAT tiles ago
      wait_bar_inl(&a_credit[ab], (ua & 1) ^ 1u);
#pragma unroll
      for (int e8 = 0; e8 < kPer / 8; ++e8) {                         // my 16-byte chunks of the row: clusters 8 c .. 8 c + 7
        const int c = part * (kPer / 8) + e8;
        const uint32_t off = ab * C::kATileBytes + (c >> 3) * kSubBytes + sw128_offset(f, c & 7);
#pragma unroll
        for (int d = 0; d < kC; ++d)
          st_async_v4_b32(a_tile_remote[d] + off, a_full_remote[d] + ab * 8, pk[4 * e8], pk[4 * e8 + 1], pk[4 * e8 + 2], pk[4 * e8 + 3]);
      }
      if (warp == 4 && lane == 0 && i < 8) NV5_T(it, 24 + i);
      if (++i == ntv) {
        ++it; i = 0;
        ntv = it < n_iter ? vnt(it) : 0;
        nf = it < n_iter ? min(max(__ldg(num_frames + vid(it)), 0), T) : 0;
      }
    }
  } else {
    setmaxnreg_inc<160>();
    // ============================ epilogue: residual, norms, output (per video) ============================
    // warp e = (lane quadrant q, cluster half h): rows q*32 .. +31 of every 128-row accumulator block, clusters 32 h .. + 31.
    // The accumulator is read from TMEM once (and handed back to phase 1 right away); the corrected values stay in registers
    // across the exchange of the per-cluster norms.
    if constexpr (KC == 64) {
    const int e = warp - 8;
    const int q = e & 3, h = e >> 2;
    const int et = e * 32 + lane;                             // 0..255
    // accumulator blocks in which this warp's 32 rows exist (a prefix: only the last block can be half valid)
    const int nmb_w = (DH - q * 32 + 127) / 128 > 0 ? (DH - q * 32 + 127) / 128 : 0;
    const uint32_t tv = taddr_of(tmem_base, q) + C::kVCol + h * 32;        // this warp's V columns: + m * KC per block
    const uint32_t tc2 = taddr_of(tmem_base, q) + C::kC2Col + h * 32;      // and the matching block of cw2
    // ---- once: this warp's share of cw2 into TMEM (the residual then never touches global memory; loops below are ROLLED on
    //      purpose: sixteen warps in six roles share one instruction cache, straight-line unrolled bodies thrashed it) ----
#pragma unroll 1
    for (int m = 0; m < nmb_w; ++m) {
      // row-major cw2 [D, K]: lane = row, 8 chunks of 4 floats;  tiled: [D/32][K/4 chunks][32 rows][4 floats]
      const long long g32 = (d0 + m * 128 + q * 32) >> 5;
      const float4* c2 = TILED ? reinterpret_cast<const float4*>(cw2) + (g32 * (KC / 4) + h * 8) * 32 + lane
                               : reinterpret_cast<const float4*>(cw2 + (static_cast<long long>(d0) + m * 128 + q * 32 + lane) * KC + h * 32);
      constexpr int kCs = TILED ? 32 : 1;                     // float4 stride between consecutive chunks
      float c[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t4 = __ldg(c2 + j * kCs);
        c[4 * j] = t4.x; c[4 * j + 1] = t4.y; c[4 * j + 2] = t4.z; c[4 * j + 3] = t4.w;
      }
      tmem_st32(tc2 + m * KC, reinterpret_cast<const uint32_t*>(c));
    }
    tmem_st_wait();
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int p = it & 1;
      if (et == 0) mbar_arrive_expect_tx(&ssq_full[p], (kC - 1) * KC * 4);   // this video's partial sums from the three peers
      wait_bar_inl(&v_full[0], it & 1);
      tc_fence_after();
      if (et == 0) NV5_T(it, 48);
      // the accumulator leaves TMEM at once (3 x 32 columns per lane) and goes back to phase 1: the next video's aggregation
      // never waits for this epilogue
      float v[kMaxMb][32];
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m)
        if (m < nmb_w) tmem_ld32(tv + m * KC, reinterpret_cast<uint32_t*>(v[m]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_free[0]);
      if (et == 0) NV5_T(it, 52);
      wait_bar_inl(&asum_ready[p], (it >> 1) & 1);
      const float* asum = asum_s + p * KC + h * 32;
      float ssq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ssq[j] = 0.0f;
      // ---- pass 1 (registers): V -= a_sum * cw2 in fp32 (cw2 from TMEM, 16 columns at a time), per-cluster sum of squares ----
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m) {
        if (m < nmb_w) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float c[16];
            tmem_ld16(tc2 + m * KC + g * 16, reinterpret_cast<uint32_t*>(c));
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 as4 = *reinterpret_cast<const float4*>(asum + g * 16 + j4 * 4);     // broadcast read
              const int j = g * 16 + j4 * 4;
              v[m][j] -= as4.x * c[j4 * 4]; v[m][j + 1] -= as4.y * c[j4 * 4 + 1];
              v[m][j + 2] -= as4.z * c[j4 * 4 + 2]; v[m][j + 3] -= as4.w * c[j4 * 4 + 3];
              ssq[j] += v[m][j] * v[m][j]; ssq[j + 1] += v[m][j + 1] * v[m][j + 1];
              ssq[j + 2] += v[m][j + 2] * v[m][j + 2]; ssq[j + 3] += v[m][j + 3] * v[m][j + 3];
            }
          }
        }
      }
      if (et == 0) NV5_T(it, 54);
      ssq_w[e * 32 + lane] = warp_transpose_reduce32(ssq, lane);      // lane L: sum over this warp's rows of cluster 32 h + L
      named_bar_sync(1, 256);                                 // this CTA's partial sums are complete
      if (et == 0) NV5_T(it, 49);
      // ---- all-to-all of the KC partial sums between the four CTAs ----
      if (et < KC) {
        const int hh = et >> 5, kk = et & 31;
        const float mine = (ssq_w[(hh * 4 + 0) * 32 + kk] + ssq_w[(hh * 4 + 1) * 32 + kk]) + (ssq_w[(hh * 4 + 2) * 32 + kk] + ssq_w[(hh * 4 + 3) * 32 + kk]);
        ssq_part[(p * kC + rank) * KC + et] = mine;
#pragma unroll
        for (int s = 1; s < kC; ++s) {
          const uint32_t dst = (rank + s) & 3u;
          st_async_f32(mapa_u32(smem_u32(&ssq_part[(p * kC + rank) * KC + et]), dst), mapa_u32(smem_u32(&ssq_full[p]), dst), mine);
        }
      }
      wait_bar_cluster_inl(&ssq_full[p], (it >> 1) & 1);
      if (et == 0) NV5_T(it, 64);
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
      if (et == 0) NV5_T(it, 50);
      // ---- pass 2 (registers): rescale (intra-norm x final L2 norm), convert, store ----
#pragma unroll
      for (int m = 0; m < kMaxMb; ++m) {
        if (m < nmb_w) {
          // row-major [D, K]: 64 contiguous bytes per lane;  tiled: [D/32][K/8 chunks][32 rows][8 values]
          const long long g32 = (d0 + m * 128 + q * 32) >> 5;
          uint16_t* dst = TILED ? out + static_cast<long long>(b) * D * KC + ((g32 * (KC / 8) + h * 4) * 32 + lane) * 8
                                : out + (static_cast<long long>(b) * D + d0 + m * 128 + q * 32 + lane) * KC + h * 32;
          constexpr int kOs = TILED ? 256 : 8;                // element stride between consecutive 16-byte chunks
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
            *reinterpret_cast<uint4*>(dst + j8 * kOs) = hi;
          }
        }
      }
      if (et == 0) NV5_T(it, 51);
    }
    } else {
      // ---- K = 128: the accumulator (384 columns) and the logits fill TMEM, so cw2 comes from global memory (tiled: coalesced)
      //      and the corrected accumulator is written back to TMEM between the two passes.  warp e = (lane quadrant q, parity h):
      //      32-cluster column quarters h and h + 2. ----
      const int e = warp - 8;
      const int q = e & 3, h = e >> 2;
      const int et = e * 32 + lane;                           // 0..255
      const int nmb_w = (DH - q * 32 + 127) / 128 > 0 ? (DH - q * 32 + 127) / 128 : 0;
      const uint32_t tv = taddr_of(tmem_base, q) + C::kVCol;
      for (int it = 0; it < n_iter; ++it) {
        const int b = vid(it);
        const int p = it & 1;
        if (et == 0) mbar_arrive_expect_tx(&ssq_full[p], (kC - 1) * KC * 4);
        wait_bar_inl(&asum_ready[p], (it >> 1) & 1);
        wait_bar_inl(&v_full[0], it & 1);
        tc_fence_after();
        // ---- pass 1: V -= a_sum * cw2 (fp32), write back, per-cluster sum of squares ----
#pragma unroll 1
        for (int cq = h; cq < KC / 32; cq += 2) {
          float as[32], ssq[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t4 = *reinterpret_cast<const float4*>(asum_s + p * KC + cq * 32 + 4 * j);
            as[4 * j] = t4.x; as[4 * j + 1] = t4.y; as[4 * j + 2] = t4.z; as[4 * j + 3] = t4.w;
            ssq[4 * j] = 0.0f; ssq[4 * j + 1] = 0.0f; ssq[4 * j + 2] = 0.0f; ssq[4 * j + 3] = 0.0f;
          }
#pragma unroll 1
          for (int m = 0; m < nmb_w; ++m) {
            const long long g32 = (d0 + m * 128 + q * 32) >> 5;
            const float4* c2 = TILED ? reinterpret_cast<const float4*>(cw2) + (g32 * (KC / 4) + cq * 8) * 32 + lane
                                     : reinterpret_cast<const float4*>(cw2 + (static_cast<long long>(d0) + m * 128 + q * 32 + lane) * KC + cq * 32);
            constexpr int kCs = TILED ? 32 : 1;
            float4 cc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cc[j] = __ldg(c2 + j * kCs);
            float v[32];
            tmem_ld32(tv + m * KC + cq * 32, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] -= as[4 * j] * cc[j].x; v[4 * j + 1] -= as[4 * j + 1] * cc[j].y;
              v[4 * j + 2] -= as[4 * j + 2] * cc[j].z; v[4 * j + 3] -= as[4 * j + 3] * cc[j].w;
              ssq[4 * j] += v[4 * j] * v[4 * j]; ssq[4 * j + 1] += v[4 * j + 1] * v[4 * j + 1];
              ssq[4 * j + 2] += v[4 * j + 2] * v[4 * j + 2]; ssq[4 * j + 3] += v[4 * j + 3] * v[4 * j + 3];
            }
            tmem_st32(tv + m * KC + cq * 32, reinterpret_cast<const uint32_t*>(v));
          }
          ssq_w[(e * 2 + (cq >> 1)) * 32 + lane] = warp_transpose_reduce32(ssq, lane);
        }
        tmem_st_wait();
        named_bar_sync(1, 256);
        if (et < KC) {
          const int cq = et >> 5, kk = et & 31;                 // quarter cq lives in the warps of parity cq & 1, slot cq >> 1
          float mine = 0.0f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) mine += ssq_w[((((cq & 1) * 4 + qq) * 2) + (cq >> 1)) * 32 + kk];
          ssq_part[(p * kC + rank) * KC + et] = mine;
#pragma unroll
          for (int sidx = 1; sidx < kC; ++sidx) {
            const uint32_t dst = (rank + sidx) & 3u;
            st_async_f32(mapa_u32(smem_u32(&ssq_part[(p * kC + rank) * KC + et]), dst), mapa_u32(smem_u32(&ssq_full[p]), dst), mine);
          }
        }
        wait_bar_cluster_inl(&ssq_full[p], (it >> 1) & 1);
        named_bar_sync(1, 256);
        if (et < KC) {
          const float* sp = ssq_part + p * kC * KC + et;
          const float ss = (sp[0] + sp[KC]) + (sp[2 * KC] + sp[3 * KC]);
          const float rs = rsqrtf(fmaxf(ss, 1e-12f));
          fscale_s[et] = rs;
          contrib_s[et] = ss * rs * rs;
          if (stats && rank == 0) {
            stats[static_cast<long long>(b) * (2 * KC + 1) + et] = asum_s[p * KC + et];
            stats[static_cast<long long>(b) * (2 * KC + 1) + KC + et] = ss;
          }
        }
        named_bar_sync(1, 256);
        if (et == 0) mbar_arrive(&asum_free[p]);
        float tsum = 0.0f;
#pragma unroll
        for (int j = 0; j < KC / 32; ++j) tsum += contrib_s[j * 32 + lane];
        const float total = warp_sum(tsum);
        const float gs = rsqrtf(fmaxf(total, 1e-12f));
        if (stats && rank == 0 && et == 0) stats[static_cast<long long>(b) * (2 * KC + 1) + 2 * KC] = total;
        // ---- pass 2: rescale, convert, store ----
#pragma unroll 1
        for (int cq = h; cq < KC / 32; cq += 2) {
          float fs[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t4 = *reinterpret_cast<const float4*>(fscale_s + cq * 32 + 4 * j);
            fs[4 * j] = t4.x * gs; fs[4 * j + 1] = t4.y * gs; fs[4 * j + 2] = t4.z * gs; fs[4 * j + 3] = t4.w * gs;
          }
#pragma unroll 1
          for (int m = 0; m < nmb_w; ++m) {
            float v[32];
            tmem_ld32(tv + m * KC + cq * 32, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
            const long long g32 = (d0 + m * 128 + q * 32) >> 5;
            uint16_t* dst = TILED ? out + static_cast<long long>(b) * D * KC + ((g32 * (KC / 8) + cq * 4) * 32 + lane) * 8
                                  : out + (static_cast<long long>(b) * D + d0 + m * 128 + q * 32 + lane) * KC + cq * 32;
            constexpr int kOs = TILED ? 256 : 8;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= fs[j];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint4 hi, lo;
              if (out_f16) hi = pack8_f16(v + 8 * j8);
              else pack8_hi_lo(v + 8 * j8, hi, lo);
              *reinterpret_cast<uint4*>(dst + j8 * kOs) = hi;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_free[0]);
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

template <int KC, bool TILED>
int launch_v5(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed, const float* scale,
              const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats, cudaStream_t stream) {
  using C = Cfg<KC>;
  const int nkb_total = D / 64;
  const int nkb_hi = (nkb_total + kC - 1) / kC, nkb_lo = nkb_total / kC;
  CUtensorMap tm_xa, tm_xb, tm_cw;
  int rc;
  for (int v = 0; v < 2; ++v) {
    // X viewed as [B][D/64][T][64]: one box = 64 frames x the CTA's feature blocks (two box heights: ceil and floor of D/256)
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(nkb_total), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, 128, static_cast<uint64_t>(T) * D * 2};
    const uint32_t box[4] = {64, kF, static_cast<uint32_t>(v == 0 ? nkb_hi : nkb_lo), 1};
    if ((rc = make_tmap_bf16_nd(v == 0 ? &tm_xa : &tm_xb, x, 4, dims, strides, box)) != YT8M_OK) return rc;
  }
  {
    // Cw viewed as [D/64][KC clusters][64]: one box = one 64-feature block of all clusters
    const uint64_t dims[3] = {64, KC, static_cast<uint64_t>(nkb_total)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, 128};
    const uint32_t box[3] = {64, KC, 1};
    if ((rc = make_tmap_bf16_nd(&tm_cw, cw_packed, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  static int max_clusters = -1;
  if (max_clusters < 0) {
    YT8M_CUDA(cudaFuncSetAttribute(netvlad_v5_kernel<KC, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemTotal));
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(kC * 37, 1, 1);
    qc.blockDim = dim3(kThreads, 1, 1);
    qc.dynamicSmemBytes = C::kSmemTotal;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, netvlad_v5_kernel<KC, TILED>, &qc) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    max_clusters = n > 0 ? n : 32;              // (the query can fail under a profiler: 32 clusters always fit 148 SMs)
    if (max_clusters > 37) max_clusters = 37;
  }
  const int clusters = B < max_clusters ? B : max_clusters;
  netvlad_v5_kernel<KC, TILED><<<kC * clusters, kThreads, C::kSmemTotal, stream>>>(tm_xa, tm_xb, tm_cw, reinterpret_cast<uint16_t*>(out), num_frames, B, T,
                                                                            D, scale, shift, cw2, out_f16, stats, host_debug_timeline(),
                                                                            host_debug_flags());
  return check_launch("netvlad_v5_kernel");
}

}  // namespace

bool yt8m::netvlad_v5_supported(int T, int D, int K) {
  return (K == 64 || K == 128) && D % 64 == 0 && D / 64 >= kC && (D / 64 + kC - 1) / kC <= kMaxKb && T >= 1;
}

int yt8m::launch_netvlad_v5(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                            const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                            cudaStream_t stream) {
  YT8M_REQUIRE(netvlad_v5_supported(T, D, K), YT8M_E_BADSHAPE, "netvlad v5: T=%d D=%d K=%d", T, D, K);
  if (K == 128) return launch_v5<128, false>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out, out_f16, stats, stream);
  return launch_v5<64, false>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2, out, out_f16, stats, stream);
}

extern "C" int yt8m_netvlad_tiled_supported(int T, int D, int K) { return yt8m::netvlad_v5_supported(T, D, K) && D % 32 == 0 ? 1 : 0; }

namespace yt8m {
int launch_netvlad_v6(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2_tiled, yt8m_bf16* out_tiled, int out_f16, float* stats,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
bool netvlad_v6_supported(int T, int D, int K);
size_t netvlad_v6_workspace_bytes(int B, int T, int K);
}

extern "C" size_t yt8m_netvlad_tiled_workspace_bytes(int B, int T, int D, int K) {
  return yt8m::netvlad_v6_supported(T, D, K) ? yt8m::netvlad_v6_workspace_bytes(B, T, K) : 0;
}

extern "C" int yt8m_netvlad_fwd_tiled(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                                      const float* scale, const float* shift, const float* cw2_tiled, yt8m_bf16* out_tiled,
                                      int out_fmt, float* stats, void* workspace, size_t workspace_bytes, yt8m_stream_t stream_) {
  YT8M_REQUIRE(x && num_frames && cw_packed && cw2_tiled && out_tiled, YT8M_E_BADPTR, "yt8m_netvlad_fwd_tiled: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0, YT8M_E_BADSHAPE, "yt8m_netvlad_fwd_tiled: B=%d T=%d", B, T);
  YT8M_REQUIRE(yt8m_netvlad_tiled_supported(T, D, K), YT8M_E_UNSUPPORTED,
               "yt8m_netvlad_fwd_tiled: needs K in {64, 128} and D %% 64 == 0 with 256 <= D <= 1280 (D=%d K=%d)", D, K);
  YT8M_REQUIRE(out_fmt == YT8M_FMT_BF16 || out_fmt == YT8M_FMT_F16, YT8M_E_UNSUPPORTED, "yt8m_netvlad_fwd_tiled: out_fmt");
  YT8M_REQUIRE(aligned16(out_tiled) && aligned16(cw2_tiled), YT8M_E_BADPTR, "yt8m_netvlad_fwd_tiled: out / cw2 must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // two streaming kernels (assignment, then aggregation: yt8m_netvlad_v6.cu) when the caller provides the scratch for the
  // assignment; the one-pass four-CTA-cluster kernel otherwise (and for D > 1152, or with debug flag 1 << 21)
  if (workspace && yt8m::netvlad_v6_supported(T, D, K) && workspace_bytes >= yt8m::netvlad_v6_workspace_bytes(B, T, K) &&
      !(host_debug_flags() & (1 << 21)))
    return yt8m::launch_netvlad_v6(x, num_frames, B, T, D, K, cw_packed, scale, shift, cw2_tiled, out_tiled, out_fmt == YT8M_FMT_F16, stats,
                                   workspace, workspace_bytes, stream);
  if (K == 128)
    return launch_v5<128, true>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2_tiled, out_tiled, out_fmt == YT8M_FMT_F16, stats, stream);
  return launch_v5<64, true>(x, num_frames, B, T, D, cw_packed, scale, shift, cw2_tiled, out_tiled, out_fmt == YT8M_FMT_F16, stats, stream);
}
