// yt8m_b200 -- fused NetVLAD v4 (K = 64) for sm_100a: ONE pass over the frames.
//
// (NetVLAD is not part of /root/reference; definition: oracle/yt8m_oracle.py:netvlad_pool.  Same arithmetic as
//  yt8m_netvlad.cu, except that the residual  - a_sum[k] * cw2[d, k]  is subtracted in exact fp32.)
//
// v3 streams every video twice (assignment GEMM from HBM, aggregation GEMM again from L2) and parks the
// un-normalised descriptor in HBM before rescaling it: ~2.1 MB moved L2 -> SM per video for 0.84 MB of
// algorithmic traffic, two latency-bound passes per video (profiles/r01c_netvlad_v3_timeline.txt).  v4 removes both:
//
//   * a CLUSTER OF TWO CTAs owns a video and splits the feature axis: CTA r holds columns [r*D/2, (r+1)*D/2).
//     Its half of the assignment centres Cw (64 x D/2 bf16, 72 KB) stays RESIDENT in shared memory for the whole
//     (persistent) kernel; the video streams through a 3-slot ring of 32-frame tiles (36 KB, one TMA op each).
//   * per tile:  phase 0  S^T[k, f] = Cw_r . X_r^T   (UMMA M = 64 clusters, N = 32 frames)
//     gives each CTA the PARTIAL logits over its half of D.  Two "reader" warps move them TMEM -> registers,
//     push them into the peer CTA's shared memory (st.async: 16-byte DSMEM stores that complete a transaction
//     count on the peer's mbarrier -- no fences, no remote arrives), add the peer's partial in place, and four
//     softmax warps (4 threads per frame) produce the bf16 assignment tile.  Both CTAs compute the same softmax.
//   * phase 1  V_r^T[d, k] += X_r^T . a  reads the SAME shared-memory tile (MN-major operands), so a frame is read
//     from HBM exactly once and never again.  V_r^T (D/2 x 64 fp32 = 320 TMEM columns) stays in TMEM for the video.
//   * end of video: pass 1 subtracts the residual (cw2 rows fetched coalesced, transposed through a 4 KB per-warp
//     staging tile), writes the corrected value back to TMEM (tcgen05.st) and accumulates the per-cluster sum of
//     squares; the two CTAs exchange their 64 partial sums through DSMEM; pass 2 rescales (intra-norm x final
//     L2 norm), converts and writes the descriptor ONCE (transposed through the same per-warp tile: every store
//     instruction writes 512 contiguous bytes).  Accumulator blocks are handed back to the MMA warp one by one,
//     so the next video's aggregation starts behind the draining epilogue.
//   * the MMA warp is a small scheduler: it issues phase 0 of the next tiles or phase 1 of the oldest tile,
//     whichever has its operands ready (non-blocking mbarrier probes), so neither phase waits behind the other.
//
// Algorithmic traffic per video: 300*1152*2 B of frames in, 1152*64*2 B of descriptor out; cw2 (295 KB fp32)
// comes from L2.  Warp roles (512 threads, four warpgroups with their own register budgets): 0 = TMA producer,
// 1 = MMA issuer | 4-7 = epilogue | 8-11 = softmax | 12-15 = logit readers (one per TMEM lane quadrant).
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace yt8m {
int launch_netvlad_v4(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed,
                      const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                      cudaStream_t stream);
}

namespace {

constexpr int KC = 64;                       // clusters
constexpr int kFT = 32;                      // frames per tile
constexpr int kMaxKb = 9;                    // 64-wide feature blocks per CTA (D/2 <= 576)
constexpr int kMaxMb = 5;                    // 128-row accumulator blocks per CTA
constexpr int kSubBytes = kFT * 128;         // one 64-feature sub-tile of a frame tile: 32 rows x 128 B
constexpr int kXSlotBytes = kMaxKb * kSubBytes;     // 36 KB
constexpr int kSlots = 3;
constexpr int kCwSubBytes = KC * 128;        // one 64-feature block of the centres: 64 rows x 128 B
constexpr int kATileBytes = kFT * 128;       // assignment tile: 32 frames x 64 clusters bf16
constexpr int kPeerBytes = kFT * KC * 4;     // logits of one tile, fp32 frame-major
constexpr int kStageBytes = 128 * 128;       // output staging tile / 4 x 4 KB cw2 transposition tiles

constexpr int kOffCw = 0;
constexpr int kOffX = kOffCw + kMaxKb * kCwSubBytes;          //  72 KB
constexpr int kOffA = kOffX + kSlots * kXSlotBytes;           // 180 KB
constexpr int kOffPeer = kOffA + 2 * kATileBytes;             // 188 KB
constexpr int kOffStage = kOffPeer + 2 * kPeerBytes;          // 204 KB
constexpr int kOffSmall = kOffStage + kStageBytes;            // 220 KB
constexpr int kSmallFloats = 2 * KC /*scale, shift*/ + 2 * KC /*asum[2]*/ + KC /*ssq*/ + 2 * KC /*ssq_peer[2]*/ + KC /*fscale*/ +
                             KC /*contrib*/ + 4 * KC /*ssq per epilogue warp*/ + 4 * KC /*asum per softmax warp*/;
constexpr int kSmallBytes = kSmallFloats * 4 + 512;   // + barriers, TMEM slot, schedule (1 + 2 * kMaxIter ints)
constexpr int kSmemTotal = kOffSmall + kSmallBytes;
static_assert(kSmemTotal <= 227 * 1024, "NetVLAD v4 shared-memory budget exceeded");

constexpr int kSCol = 0;                     // TMEM: S^T double buffer, 2 x 32 columns
constexpr int kVCol = 64;                    // TMEM: V^T, kMaxMb x 64 columns
constexpr int kThreads = 512;
constexpr int kSms = 148;
constexpr int kMaxSchedB = 1024;             // batches up to this size get the longest-first schedule (O(B^2 / 512) per CTA)
constexpr int kMaxIter = 16;                 // >= ceil(kMaxSchedB / 74) videos per cluster

// debug-only phase timeline (globaltimer ns) of cluster 0 / CTA rank 0, first 3 videos, 128 stamps per video:
// [0,10) mma: tile landed   [10,20) reader: S ready   [20,30) softmax: logits ready   [30,40) mma: assignment ready
// [40,50) reader: peer partial landed   50 epi: video complete  51 pass 1 done  52 norms exchanged  53 pass 2 done
// [54,64) softmax: assignment written   [64,74) epi: pass-1 step done   [74,79) epi: pass-2 block done   80 softmax: a_sum done
#define NV4_T(itv, slot)                                                                                          \
  do {                                                                                                            \
    if (timeline && blockIdx.x == 0 && (itv) < 3) timeline[(itv) * 128 + (slot)] = global_timer_ns();              \
  } while (0)

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

// The exchanged logits of a tile live cluster-major: row k = 32 frames (128 B), whose 16-byte chunks are XOR-swizzled
// with (k & 7) -- the readers (one cluster per thread) move whole chunks without bank conflicts.
__device__ __forceinline__ uint32_t xch_chunk_off(int k, int c) {
  return static_cast<uint32_t>(k) * 128u + ((static_cast<uint32_t>(c) ^ (static_cast<uint32_t>(k) & 7u)) << 4);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
netvlad_v4_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_cw, uint16_t* __restrict__ out,
                  const int* __restrict__ num_frames, int B, int T, int D, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ cw2, int out_f16, float* __restrict__ stats,
                  unsigned long long* __restrict__ timeline, int dbg_flags) {
  // no static shared memory in this kernel: the dynamic window starts at the CTA's shared base (1024-byte aligned, checked
  // below), and pointers derived from the array keep their address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* cws = smem + kOffCw;
  uint8_t* xs = smem + kOffX;
  uint8_t* atile = smem + kOffA;
  uint8_t* peerbuf = smem + kOffPeer;
  uint8_t* stage = smem + kOffStage;
  float* scale_s = reinterpret_cast<float*>(smem + kOffSmall);
  float* shift_s = scale_s + KC;
  float* asum_s = shift_s + KC;                  // [2][KC] by video parity
  float* ssq_s = asum_s + 2 * KC;                // [KC] this CTA's partial sums of squares
  float* ssq_peer = ssq_s + KC;                  // [2][KC] written by the peer CTA
  float* fscale_s = ssq_peer + 2 * KC;           // [KC]
  float* contrib_s = fscale_s + KC;              // [KC]
  float* ssq_w = contrib_s + KC;                 // [4][KC] per epilogue warp (fp32 shared atomics are CAS loops on sm_100)
  float* asum_w = ssq_w + 4 * KC;                // [4][KC] per softmax warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(asum_w + 4 * KC);
  uint64_t* cw_full = bars;                      // [1]
  uint64_t* x_full = cw_full + 1;                // [3]
  uint64_t* x_empty = x_full + 3;                // [3]
  uint64_t* s_full = x_empty + 3;                // [2]  MMA -> readers
  uint64_t* s_free = s_full + 2;                 // [2]  readers -> MMA
  uint64_t* peer_full = s_free + 2;              // [2]  the peer's partial logits have landed (st.async transaction bytes)
  uint64_t* peer_free = peer_full + 2;           // [2]  peer softmax -> my readers: the PEER's buffer may be overwritten (remote, 4)
  uint64_t* sum_ready = peer_free + 2;           // [2]  readers -> softmax (4)
  uint64_t* a_ready = sum_ready + 2;             // [2]  softmax -> MMA (4)
  uint64_t* a_free = a_ready + 2;                // [2]  MMA -> softmax
  uint64_t* asum_ready = a_free + 2;             // [2]  softmax -> epilogue, by video parity (4): with one- and two-tile videos the
                                                 //      softmax warps finish video it+1 before the epilogue has looked at video it
  uint64_t* v_full = asum_ready + 2;             // [1]  MMA -> epilogue, per video
  uint64_t* v_free = v_full + 1;                 // [5]  epilogue -> MMA, per accumulator block and video (4)
  uint64_t* ssq_full = v_free + kMaxMb;          // [1]  the peer's 64 partial sums have landed (st.async bytes); [2] by video parity
  uint64_t* asum_free = ssq_full + 2;            // [2]  epilogue -> softmax, by video parity: a_sum of video it-2 has been consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(asum_free + 2);
  int* sched = reinterpret_cast<int*>(tmem_slot + 2);   // [0] = videos of this cluster, [1 + w] = video of wave w, [1 + kMaxIter + w] = its tiles

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;
  const int n_clusters = static_cast<int>(gridDim.x) >> 1;
  const int cid = static_cast<int>(blockIdx.x) >> 1;
  const int NT = (T + kFT - 1) / kFT;
  // PADDED FRAMES ARE NOT READ: a video is streamed for ceil(num_frames / 32) tiles only (at least one, which also
  // zero-initialises the accumulators of an empty video).  Videos are handed to the clusters longest first in
  // serpentine order (wave w: cluster c takes rank w*C + c, or w*C + C-1-c when w is odd), so the per-cluster sums of
  // tiles stay within one video of each other; every CTA derives the same schedule from num_frames on its own.
  auto tiles_of = [&](int nfv) { return min(max((min(nfv, T) + kFT - 1) / kFT, 1), NT); };
  const bool use_list = B <= kMaxSchedB;
  if (use_list) {
    if (threadIdx.x == 0) sched[0] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += kThreads) {
      const int ntb = tiles_of(__ldg(num_frames + b));
      int r = 0;
      for (int j = 0; j < B; ++j) {
        const int ntj = tiles_of(__ldg(num_frames + j));
        r += (ntj > ntb || (ntj == ntb && j < b)) ? 1 : 0;
      }
      const int w = r / n_clusters, pos = r - w * n_clusters;
      if (((w & 1) ? n_clusters - 1 - pos : pos) == cid) {
        sched[1 + w] = b;
        sched[1 + kMaxIter + w] = ntb;
        atomicMax(&sched[0], w + 1);
      }
    }
    __syncthreads();
  }
  const int n_iter = use_list ? sched[0] : (B - cid + n_clusters - 1) / n_clusters;
  auto vid = [&](int it) { return use_list ? sched[1 + it] : cid + it * n_clusters; };
  auto vnt = [&](int it) { return use_list ? sched[1 + kMaxIter + it] : tiles_of(__ldg(num_frames + cid + it * n_clusters)); };
  const int nkb = D / 128;                      // 64-wide feature blocks of this CTA's half
  const int DH = nkb * 64;
  const int nmb = (nkb + 1) >> 1;               // 128-row accumulator blocks (the last one may be half valid)
  int total_tiles = 0;
  for (int it = 0; it < n_iter; ++it) total_tiles += vnt(it);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_cw);
    mbar_init(cw_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4);
      mbar_init(&peer_full[i], 1); mbar_init(&peer_free[i], 4);
      mbar_init(&sum_ready[i], 4);
      mbar_init(&a_ready[i], 4); mbar_init(&a_free[i], 1);
    }
    mbar_init(&asum_ready[0], 4); mbar_init(&asum_ready[1], 4);
    mbar_init(v_full, 1);
    for (int i = 0; i < kMaxMb; ++i) mbar_init(&v_free[i], 4);
    mbar_init(&ssq_full[0], 1); mbar_init(&ssq_full[1], 1);
    mbar_init(&asum_free[0], 1); mbar_init(&asum_free[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int k = threadIdx.x; k < KC; k += kThreads) {
    scale_s[k] = scale ? scale[k] : 1.0f;
    shift_s[k] = shift ? shift[k] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  setmaxnreg_dec<96>();                          // one instruction for the whole warpgroup (.sync.aligned)
  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (elect_one()) {
      mbar_arrive_expect_tx(cw_full, nkb * kCwSubBytes);
      tma_load_3d(cws, &tm_cw, cw_full, 0, 0, static_cast<int>(rank) * nkb, kEvictLast);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    // L2 prefetch cursor: the ring holds only three tiles (a slot lives ~3 us from landing to the end of its phase 1), so a
    // load issued when its slot frees pays the full HBM latency.  Every load therefore also asks L2 for the tile
    // kAhead tiles further down this CTA's stream; the later shared-memory load then hits L2.
    constexpr int kAhead = 3;
    const bool do_pf = !(dbg_flags & 65536);
    int itp = 0, ip = 0, ntp = n_iter > 0 ? vnt(0) : 0;
    auto prefetch_next = [&]() {
      if (itp >= n_iter) return;
      if (do_pf && elect_one()) tma_prefetch_4d(&tm_x, 0, ip * kFT, static_cast<int>(rank) * nkb, vid(itp));
      __syncwarp();
      if (++ip == ntp) { ++itp; ip = 0; ntp = itp < n_iter ? vnt(itp) : 0; }
    };
    for (int j = 0; j < kSlots + kAhead; ++j) prefetch_next();      // the first tiles: (slots + kAhead) tiles ahead of consumption
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int ntv = vnt(it);
      for (int i = 0; i < ntv; ++i) {
        mbar_wait(&x_empty[slot], phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&x_full[slot], nkb * kSubBytes);
          tma_load_4d(xs + slot * kXSlotBytes, &tm_x, &x_full[slot], 0, i * kFT, static_cast<int>(rank) * nkb, b, kEvictFirst);
        }
        __syncwarp();
        prefetch_next();
        if (++slot == kSlots) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer, phase 0 =============================
    // S^T = Cw . X^T (K-major x K-major), M = 64: cluster 16 j + i lands in TMEM lane 32 j + i (half of every lane quadrant)
    constexpr uint32_t idesc0 = make_idesc_bf16(64, kFT, 0, 0);
    mbar_wait(cw_full, 0);
    // Tiles are numbered across this CTA's videos (G).
    // Phase 0 (this warp) and phase 1 (warp 2) are issued by different warps: neither waits behind the other's
    // barrier, and the tensor pipe executes the two instruction streams in arrival order.
    int it0 = 0, i0 = 0, nt0 = n_iter > 0 ? vnt(0) : 0;       // (video, tile) of G, for the debug timeline only
    for (int G = 0; G < total_tiles; ++G) {
      const int sb = G & 1, u = G >> 1, slot = G % kSlots;
      // S^T buffer drained by the readers (tile G-2) and X tile landed (the slot is recycled by phase 1 of tile G-3)
      mbar_wait(&s_free[sb], (u & 1) ^ 1u);
      mbar_wait(&x_full[slot], (G / kSlots) & 1);
      tc_fence_after();
      if (lane == 0 && i0 < 10) NV4_T(it0, i0);
      if (elect_one()) {
        const uint32_t x_addr = smem_u32(xs + slot * kXSlotBytes);
        const uint32_t c_addr = smem_u32(cws);
        const uint32_t d_tmem = tmem_base + kSCol + sb * kFT;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t adesc0 = make_sdesc_sw128(c_addr + kb * kCwSubBytes, 16, 1024);
          const uint64_t bdesc0 = make_sdesc_sw128(x_addr + kb * kSubBytes, 16, 1024);
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
      for (int i = 0; i < ntv; ++i, ++G) {
        const int ab = G & 1, u = G >> 1, slot = G % kSlots;
        mbar_wait(&a_ready[ab], u & 1);
        tc_fence_after();
        if (lane == 0 && i < 10) NV4_T(it, 30 + i);
        const int valid = min(kFT, T - i * kFT);
        const int nsteps = (valid + 15) >> 4;
        const uint32_t x_addr = smem_u32(xs + slot * kXSlotBytes);
        const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(atile + ab * kATileBytes), kSubBytes, 1024);
        const bool first = (i == 0 && it > 0);
        if (!first) {
          if (elect_one()) {
            for (int m = 0; m < nmb; ++m) {
              // rows m*128 .. +127 of this CTA's features = sub-tiles 2m and 2m+1, one box apart (LBO); the second
              // sub-tile of a half-valid last block is whatever follows in shared memory (its 64 accumulator rows
              // are never read)
              const uint64_t adesc0 = make_sdesc_sw128(x_addr + m * 2 * kSubBytes, kSubBytes, 1024);
              const uint32_t d_tmem = tmem_base + kVCol + m * KC;
              for (int s = 0; s < nsteps; ++s)
                umma_bf16(d_tmem, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1, (i > 0 || s > 0) ? 1u : 0u);
            }
          }
        } else {
          // first tile of a new video: follow the draining epilogue of the previous one block by block
          for (int m = 0; m < nmb; ++m) {
            mbar_wait(&v_free[m], (it - 1) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc0 = make_sdesc_sw128(x_addr + m * 2 * kSubBytes, kSubBytes, 1024);
              const uint32_t d_tmem = tmem_base + kVCol + m * KC;
              for (int s = 0; s < nsteps; ++s)
                umma_bf16(d_tmem, sdesc_advance(adesc0, s * 2048), sdesc_advance(bdesc0, s * 2048), idesc1, s > 0 ? 1u : 0u);
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          umma_commit(&x_empty[slot]);
          umma_commit(&a_free[ab]);
          if (i == ntv - 1) umma_commit(v_full);
        }
        __syncwarp();
      }
    }
  }
  } else if (warp >= 12) {
    setmaxnreg_dec<96>();
    // ============================ logit readers: TMEM -> peer CTA + in-place sum ============================
    const int q = warp & 3;                                   // lane quadrant q holds clusters 16 q .. 16 q + 15 in its lanes 0..15
    const int k = q * 16 + (lane & 15);
    const bool act = lane < 16;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kSCol;
    const uint32_t peer_buf_remote = mapa_u32(smem_u32(peerbuf), peer);
    const uint32_t peer_full_remote = mapa_u32(smem_u32(peer_full), peer);
    int itr = 0, ir = 0, ntr = n_iter > 0 ? vnt(0) : 0;       // (video, tile) of G, for the debug timeline only
    for (int G = 0; G < total_tiles; ++G) {
      const int sb = G & 1, u = G >> 1;
      const int tl_it = itr, tl_i = ir;
      if (++ir == ntr) { ++itr; ir = 0; ntr = itr < n_iter ? vnt(itr) : 0; }
      // this tile's incoming partial: 8 KB of st.async transaction bytes from the peer's readers
      if (warp == 12 && lane == 0) mbar_arrive_expect_tx(&peer_full[sb], kPeerBytes);
      mbar_wait(&s_full[sb], u & 1);
      tc_fence_after();
      if (warp == 12 && lane == 0 && tl_i < 10) NV4_T(tl_it, 10 + tl_i);
      float r[kFT];
      tmem_ld32(taddr + sb * kFT, reinterpret_cast<uint32_t*>(r));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);
      // the peer's softmax warps are done with what we pushed two tiles ago
      mbar_wait(&peer_free[sb], (u & 1) ^ 1u);
      if (act) {
#pragma unroll
        for (int c = 0; c < kFT / 4; ++c)
          st_async_v4(peer_buf_remote + sb * kPeerBytes + xch_chunk_off(k, c), peer_full_remote + sb * 8, r[4 * c], r[4 * c + 1],
                      r[4 * c + 2], r[4 * c + 3]);
      }
      // the peer's partial logits have landed in OUR buffer: add ours in place
      mbar_wait(&peer_full[sb], u & 1);
      if (warp == 12 && lane == 0 && tl_i < 10) NV4_T(tl_it, 40 + tl_i);
      uint8_t* pb = peerbuf + sb * kPeerBytes;
      if (act) {
#pragma unroll
        for (int c = 0; c < kFT / 4; ++c) {
          float4* p = reinterpret_cast<float4*>(pb + xch_chunk_off(k, c));
          float4 t4 = *p;
          t4.x += r[4 * c]; t4.y += r[4 * c + 1]; t4.z += r[4 * c + 2]; t4.w += r[4 * c + 3];
          *p = t4;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sum_ready[sb]);
    }
  } else if (warp < 8) {
    setmaxnreg_inc<216>();
    // ============================ epilogue: residual, norms, output (per video) ============================
    const int q = warp & 3;                                   // TMEM lane quadrant
    const int wi = warp - 4;                                  // staging tile of this warp
    const int et = wi * 32 + lane;                            // 0..127
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kVCol;
    uint8_t* wst = stage + wi * 4096;
    const uint32_t ssq_peer_remote = mapa_u32(smem_u32(ssq_peer), peer);
    const uint32_t ssq_full_remote = mapa_u32(smem_u32(ssq_full), peer);
    // accumulator blocks in which this warp's 32 rows exist (a prefix: only the last block can be half valid)
    const int nmb_w = (DH - q * 32 + 127) / 128 > 0 ? (DH - q * 32 + 127) / 128 : 0;
    const int nsteps1 = 2 * nmb_w;                            // pass 1 = two sweeps (32 clusters each) over those blocks
    // coalesced fetch of the 32 x 32 fp32 block cw2[gd0 .. +32, h*32 .. +32): 4 rows (of 128 B) per instruction
    auto fetch_c2 = [&](int s, float4* g) {
      const int h = s >= nmb_w ? 1 : 0, m = s - h * nmb_w;
      const long long gd0 = static_cast<long long>(rank) * DH + m * 128 + q * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + (lane >> 3), ch = lane & 7;
        g[j] = __ldg(reinterpret_cast<const float4*>(cw2 + (gd0 + r) * KC + h * 32 + ch * 4));
      }
    };
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int p = it & 1;
      // the cw2 blocks of the first two pass-1 steps are in flight while we wait for the video to complete; after that the
      // fetch runs two steps ahead of its use (L2 latency ~ two steps of work)
      float4 ga[8], gb[8];
      if (nsteps1 > 0) { fetch_c2(0, ga); fetch_c2(1, gb); }
      if (et == 0) mbar_arrive_expect_tx(&ssq_full[p], KC * 4);   // this video's 64 partial sums from the peer
      if (nsteps1 == 0) { ssq_w[wi * KC + lane] = 0.0f; ssq_w[wi * KC + 32 + lane] = 0.0f; }
      mbar_wait(&asum_ready[p], (it >> 1) & 1);
      mbar_wait(v_full, it & 1);
      tc_fence_after();
      if (et == 0) NV4_T(it, 50);
      const float* asum = asum_s + p * KC;
      // ---- pass 1: V -= a_sum * cw2 (fp32), write back, per-cluster sum of squares ----
      float ssq[32], as[32];
      auto step1 = [&](int s, float4 (&g)[8]) {
        const int h = s >= nmb_w ? 1 : 0, m = s - h * nmb_w;
        if (m == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { ssq[j] = 0.0f; as[j] = asum[h * 32 + j]; }
        }
        float v[32];
        tmem_ld32(tlane + m * KC + h * 32, reinterpret_cast<uint32_t*>(v));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 4 * j + (lane >> 3), ch = lane & 7;
          *reinterpret_cast<float4*>(wst + r * 128 + ((ch ^ (r & 7)) << 4)) = g[j];
        }
        __syncwarp();
        if (s + 2 < nsteps1) fetch_c2(s + 2, g);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 c2 = *reinterpret_cast<const float4*>(wst + lane * 128 + ((c ^ (lane & 7)) << 4));
          v[4 * c + 0] -= as[4 * c + 0] * c2.x;
          v[4 * c + 1] -= as[4 * c + 1] * c2.y;
          v[4 * c + 2] -= as[4 * c + 2] * c2.z;
          v[4 * c + 3] -= as[4 * c + 3] * c2.w;
        }
        __syncwarp();                                         // the staging tile may be overwritten
        tmem_st32(tlane + m * KC + h * 32, reinterpret_cast<const uint32_t*>(v));
#pragma unroll
        for (int j = 0; j < 32; ++j) ssq[j] += v[j] * v[j];
        if (m == nmb_w - 1) {
          const float tot = warp_transpose_reduce32(ssq, lane);
          ssq_w[wi * KC + h * 32 + lane] = tot;
        }
        if (et == 0 && s < 10) NV4_T(it, 64 + s);
      };
#pragma unroll 1
      for (int s = 0; s < nsteps1; s += 2) {                  // nsteps1 is even
        step1(s, ga);
        step1(s + 1, gb);
      }
      tmem_st_wait();
      named_bar_sync(1, 128);                                 // this CTA's partial sums are complete
      if (et == 0) NV4_T(it, 51);
      // ---- exchange the 64 partial sums with the peer CTA ----
      if (et < KC) {
        ssq_s[et] = (ssq_w[et] + ssq_w[KC + et]) + (ssq_w[2 * KC + et] + ssq_w[3 * KC + et]);
        st_async_f32(ssq_peer_remote + (p * KC + et) * 4, ssq_full_remote + p * 8, ssq_s[et]);
      }
      mbar_wait(&ssq_full[p], (it >> 1) & 1);
      if (et < KC) {
        const float ss = ssq_s[et] + ssq_peer[p * KC + et];
        const float rs = rsqrtf(fmaxf(ss, 1e-12f));
        fscale_s[et] = rs;
        contrib_s[et] = ss * rs * rs;
        if (stats && rank == 0) {                             // saved for the backward pass: a_sum, ||V_k||^2
          stats[static_cast<long long>(b) * (2 * KC + 1) + et] = asum[et];
          stats[static_cast<long long>(b) * (2 * KC + 1) + KC + et] = ss;
        }
      }
      named_bar_sync(1, 128);
      // every epilogue thread is done with asum_s[p]: the softmax warps may publish video it+2 into it (short videos --
      // one or two tiles -- let them run that far ahead)
      if (et == 0) mbar_arrive(&asum_free[p]);
      const float total = warp_sum(contrib_s[lane] + contrib_s[lane + 32]);     // same tree on every warp of both CTAs
      const float gs = rsqrtf(fmaxf(total, 1e-12f));
      if (stats && rank == 0 && et == 0) stats[static_cast<long long>(b) * (2 * KC + 1) + 2 * KC] = total;
      if (et == 0) NV4_T(it, 52);
      // ---- pass 2: rescale, convert, transpose through the warp's staging tile, store 512 contiguous bytes per
      //      instruction; accumulator blocks go back to the MMA warp one by one ----
      float fs[KC];
#pragma unroll
      for (int j = 0; j < KC; ++j) fs[j] = fscale_s[j] * gs;
#pragma unroll 1
      for (int m = 0; m < nmb; ++m) {
        const bool active = m < nmb_w;
        if (active) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[32];
            tmem_ld32(tlane + m * KC + h * 32, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= fs[h * 32 + j];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint4 hi, lo;
              if (out_f16) hi = pack8_f16(v + 8 * j8);
              else pack8_hi_lo(v + 8 * j8, hi, lo);
              *reinterpret_cast<uint4*>(wst + lane * 128 + (((h * 4 + j8) ^ (lane & 7)) << 4)) = hi;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_free[m]);
        if (active) {
          // rows q*32 .. +31 of block m are 32 x 128 B = 4 KB contiguous in the output
          uint16_t* dst = out + (static_cast<long long>(b) * D + static_cast<long long>(rank) * DH + m * 128 + q * 32) * KC;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = 4 * j + (lane >> 3), ch = lane & 7;
            const uint4 t4 = *reinterpret_cast<const uint4*>(wst + r * 128 + ((ch ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(dst + r * KC + ch * 8) = t4;
          }
          __syncwarp();                                       // the staging tile may be overwritten
        }
        if (et == 0 && m < 5) NV4_T(it, 74 + m);
      }
      if (et == 0) NV4_T(it, 53);
    }
  } else {
    setmaxnreg_dec<104>();                        // warps 8-11
    // ============================ softmax: 4 threads per frame, 16 clusters each ============================
    const int sw = warp - 8;                                  // 0..3
    const int st = sw * 32 + lane;                            // 0..127
    const int f = st >> 2, q4 = st & 3;
    const uint32_t peer_free_remote = mapa_u32(smem_u32(peer_free), peer);
    float sc[16], sh[16], acc[16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sc[4 * j + i] = scale_s[4 * (q4 + 4 * j) + i];
        sh[4 * j + i] = shift_s[4 * (q4 + 4 * j) + i];
        acc[4 * j + i] = 0.0f;
      }
    int G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = vid(it);
      const int nf = min(max(num_frames[b], 0), T);
      const int ntv = vnt(it);
      for (int i = 0; i < ntv; ++i, ++G) {
        const int sb = G & 1, u = G >> 1;
        mbar_wait(&sum_ready[sb], u & 1);
        if (st == 0 && i < 10) NV4_T(it, 20 + i);
        const uint8_t* pb = peerbuf + sb * kPeerBytes;
        float l[16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int k = 4 * (q4 + 4 * j) + i4;
            l[4 * j + i4] = *reinterpret_cast<const float*>(pb + xch_chunk_off(k, f >> 2) + (f & 3) * 4);
          }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          l[j] = l[j] * sc[j] + sh[j];
          mx = fmaxf(mx, l[j]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        // every lane has consumed its loads of the logits buffer: the peer may push its next partial into it
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(peer_free_remote + sb * 8);
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          l[j] = __expf(l[j] - mx);
          sum += l[j];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.0f / sum;
        const bool valid = (i * kFT + f) < nf;
        mbar_wait(&a_free[sb], (u & 1) ^ 1u);                 // phase 1 of two tiles ago has read this assignment tile
        uint8_t* at = atile + sb * kATileBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = q4 + 4 * j;                           // clusters 4c .. 4c+3
          // select, not multiply: rows of frames >= num_frames may hold non-finite garbage
          const __nv_bfloat16 h0 = __float2bfloat16_rn(valid ? l[4 * j + 0] * inv : 0.0f);
          const __nv_bfloat16 h1 = __float2bfloat16_rn(valid ? l[4 * j + 1] * inv : 0.0f);
          const __nv_bfloat16 h2 = __float2bfloat16_rn(valid ? l[4 * j + 2] * inv : 0.0f);
          const __nv_bfloat16 h3 = __float2bfloat16_rn(valid ? l[4 * j + 3] * inv : 0.0f);
          acc[4 * j + 0] += __bfloat162float(h0);             // a_sum uses the rounded assignment too
          acc[4 * j + 1] += __bfloat162float(h1);
          acc[4 * j + 2] += __bfloat162float(h2);
          acc[4 * j + 3] += __bfloat162float(h3);
          *reinterpret_cast<uint2*>(at + sw128_offset(f, c >> 1) + (c & 1) * 8) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[sb]);
        if (st == 0 && i < 10) NV4_T(it, 54 + i);
      }
      // ---- a_sum of this video: reduce over the 32 threads that share q4 (8 per warp, 4 warps) ----
      const int p = it & 1;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = acc[j];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        acc[j] = a;
      }
      if (lane < 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) asum_w[sw * KC + 4 * (q4 + 4 * j) + i] = acc[4 * j + i];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
      if (it >= 2) mbar_wait(&asum_free[p], ((it >> 1) - 1) & 1);   // the epilogue has consumed a_sum of video it-2
      named_bar_sync(2, 128);
      if (st < KC) asum_s[p * KC + st] = (asum_w[st] + asum_w[KC + st]) + (asum_w[2 * KC + st] + asum_w[3 * KC + st]);
      named_bar_sync(2, 128);
      if (lane == 0) mbar_arrive(&asum_ready[p]);
      if (st == 0) NV4_T(it, 80);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody exits while its peer may still touch its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int yt8m::launch_netvlad_v4(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, const yt8m_bf16* cw_packed,
                            const float* scale, const float* shift, const float* cw2, yt8m_bf16* out, int out_f16, float* stats,
                            cudaStream_t stream) {
  YT8M_REQUIRE(D % 128 == 0 && D / 128 >= 1 && D / 128 <= kMaxKb, YT8M_E_BADSHAPE, "netvlad v4: D=%d (need D %% 128 == 0, D <= %d)", D,
               kMaxKb * 128);
  const int nkb = D / 128;
  CUtensorMap tm_x, tm_cw;
  int rc;
  {
    // X viewed as [B][D/64][T][64]: one box = 32 frames x this CTA's D/128 feature blocks
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(D / 64), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 2, 128, static_cast<uint64_t>(T) * D * 2};
    const uint32_t box[4] = {64, kFT, static_cast<uint32_t>(nkb), 1};
    if ((rc = make_tmap_bf16_nd(&tm_x, x, 4, dims, strides, box)) != YT8M_OK) return rc;
  }
  {
    // Cw viewed as [D/64][64 clusters][64]
    const uint64_t dims[3] = {64, KC, static_cast<uint64_t>(D / 64)};
    const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, 128};
    const uint32_t box[3] = {64, KC, static_cast<uint32_t>(nkb)};
    if ((rc = make_tmap_bf16_nd(&tm_cw, cw_packed, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(netvlad_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_done = true;
  }
  const int clusters = B < kSms / 2 ? B : kSms / 2;
  netvlad_v4_kernel<<<2 * clusters, kThreads, kSmemTotal, stream>>>(tm_x, tm_cw, reinterpret_cast<uint16_t*>(out), num_frames, B, T, D, scale, shift, cw2,
                                                                    out_f16, stats, host_debug_timeline(), host_debug_flags());
  return check_launch("netvlad_v4_kernel");
}
