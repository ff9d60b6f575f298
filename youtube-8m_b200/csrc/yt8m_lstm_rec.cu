// yt8m_b200 -- persistent LSTM recurrence for sm_100a: one launch runs all T steps of one layer.
//
// Replaces the per-step launches of tf.nn.dynamic_rnn(BasicLSTMCell) -- wh/all_frame_models/lstm_model.py:30-47,
// lstm_memory_model.py:47-61 -- for H in {256, 512, 768, 1024}.  The input projection (x_t . Wx + b for every t) is
// hoisted into one big tensor-core GEMM by the caller (yt8m_lstm_fwd); what is left per step is
//     G_t[b, 4H] = xw[b, t, :] + h_{t-1}[b, :] . Wh          (M = 4H gate rows, N = batch, K = H)
// followed by the cell update.  Launching that 300 times per layer moves Wh (8 MB) and the state through L2 every
// step and is bound by per-SM ingest (measured 15 us per step on 32 CTAs).  Here:
//
//   * WEIGHT-STATIONARY: a cluster of four CTAs owns 128 gate rows (32 cells, unit-major packing
//     [i_u, j_u, f_u, o_u]); CTA `rank` of the cluster holds the K-quarter [rank*H/4, (rank+1)*H/4) of those rows
//     in shared memory (<= 64 KB, loaded once by TMA).  H/32 clusters cover the layer: 128 CTAs at H = 1024.
//   * per step every CTA loads only its K-quarter of h_{t-1} (bf16 hi + lo, <= 64 batch rows: 64 KB in one TMA
//     transaction pair), runs the UMMAs (A = weights, M = 128; B = h tile, N = 64; hi and lo into the same fp32
//     accumulator) and owns a PARTIAL G over its K-quarter.
//   * the partials are reduce-scattered through DSMEM: CTA q of the cluster finalises batch columns
//     [16q, 16q+16); every CTA pushes the matching 16 columns of its TMEM tile into q's shared memory with
//     st.async (the data completes a transaction count on q's mbarrier: no fences, no remote arrives).
//   * the cell update is then thread-local (one thread = one cell x four batch rows; c and h live in REGISTERS for
//     the whole sequence), h_t is written as bf16 hi/lo into h_seq[b, t, :] -- which is at the same time the
//     layer's output sequence, the next layer's GEMM operand, and next step's recurrent operand.
//   * steps are separated by one grid-wide barrier (a monotonic counter in global memory; all CTAs are
//     co-resident: cooperative launch).  Writers: st.global -> fence.proxy.async -> bar -> fence + atomic;
//     reader (the TMA warp): ld.acquire -> fence.proxy.async -> TMA load.
//
// dynamic_rnn(sequence_length) semantics: for t >= num_frames[b] the output row is zero and the state row is
// carried.  A frozen row stays frozen, so zeros are what the recurrence reads back for it (never used).
//
// Per step the layer moves B x 4H fp32 of xw (HBM, read once) and n_cta x 64 KB of h (L2); the roofline that
// bounds it is latency: barrier + TMA + MMA + DSMEM + cell update are serial by data dependence.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace {

constexpr int kKS = 4;                         // CTAs per cluster = K split
constexpr int kNB = 64;                        // batch columns per pass (UMMA N)
constexpr int kMaxKb = 4;                      // 64-wide k-blocks per CTA (H <= 1024)
constexpr int kWTileBytes = 128 * 128;         // 128 gate rows x 64 bf16
constexpr int kHTileBytes = kNB * 128;         // 64 batch rows x 64 bf16
constexpr int kOffW = 0;
constexpr int kOffHhi = kOffW + kMaxKb * kWTileBytes;        //  64 KB
constexpr int kOffHlo = kOffHhi + kMaxKb * kHTileBytes;      //  96 KB
constexpr int kOffRecv = kOffHlo + kMaxKb * kHTileBytes;     // 128 KB
constexpr int kRecvSrcBytes = 4 * 32 * 64;                   // one source CTA: [bq 4][cell 32][slot 4] float4 = 8 KB
constexpr int kRecvBytes = kKS * kRecvSrcBytes;              // 32 KB
constexpr int kOffBars = kOffRecv + kRecvBytes;              // 160 KB
constexpr int kSmemTotal = kOffBars + 64;
constexpr int kThreads = 192;                  // warp 0 = TMA + grid barrier, 1 = MMA, 2-5 = exchange + cell update

struct LstmRecParams {
  const float* xw;            // [Bc, T, 4H] fp32, unit-major gate columns, bias included
  const int* num_frames;      // [Bc]
  uint16_t* h_hi;             // [Bc, T, H] bf16
  uint16_t* h_lo;
  float* out_seq;             // nullable [Bc, T, H]
  float* c_out;               // final state, row stride ld_state
  float* h_out;
  long long ld_state;
  unsigned int* counter;      // zeroed before the launch
  unsigned long long* timeline;   // debug only (yt8m_debug_set_timeline): 16 stamps per step of CTA 0, steps 8..15
  int Bc, T, H;
  float forget_bias;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// release at gpu scope: the CTA's earlier global stores (ordered before this thread by bar.sync) are visible to whoever
// acquires the new counter value
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// orders generic-proxy global accesses with async-proxy (TMA) global accesses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// debug-only step timeline (globaltimer ns) of CTA 0, steps 8..15, 16 stamps per step:
// 0 tma: grid barrier passed   1 mma: h tiles landed   2 epi: accumulator complete   3 epi: partials pushed
// 4 epi: peers' partials landed   5 epi: h_t stored   6 epi: arrived on the grid barrier
#define LR_T(t, slot)                                                                                             \
  do {                                                                                                            \
    if (p.timeline && blockIdx.x == 0 && (t) >= 8 && (t) < 16) p.timeline[((t) - 8) * 16 + (slot)] = global_timer_ns(); \
  } while (0)

__global__ void __cluster_dims__(kKS, 1, 1) __launch_bounds__(kThreads, 1)
lstm_rec_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_hhi,
                const __grid_constant__ CUtensorMap tm_hlo, const LstmRecParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ws = smem + kOffW;
  uint8_t* hhi = smem + kOffHhi;
  uint8_t* hlo = smem + kOffHlo;
  uint8_t* recv = smem + kOffRecv;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* w_full = bars;
  uint64_t* h_full = bars + 1;
  uint64_t* acc_full = bars + 2;
  uint64_t* recv_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nt = static_cast<int>(blockIdx.x) / kKS;        // 128-row gate tile = cells [32 nt, 32 nt + 32)
  const int nkb = p.H / (64 * kKS);
  const unsigned int n_cta = gridDim.x;
  const int T = p.T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_hhi); tma_prefetch_desc(&tm_hlo);
    mbar_init(w_full, 1);
    mbar_init(h_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(recv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peers' barriers exist before anyone completes bytes on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer + grid barrier =================================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, nkb * kWTileBytes);
      tma_load_3d(ws, &tm_w, w_full, 0, nt * 128, static_cast<int>(rank) * nkb, kEvictLast);
    }
    __syncwarp();
    for (int t = 1; t < T; ++t) {
      // every CTA has written its slice of h_{t-1}
      const unsigned int target = static_cast<unsigned int>(t) * n_cta;
      if (ld_acquire_gpu(p.counter) < target) {
        const uint64_t t0 = global_timer_ns();
        uint32_t spins = 0;
        while (ld_acquire_gpu(p.counter) < target) {
          if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > YT8M_WAIT_TIMEOUT_NS) { __trap(); }
        }
      }
      fence_proxy_async_global();
      __syncwarp();
      if (lane == 0) LR_T(t, 0);
      if (elect_one()) {
        mbar_arrive_expect_tx(h_full, 2 * nkb * kHTileBytes);
        tma_load_4d(hhi, &tm_hhi, h_full, 0, 0, static_cast<int>(rank) * nkb, t - 1, kEvictNormal);
        tma_load_4d(hlo, &tm_hlo, h_full, 0, 0, static_cast<int>(rank) * nkb, t - 1, kEvictNormal);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer ==========================================
    constexpr uint32_t idesc = make_idesc_bf16(128, kNB, 0, 0);     // A = weights (K-major), B = h (K-major)
    mbar_wait(w_full, 0);
    for (int t = 1; t < T; ++t) {
      mbar_wait(h_full, (t - 1) & 1);
      tc_fence_after();
      if (lane == 0) LR_T(t, 1);
      if (elect_one()) {
        const uint32_t w_addr = smem_u32(ws), hi_addr = smem_u32(hhi), lo_addr = smem_u32(hlo);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t adesc = make_sdesc_sw128(w_addr + kb * kWTileBytes, 16, 1024);
          const uint64_t bhi = make_sdesc_sw128(hi_addr + kb * kHTileBytes, 16, 1024);
          const uint64_t blo = make_sdesc_sw128(lo_addr + kb * kHTileBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem_base, sdesc_advance(adesc, k * 32), sdesc_advance(bhi, k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_bf16(tmem_base, sdesc_advance(adesc, k * 32), sdesc_advance(blo, k * 32), idesc, 1u);
          }
        }
        umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ============================ exchange + cell update (128 threads) ===========================
    const int ew = warp - 2;                     // 0..3
    const int q = warp & 3;                      // TMEM lane quadrant this warp may read
    // exchange role: this thread holds gate row `row` of the tile for all 64 batch columns
    const int row = q * 32 + lane;
    const int c_r = row >> 2, g_r = row & 3;
    const uint32_t send_off = static_cast<uint32_t>(rank) * kRecvSrcBytes + static_cast<uint32_t>(c_r) * 64u +
                              (static_cast<uint32_t>(g_r ^ ((c_r >> 1) & 3)) << 4);
    uint32_t rbase[kKS], rbar[kKS];
#pragma unroll
    for (int d = 0; d < kKS; ++d) {
      rbase[d] = mapa_u32(smem_u32(recv), d) + send_off;
      rbar[d] = mapa_u32(smem_u32(recv_full), d);
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // cell-update role: cell c of the tile, batch rows rank*16 + bq*4 + {0..3}
    const int c = lane, bq = ew;
    const int cell = nt * 32 + c;
    const int b0 = static_cast<int>(rank) * 16 + bq * 4;
    const int H = p.H;
    int nf[4];
    bool valid[4];
    float cst[4], hst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      valid[i] = b0 + i < p.Bc;
      nf[i] = valid[i] ? __ldg(p.num_frames + b0 + i) : 0;
      cst[i] = 0.0f;
      hst[i] = 0.0f;
    }
    const uint32_t swz = static_cast<uint32_t>((c >> 1) & 3);
    const uint8_t* rd0 = recv + (bq * 32 + c) * 64;

    for (int t = 0; t < T; ++t) {
      // input projection of this step (independent of the recurrence: in flight while we wait)
      float4 xg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid[i])
          xg[i] = __ldg(reinterpret_cast<const float4*>(p.xw + (static_cast<long long>(b0 + i) * T + t) * 4 * H + nt * 128 + 4 * c));
      }
      float G[4][4];                             // [batch i][gate]
#pragma unroll
      for (int i = 0; i < 4; ++i) { G[i][0] = xg[i].x; G[i][1] = xg[i].y; G[i][2] = xg[i].z; G[i][3] = xg[i].w; }
      if (t > 0) {
        if (ew == 0 && lane == 0) mbar_arrive_expect_tx(recv_full, kRecvBytes);
        mbar_wait(acc_full, (t - 1) & 1);
        tc_fence_after();
        if (ew == 0 && lane == 0) LR_T(t, 2);
        float v[kNB];
        tmem_ld32(taddr, reinterpret_cast<uint32_t*>(v));
        tmem_ld32(taddr + 32, reinterpret_cast<uint32_t*>(v + 32));
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int d = 0; d < kKS; ++d)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_async_v4(rbase[d] + j * 2048, rbar[d], v[16 * d + 4 * j], v[16 * d + 4 * j + 1], v[16 * d + 4 * j + 2],
                        v[16 * d + 4 * j + 3]);
        if (ew == 0 && lane == 0) LR_T(t, 3);
        mbar_wait(recv_full, (t - 1) & 1);
        if (ew == 0 && lane == 0) LR_T(t, 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = (static_cast<uint32_t>(g) ^ swz) << 4;
#pragma unroll
          for (int s = 0; s < kKS; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(rd0 + s * kRecvSrcBytes + off);
            G[0][g] += a.x; G[1][g] += a.y; G[2][g] += a.z; G[3][g] += a.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!valid[i]) continue;
        const float c2 = cst[i] * sigmoidf_(G[i][2] + p.forget_bias) + sigmoidf_(G[i][0]) * tanhf_(G[i][1]);
        const float h2 = tanhf_(c2) * sigmoidf_(G[i][3]);
        const bool live = t < nf[i];
        if (live) { cst[i] = c2; hst[i] = h2; }
        const float ho = live ? h2 : 0.0f;
        __nv_bfloat16 hi, lo;
        split_bf16(ho, hi, lo);
        const long long o = (static_cast<long long>(b0 + i) * T + t) * H + cell;
        p.h_hi[o] = __bfloat16_as_ushort(hi);
        p.h_lo[o] = __bfloat16_as_ushort(lo);
        if (p.out_seq) p.out_seq[o] = ho;
      }
      if (ew == 0 && lane == 0) LR_T(t, 5);
      if (t + 1 < T) {
        fence_proxy_async_global();
        named_bar_sync(1, 128);                  // every thread's h_t stores (and recv reads) are done
        if (ew == 0 && lane == 0) {
          red_release_gpu_add(p.counter, 1u);
          LR_T(t, 6);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!valid[i]) continue;
      p.c_out[static_cast<long long>(b0 + i) * p.ld_state + cell] = cst[i];
      p.h_out[static_cast<long long>(b0 + i) * p.ld_state + cell] = hst[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody exits while a peer may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace

// all CTAs of a launch spin on each other, so they must be co-resident: ask the occupancy calculator once
bool yt8m::lstm_rec_available(int H) {
  if (!lstm_rec_supported(H)) return false;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal) != cudaSuccess) {
      (void)cudaGetLastError();
      max_clusters = 0;
    } else {
      cudaLaunchConfig_t qc{};
      qc.gridDim = dim3(kKS * 32, 1, 1);
      qc.blockDim = dim3(kThreads, 1, 1);
      qc.dynamicSmemBytes = kSmemTotal;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kKS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      qc.attrs = attr;
      qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_kernel, &qc) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
      max_clusters = n;
    }
  }
  return H / 32 <= max_clusters;
}
bool yt8m::lstm_rec_supported(int H) { return H % (64 * kKS) == 0 && H >= 64 * kKS && H <= 64 * kKS * kMaxKb; }
int yt8m::lstm_rec_batch_chunk() { return kNB; }

int yt8m::launch_lstm_rec(const float* xw, const int* num_frames, int B, int T, int H, const yt8m_bf16* w_rec, long long ldw,
                          float forget_bias, yt8m_bf16* h_hi, yt8m_bf16* h_lo, float* out_seq, float* c_out, float* h_out,
                          long long ld_state, unsigned int* counters, cudaStream_t stream) {
  YT8M_REQUIRE(lstm_rec_supported(H), YT8M_E_UNSUPPORTED, "lstm_rec: H=%d", H);
  const int nkb = H / (64 * kKS);
  const int clusters = H / 32;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_done = true;
  }
  YT8M_REQUIRE(lstm_rec_available(H), YT8M_E_UNSUPPORTED, "lstm_rec: %d clusters of %d CTAs cannot be co-resident", clusters, kKS);

  CUtensorMap tm_w;
  int rc;
  {
    // recurrent columns of the packed matrix viewed as [H/64][4H rows][64]
    const uint64_t dims[3] = {64, static_cast<uint64_t>(4) * H, static_cast<uint64_t>(H / 64)};
    const uint64_t strides[2] = {static_cast<uint64_t>(ldw) * 2, 128};
    const uint32_t box[3] = {64, 128, static_cast<uint32_t>(nkb)};
    if ((rc = make_tmap_bf16_nd(&tm_w, w_rec, 3, dims, strides, box)) != YT8M_OK) return rc;
  }
  const int n_chunks = (B + kNB - 1) / kNB;
  YT8M_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * n_chunks, stream));
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int b0 = ch * kNB;
    const int Bc = B - b0 < kNB ? B - b0 : kNB;
    const long long seq_off = static_cast<long long>(b0) * T * H;
    CUtensorMap tm_hhi, tm_hlo;
    {
      // h_seq chunk [Bc][T][H] viewed as [T][H/64][Bc][64]: one box = all batch rows x this CTA's k-blocks of one step
      const uint64_t dims[4] = {64, static_cast<uint64_t>(Bc), static_cast<uint64_t>(H / 64), static_cast<uint64_t>(T)};
      const uint64_t strides[3] = {static_cast<uint64_t>(T) * H * 2, 128, static_cast<uint64_t>(H) * 2};
      const uint32_t box[4] = {64, kNB, static_cast<uint32_t>(nkb), 1};
      if ((rc = make_tmap_bf16_nd(&tm_hhi, h_hi + seq_off, 4, dims, strides, box)) != YT8M_OK) return rc;
      if ((rc = make_tmap_bf16_nd(&tm_hlo, h_lo + seq_off, 4, dims, strides, box)) != YT8M_OK) return rc;
    }
    LstmRecParams p{};
    p.xw = xw + static_cast<long long>(b0) * T * 4 * H;
    p.num_frames = num_frames + b0;
    p.h_hi = reinterpret_cast<uint16_t*>(h_hi) + seq_off;
    p.h_lo = reinterpret_cast<uint16_t*>(h_lo) + seq_off;
    p.out_seq = out_seq ? out_seq + seq_off : nullptr;
    p.c_out = c_out + static_cast<long long>(b0) * ld_state;
    p.h_out = h_out + static_cast<long long>(b0) * ld_state;
    p.ld_state = ld_state;
    p.counter = counters + ch;
    p.timeline = ch == 0 ? host_debug_timeline() : nullptr;
    p.Bc = Bc; p.T = T; p.H = H;
    p.forget_bias = forget_bias;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kKS * clusters, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_rec_kernel, tm_w, tm_hhi, tm_hlo, p);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
      // The CTAs spin on a grid-wide barrier, so a launch WITHOUT the co-residency guarantee could deadlock a busy GPU: never
      // retry non-cooperatively.  The first chunk of the first layer reports "unsupported" and the caller runs the per-step
      // GEMM recurrence instead.
      (void)cudaGetLastError();
      set_error("lstm_rec_kernel: cooperative launch refused (%s)", cudaGetErrorString(e));
      return YT8M_E_UNSUPPORTED;
    }
    if (e != cudaSuccess) {
      set_error("lstm_rec_kernel: launch failed: %s", cudaGetErrorString(e));
      return YT8M_E_CUDA;
    }
    if ((rc = check_launch("lstm_rec_kernel")) != YT8M_OK) return rc;
  }
  return YT8M_OK;
}
