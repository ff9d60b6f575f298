// yt8m_b200 -- warp-specialised tcgen05 GEMM main loop with fused epilogues (sm_100a).
//
//   D[M, N] = epilogue( A[M, K] . W[N, K]^T )        A, W bf16 (K contiguous), fp32 accumulate in TMEM
//
// One CTA computes one 128 x BLOCK_N output tile (optionally one K-split of it):
//   warps 0, 6  : TMA producers (A tiles / W tiles; one thread issues a TMA op every ~140 ns whatever the
//                 box size -- measured, tools/micro/tma_ingest.cu -- so the two operands get a warp each)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16)
//   warps 2..5  : epilogue -- tcgen05.ld the accumulator row owned by each thread, apply the fused
//                 epilogue (bias/activation, MoE softmax x sigmoid, LSTM gates) and store.
// A may be given as a bf16 hi + lo pair (A_SPLIT = 2): both are multiplied against the same W tile
// and accumulated into the same TMEM tile, which carries ~16 mantissa bits of the activation through
// the tensor cores (used for recurrent state / chained activations; weights are bf16-exact).
#pragma once
#include "yt8m_common.cuh"

namespace yt8m {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 224;     // warp 0: A producer, 1: MMA, 2-5: epilogue, 6: W producer

struct GemmShape {
  int M, N, K;
  int kb_per_split;                    // K-blocks handled by one blockIdx.z
  int a_f16;                           // 1: A AND W hold IEEE fp16 (11 significant bits) instead of bf16 (fp16 x bf16 in one
                                       //    instruction is an illegal-instruction fault on sm_100a -- measured, tools/f16_diag.py)
  // L2 eviction priority of the TMA loads: a weight matrix that is streamed once per step and is larger than L2 must
  // not push out what the next kernel re-reads (evict-first); a head that fits (MoE: 50 MB) stays for the next step
  // (evict-last)
  unsigned long long hint_a = kEvictNormal;
  unsigned long long hint_w = kEvictNormal;
  // debug only (yt8m_debug_set_timeline; tools/gemm_timeline.py): globaltimer stamps of CTA 0 --
  // [0, 48) MMA warp: k-block c landed   [48, 64) epilogue: tile start / end   [64, 112) A producer: slot of k-block c free
  unsigned long long* timeline = nullptr;
};

// MN = false: operands stored [rows, K] with K contiguous (forward / dgrad): stage = 64 K-elements.
// MN = true : operands stored [K, rows] with the OUTPUT index contiguous (wgrad: out = A^T . B, contraction over
//             the batch rows): stage = 128 contraction rows, tiles are 64-column boxes of 128 rows.
// MT = number of 128-row M tiles one CTA owns (1 or 2): with MT = 2 every W tile is used for 256 output rows, i.e.
// 1.5x the tensor work per byte brought into shared memory -- the rings are bytes-in-flight / latency bound.
// CTAS = CTAs meant to be co-resident on one SM (1 or 2): with 2, each gets half the shared-memory budget and the
// epilogue of one overlaps the main loop of the other (short-K problems: the MoE head at small batch).
template <int BLOCK_N, int A_SPLIT, bool MN = false, int MT = 1, int CTAS = 1>
struct GemmSmem {
  static constexpr int kATile = MN ? 2 * 128 * 128 : kBlockM * kBlockK * 2;       // one 128-row A tile: 16 KB (K-major) / 32 KB (MN)
  static constexpr int kABytes = MT * kATile;                                      // all M tiles of one operand half (hi or lo)
  static constexpr int kBBytes = MN ? (BLOCK_N / 64) * 128 * 128 : BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = A_SPLIT * kABytes + kBBytes;
  static constexpr int kBudget = CTAS == 2 ? 96 * 1024 : 200 * 1024;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024 for alignment slack
  static_assert(kStages >= 2, "need at least two pipeline stages");
};

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// Epilogue concept:
//   struct Epi { struct Params {...};
//     template <int BLOCK_N> static __device__ void run(const Params&, const GemmShape&, int row (global M index),
//                                       int n0 (global N index of tile col 0), uint32_t tmem_row_addr, bool row_valid); }
// run() is called by every epilogue thread (uniformly per warp: tcgen05.ld is warp-collective).

template <int BLOCK_N, int A_SPLIT, class Epi, bool MN = false, int MT = 1, int CTAS = 1>
__global__ void __launch_bounds__(kGemmThreads, CTAS)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b, const GemmShape shape, const typename Epi::Params ep) {
  using S = GemmSmem<BLOCK_N, A_SPLIT, MN, MT, CTAS>;
  // PERSISTENT: the grid is min(#tiles, SMs x CTAS) and every CTA walks tiles blockIdx.x, +gridDim.x, ...  The
  // accumulator is double-buffered in TMEM when it fits (2 x MT x BLOCK_N columns), so the epilogue of tile i
  // overlaps the loads and MMAs of tile i+1, and the smem ring never drains between tiles.
  constexpr int kAccCols = MT * BLOCK_N;
  constexpr int kAccBufs = (CTAS * 2 * kAccCols <= 512) ? 2 : 1;
  static_assert(CTAS * kAccBufs * kAccCols <= 512, "accumulators exceed TMEM");
  constexpr uint32_t kTmemCols = tmem_cols_for(kAccBufs * kAccCols);
  constexpr int kStageK = MN ? 128 : kBlockK;          // contraction elements per pipeline stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + S::kStages * S::kStageBytes + S::kBarrierBytes;     // Epi::kSmemBytes, a quarter per epilogue warp

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int num_kb_total = (shape.K + kStageK - 1) / kStageK;
  const int n_tiles = (shape.N + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (shape.M + MT * kBlockM - 1) / (MT * kBlockM);
  const int splits = (num_kb_total + shape.kb_per_split - 1) / shape.kb_per_split;
  const int total_tiles = n_tiles * m_tiles * splits;

  // tile id -> (n_tile fastest, then m_tile, then K split): CTAs running side by side share the A rows
  struct Tile { int n_tile, m_tile, split, kb_begin, num_kb, kb_rot; };
  auto decode = [&](int t) {
    Tile r;
    r.n_tile = t % n_tiles;
    const int rest = t / n_tiles;
    r.m_tile = rest % m_tiles;
    r.split = rest / m_tiles;
    r.kb_begin = r.split * shape.kb_per_split;
    r.num_kb = min(r.kb_begin + shape.kb_per_split, num_kb_total) - r.kb_begin;     // >= 1 by construction of `splits`
    // CTAs that share an operand tile start their K loop at different k-blocks, so they do not all request the same
    // L2 lines at the same instant (no TMA multicast in this kernel)
    r.kb_rot = static_cast<int>((r.n_tile * 5u + r.m_tile * 3u) % static_cast<unsigned>(r.num_kb));
    return r;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    if (A_SPLIT == 2) tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 2);      // one arrive.expect_tx from each producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);                // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // programmatic dependent launch (used by the per-frame GEMMs of the reverse LSTM recurrence): everything above -- barrier
  // init, TMEM allocation, descriptor prefetch -- ran while the previous kernel was still executing; nothing below touches
  // global memory before the previous grid has completed.  Plain launches: both are no-ops.
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------- TMA producer: A (hi [+ lo]) -------------------------------
    int stage = 0, dbg_c = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      for (int kbi = 0; kbi < tl.num_kb; ++kbi) {
        const int kb = tl.kb_begin + (kbi + tl.kb_rot >= tl.num_kb ? kbi + tl.kb_rot - tl.num_kb : kbi + tl.kb_rot);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (shape.timeline && blockIdx.x == 0 && lane == 0 && dbg_c < 48) shape.timeline[64 + dbg_c] = global_timer_ns();
        ++dbg_c;
        if (elect_one()) {
          uint8_t* st = tiles + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], A_SPLIT * S::kABytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int row0 = (tl.m_tile * MT + mt) * kBlockM;
            uint8_t* at = st + mt * S::kATile;
            if (MN) {
              // two 64-column boxes of 128 contraction rows per operand half
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_2d(at + h * 16384, &tm_a_hi, &full_bar[stage], row0 + h * 64, kb * 128, shape.hint_a);
                if (A_SPLIT == 2)
                  tma_load_2d(at + S::kABytes + h * 16384, &tm_a_lo, &full_bar[stage], row0 + h * 64, kb * 128, shape.hint_a);
              }
            } else {
              tma_load_2d(at, &tm_a_hi, &full_bar[stage], kb * kBlockK, row0, shape.hint_a);
              if (A_SPLIT == 2) tma_load_2d(at + S::kABytes, &tm_a_lo, &full_bar[stage], kb * kBlockK, row0, shape.hint_a);
            }
          }
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 6) {
    // ------------------------------- TMA producer: W -------------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode(t);
      for (int kbi = 0; kbi < tl.num_kb; ++kbi) {
        const int kb = tl.kb_begin + (kbi + tl.kb_rot >= tl.num_kb ? kbi + tl.kb_rot - tl.num_kb : kbi + tl.kb_rot);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* st = tiles + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kBBytes);
          if (MN) {
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h)
              tma_load_2d(st + A_SPLIT * S::kABytes + h * 16384, &tm_b, &full_bar[stage], tl.n_tile * BLOCK_N + h * 64, kb * 128, shape.hint_w);
          } else {
            tma_load_2d(st + A_SPLIT * S::kABytes, &tm_b, &full_bar[stage], kb * kBlockK, tl.n_tile * BLOCK_N, shape.hint_w);
          }
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // a_format (bits 7..9) and b_format (bits 10..12): 1 = bf16, 0 = fp16; the hardware wants them equal
    const uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, MN ? 1 : 0, MN ? 1 : 0) ^ (shape.a_f16 ? ((1u << 7) | (1u << 10)) : 0u);
    int stage = 0, it = 0, dbg_c = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode(t);
      const int buf = it % kAccBufs, use = it / kAccBufs;
      if (use > 0) {                                   // the epilogue has drained this accumulator buffer
        mbar_wait(&tmem_empty_bar[buf], (use - 1) & 1);
        tc_fence_after();
      }
      for (int kb = 0; kb < tl.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (shape.timeline && blockIdx.x == 0 && lane == 0 && dbg_c < 48) shape.timeline[dbg_c] = global_timer_ns();
        ++dbg_c;
        if (elect_one()) {
          const uint32_t s_addr = smem_u32(tiles + stage * S::kStageBytes);
          const uint32_t b_addr = s_addr + A_SPLIT * S::kABytes;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t a_addr = s_addr + mt * S::kATile;
            const uint32_t d_tmem = tmem_base + buf * kAccCols + mt * BLOCK_N;
            if (MN) {
              // MN-major SWIZZLE_128B: 8-row groups 1024 B apart (SBO), 64-column blocks one box apart (LBO);
              // one UMMA consumes 16 contraction rows = 2048 B
              const uint64_t adesc0 = make_sdesc_sw128(a_addr, 16384, 1024);
              const uint64_t adesc1 = make_sdesc_sw128(a_addr + S::kABytes, 16384, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(b_addr, 16384, 1024);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t bdesc = sdesc_advance(bdesc0, k * 2048);
                umma_bf16(d_tmem, sdesc_advance(adesc0, k * 2048), bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                if (A_SPLIT == 2) umma_bf16(d_tmem, sdesc_advance(adesc1, k * 2048), bdesc, idesc, 1u);
              }
            } else {
              const uint64_t adesc0 = make_sdesc_sw128(a_addr, 16, 1024);
              const uint64_t adesc1 = make_sdesc_sw128(a_addr + S::kABytes, 16, 1024);
              const uint64_t bdesc0 = make_sdesc_sw128(b_addr, 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t bdesc = sdesc_advance(bdesc0, k * (kUmmaK * 2));
                umma_bf16(d_tmem, sdesc_advance(adesc0, k * (kUmmaK * 2)), bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                if (A_SPLIT == 2) umma_bf16(d_tmem, sdesc_advance(adesc1, k * (kUmmaK * 2)), bdesc, idesc, 1u);
              }
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == tl.num_kb - 1) umma_commit(&tmem_full_bar[buf]);
        }
        __syncwarp();
        if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    const int row_in_tile = q * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode(t);
      const int buf = it % kAccBufs, use = it / kAccBufs;
      // per-tile operands of the epilogue (e.g. the MoE expert biases) are staged in this warp's shared-memory quarter
      // while the main loop of the tile is still running
      Epi::stage(ep, tl.n_tile * BLOCK_N, epi_smem + q * (Epi::kSmemBytes / 4), lane);
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      if (shape.timeline && blockIdx.x == 0 && threadIdx.x == 64 && it < 8) shape.timeline[48 + 2 * it] = global_timer_ns();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int row = (tl.m_tile * MT + mt) * kBlockM + row_in_tile;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccCols + mt * BLOCK_N;
        Epi::template run<BLOCK_N>(ep, shape, row, tl.n_tile * BLOCK_N, taddr, row < shape.M, true, tl.split,
                                   epi_smem + q * (Epi::kSmemBytes / 4));
      }
      tc_fence_before();
      __syncwarp();
      if (shape.timeline && blockIdx.x == 0 && threadIdx.x == 64 && it < 8) shape.timeline[48 + 2 * it + 1] = global_timer_ns();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Epilogue 1: linear  -- y = act((acc) * col_scale + col_shift)  [col_shift doubles as the bias]
//   outputs: fp32 and/or bf16 hi (+ lo).  With split-K > 1 the raw accumulator is atomically added
//   into a zero-initialised fp32 workspace and a finalize kernel applies the affine/activation.
// ---------------------------------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU6 = 2, ACT_SIGMOID = 3, ACT_TANH = 4 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_RELU6: return fminf(fmaxf(v, 0.0f), 6.0f);
    case ACT_SIGMOID: return sigmoidf_(v);
    case ACT_TANH: return tanhf_(v);
    default: return v;
  }
}

struct EpiLinear {
  template <class P> static __device__ __forceinline__ void stage(const P&, int, uint8_t*, int) {}
  // (a per-warp transpose through shared memory for 128-byte row stores was measured SLOWER: 8192^3 GEMM 1082 -> 894
  //  TFLOP/s, wgrad +14% -- the strided 16-byte stores are not what bounds these epilogues)
  static constexpr int kSmemBytes = 0;
  struct Params {
    float* out_f32;            // nullable
    __nv_bfloat16* out_hi;     // nullable
    __nv_bfloat16* out_lo;     // nullable
    long long ld_out;          // row stride (elements) of all outputs
    const float* col_scale;    // nullable (=> 1)
    const float* col_shift;    // nullable (=> 0): bias / folded BN shift
    int act;
    int split_k;               // > 1: atomicAdd raw partials into out_f32, nothing else
    int out_f16;               // 1: out_hi receives fp16 (one 11-bit operand for the next GEMM), out_lo unused
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& s, int row, int n0, uint32_t taddr,
                                             bool row_valid, bool have_acc, int /*split*/, uint8_t* /*smem*/) {
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t r[32];
      if (have_acc) {
        tmem_ld32(taddr + c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (!row_valid) continue;
      const int col0 = n0 + c;
      if (col0 >= s.N) continue;
      const long long base = static_cast<long long>(row) * p.ld_out + col0;
      if (p.split_k > 1) {
        if (col0 + 32 <= s.N && (p.ld_out & 3) == 0) {
          // 16-byte vector reductions: a quarter of the L2 atomic operations of the scalar form
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(p.out_f32 + base + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                       __uint_as_float(r[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < s.N) atomicAdd(p.out_f32 + base + j, __uint_as_float(r[j]));
        }
        continue;
      }
      // The per-column affine and the activation are applied chunk-wise with the run-time tests OUTSIDE the element loops: the
      // first version tested col < N, col_scale, col_shift and switched on the activation for each of the 32 elements (~70 SASS
      // instructions per element): the weight-gradient GEMMs (302 MB of fp32 output, no affine, no activation) were bound by
      // the epilogue's INSTRUCTION count at 615 GB/s, whatever the contraction length (profiles/r02d_wgrad_ncu.md).
      const bool full = (col0 + 32 <= s.N) && ((p.ld_out & 7) == 0);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (col0 + 32 <= s.N) {
        if (p.col_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= __ldg(p.col_scale + col0 + j);
        }
        if (p.col_shift) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(p.col_shift + col0 + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (col0 + j < s.N) {
            if (p.col_scale) v[j] *= __ldg(p.col_scale + col0 + j);
            if (p.col_shift) v[j] += __ldg(p.col_shift + col0 + j);
          }
        }
      }
      switch (p.act) {
        case ACT_NONE: break;
        case ACT_RELU:
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          break;
        case ACT_RELU6:
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], 0.0f), 6.0f);
          break;
        case ACT_SIGMOID:
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j]);
          break;
        default:
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          break;
      }
      if (p.out_f32) {
        if (full) {
          float4* dst = reinterpret_cast<float4*>(p.out_f32 + base);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < s.N) p.out_f32[base + j] = v[j];
        }
      }
      if (p.out_hi) {
        if (full) {
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + base);
          uint4* dl = p.out_lo ? reinterpret_cast<uint4*>(p.out_lo + base) : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h, l;
            if (p.out_f16) {
              h = pack8_f16(v + 8 * j);
            } else {
              pack8_hi_lo(v + 8 * j, h, l);
              if (dl) dl[j] = l;
            }
            dh[j] = h;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < s.N) {
              if (p.out_f16) {
                reinterpret_cast<__half*>(p.out_hi)[base + j] = __float2half_rn(v[j]);
              } else {
                __nv_bfloat16 h, l;
                split_bf16(v[j], h, l);
                p.out_hi[base + j] = h;
                if (p.out_lo) p.out_lo[base + j] = l;
              }
            }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Epilogue 2: MoE head (wh/all_video_models/moe_model.py:54-64 of the reference).
//   Packed weight rows: tile of 128 rows = CPT = floor(128/(2M+1)) classes, class-major, each class
//   = [gate_0..gate_M, expert_0..expert_{M-1}], zero padded to 128.  The epilogue thread owning batch
//   row b turns its 128 accumulators into CPT probabilities:
//       p[b, v] = sum_{m<M} softmax_{M+1}(gate)[m] * sigmoid(expert[m] + bias[m])
//   The B x V(2M+1) activations never reach HBM.  Optional max over `heads` consecutive rows
//   (MoeExtendModel / attention-max-pooling) is done by the caller's reduce kernel.
// ---------------------------------------------------------------------------------------------
template <int NMIX>
struct EpiMoe {
  // the 128 packed biases of the tile, one copy per epilogue warp: loading them from global memory right where they are
  // used put an exposed L2 round trip in front of every sigmoid (measured: 12-15 us per tile, tools/gemm_timeline.py)
  static constexpr int kSmemBytes = 4 * 512;
  static constexpr int kPer = 2 * NMIX + 1;
  static constexpr int kCpt = 128 / kPer;
  struct Params {
    float* out;                 // [M, vocab] fp32
    long long ld_out;
    const float* bias_packed;   // [n_tiles * 128]: expert biases in packed order (0 for gates / padding)
    int vocab;
  };
  static __device__ __forceinline__ void stage(const Params& p, int n0, uint8_t* smem, int lane) {
    __syncwarp();                                    // the previous tile's reads of this quarter are done
    reinterpret_cast<float4*>(smem)[lane] = __ldg(reinterpret_cast<const float4*>(p.bias_packed + n0) + lane);
    __syncwarp();
  }
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& /*s*/, int row, int n0, uint32_t taddr,
                                             bool row_valid, bool /*have_acc*/, int /*split*/, uint8_t* smem) {
    static_assert(BLOCK_N == 128, "MoE epilogue expects 128-column tiles");
    // 32 accumulator columns (= kCpc whole classes) at a time: with all 128 live the compiler had no registers left to
    // overlap the exp / reciprocal chains of neighbouring classes (measured 8-15 us per tile)
    constexpr int kCpc = 32 / kPer;                         // classes per chunk
    constexpr int kChunks = (kCpt + kCpc - 1) / kCpc;
    const int v0 = (n0 / 128) * kCpt;
    const float* bias = reinterpret_cast<const float*>(smem);
    float* out = p.out + static_cast<long long>(row) * p.ld_out + v0;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      const int c_first = ch * kCpc;
      const int col_want = c_first * kPer;
      const int col0 = col_want + 32 <= 128 ? col_want : 128 - 32;      // the last chunk is shifted back into the tile
      const int loc = col_want - col0;
      float a[32];
      tmem_ld32(taddr + col0, reinterpret_cast<uint32_t*>(a));
      tmem_ld_wait();
      if (row_valid) {
#pragma unroll
        for (int cc = 0; cc < kCpc; ++cc) {
          const int c = c_first + cc;
          if (c < kCpt) {
            const int o = loc + cc * kPer;
            const int ob = c * kPer;
            float mx = a[o];
#pragma unroll
            for (int m = 1; m <= NMIX; ++m) mx = fmaxf(mx, a[o + m]);
            float den = 0.0f, num = 0.0f;
#pragma unroll
            for (int m = 0; m <= NMIX; ++m) {
              const float e = __expf(a[o + m] - mx);
              den += e;
              if (m < NMIX) num += e * sigmoidf_(a[o + NMIX + 1 + m] + bias[ob + NMIX + 1 + m]);
            }
            if (v0 + c < p.vocab) out[c] = num / den;
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Epilogue 3: one BasicLSTMCell step (TF 1.0 semantics; call site wh/all_frame_models/lstm_model.py:34-47).
//   Packed weight rows: unit-major, [i_u, j_u, f_u, o_u] per hidden unit u (128-row tile = 32 units).
//   acc = [x_t, h_{t-1}] . W   (or only the h part when the x projection was hoisted: then xw holds
//   x_t . W_x + b in the same packed column order).
//     c' = c * sigmoid(f + forget_bias) + sigmoid(i) * tanh(j);  h' = tanh(c') * sigmoid(o)
//   dynamic_rnn(sequence_length): rows with t >= num_frames[b] keep (c, h) and emit 0.
// ---------------------------------------------------------------------------------------------
struct EpiLstm {
  static constexpr int kSmemBytes = 0;
  template <class P> static __device__ __forceinline__ void stage(const P&, int, uint8_t*, int) {}
  struct Params {
    const float* xw;            // nullable: [B, 4H] slice for this t (row stride ld_xw), packed order
    long long ld_xw;
    const float* bias_packed;   // nullable: [4H] packed order
    const float* c_in;          // [B, H]
    const float* h_in;          // [B, H] fp32
    float* c_out;
    float* h_out;
    __nv_bfloat16* a0_hi;       // destination 0 for bf16 hi/lo of h' (own recurrent operand), row stride ld_a0
    __nv_bfloat16* a0_lo;
    long long ld_a0;
    __nv_bfloat16* a1_hi;       // nullable destination 1 (next layer's input operand)
    __nv_bfloat16* a1_lo;
    long long ld_a1;
    float* out_seq;             // nullable: [B, T, H] fp32 outputs (pre-offset to time t), row stride ld_seq
    __nv_bfloat16* out_seq_bf;  // nullable: bf16 copy of the outputs (operand for attention logits)
    long long ld_seq;
    const int* num_frames;      // [B]
    int t;
    int hidden;
    float forget_bias;
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& /*s*/, int row, int n0, uint32_t taddr,
                                             bool row_valid, bool /*have_acc*/, int /*split*/, uint8_t* /*smem*/) {
    static_assert(BLOCK_N == 128, "LSTM epilogue expects 128-column tiles (32 units)");
    const int u0 = n0 / 4;
    const bool live = row_valid && (p.t < __ldg(p.num_frames + (row_valid ? row : 0)));
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c, r);
      tmem_ld_wait();
      if (!row_valid) continue;
      float g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(r[j]);
      if (p.xw) {
        const float4* xw = reinterpret_cast<const float4*>(p.xw + static_cast<long long>(row) * p.ld_xw + n0 + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = __ldg(xw + j);
          g[4 * j] += x.x; g[4 * j + 1] += x.y; g[4 * j + 2] += x.z; g[4 * j + 3] += x.w;
        }
      }
      if (p.bias_packed) {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] += __ldg(p.bias_packed + n0 + c + j);
      }
      const int ub = u0 + c / 4;                          // 8 units in this chunk
      const long long sbase = static_cast<long long>(row) * p.hidden + ub;
      float cn[8], hn[8], ho[8];
      {
        const float4* ci = reinterpret_cast<const float4*>(p.c_in + sbase);
        const float4* hi4 = reinterpret_cast<const float4*>(p.h_in + sbase);
        const float4 c0 = ci[0], c1 = ci[1], h0 = hi4[0], h1 = hi4[1];
        const float cp[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float hp[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float gi = g[4 * u], gj = g[4 * u + 1], gf = g[4 * u + 2], go = g[4 * u + 3];
          const float c2 = cp[u] * sigmoidf_(gf + p.forget_bias) + sigmoidf_(gi) * tanhf_(gj);
          const float h2 = tanhf_(c2) * sigmoidf_(go);
          cn[u] = live ? c2 : cp[u];
          hn[u] = live ? h2 : hp[u];
          ho[u] = live ? h2 : 0.0f;
        }
      }
      float4* co = reinterpret_cast<float4*>(p.c_out + sbase);
      float4* hofp = reinterpret_cast<float4*>(p.h_out + sbase);
      co[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
      co[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
      hofp[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
      hofp[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
      uint4 hi, lo;
      pack8_hi_lo(hn, hi, lo);
      *reinterpret_cast<uint4*>(p.a0_hi + static_cast<long long>(row) * p.ld_a0 + ub) = hi;
      *reinterpret_cast<uint4*>(p.a0_lo + static_cast<long long>(row) * p.ld_a0 + ub) = lo;
      if (p.a1_hi) {
        // the layer above consumes this step's h' (finished rows are frozen there too, so what they
        // receive is irrelevant)
        *reinterpret_cast<uint4*>(p.a1_hi + static_cast<long long>(row) * p.ld_a1 + ub) = hi;
        *reinterpret_cast<uint4*>(p.a1_lo + static_cast<long long>(row) * p.ld_a1 + ub) = lo;
      }
      if (p.out_seq) {
        float4* os = reinterpret_cast<float4*>(p.out_seq + static_cast<long long>(row) * p.ld_seq + ub);
        os[0] = make_float4(ho[0], ho[1], ho[2], ho[3]);
        os[1] = make_float4(ho[4], ho[5], ho[6], ho[7]);
      }
      if (p.out_seq_bf) {
        uint4 oh, ol;
        pack8_hi_lo(ho, oh, ol);
        *reinterpret_cast<uint4*>(p.out_seq_bf + static_cast<long long>(row) * p.ld_seq + ub) = oh;
      }
    }
  }
};

}  // namespace yt8m
