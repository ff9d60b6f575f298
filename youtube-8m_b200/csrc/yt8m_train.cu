// yt8m_b200 -- training-step kernels: backward of the Logistic / MoE heads (wgrad as an MN-major tcgen05
// GEMM, dlogits from a fused backward epilogue), bias-gradient column sums, L2-regulariser + per-tensor
// clip + TF-1.0 Adam.  Reference semantics: wh/train.py:440-466 (reg, clip, Adam), wh/utils.py:164-174
// (clip_by_norm per tensor), wh/losses.py:114-130 (loss; its gradient comes from yt8m_xent_fwd_bwd).
//
// Master weights, gradients and Adam moments live in the PACKED (K-contiguous [out, in]) layout the
// forward kernels consume, so no transposes or re-packing sit on the step's critical path; the wgrad GEMM
// writes straight into that layout.
#include "yt8m_gemm.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {
constexpr int kNumSms = 148;

int grid_for(long long total, int per_block) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((total + per_block - 1) / per_block, kNumSms * 16)));
}

// dz = dp * p * (1 - p)  -> bf16 hi/lo  (sigmoid backward of LogisticModel, logistic_model.py:23-25)
__global__ void logistic_bwd_dz_kernel(const float* __restrict__ dp, const float* __restrict__ p, long long rows, int cols,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float pv = p[i];
    __nv_bfloat16 h, l;
    split_bf16(dp[i] * pv * (1.0f - pv), h, l);
    hi[r * ld + c] = h;
    lo[r * ld + c] = l;
  }
}

// column sums over the rows of a bf16 hi(+lo) matrix: out[n] = sum_b (hi + lo)[b, n].  A CTA owns 64 columns (32 lanes x 2
// columns, 8 row groups) of one row SLICE (blockIdx.y); with more than one slice the partial sums go to scratch and the last
// slice of a column block to finish adds them in slice order (a self-cleaning ticket: no atomics on the data, the result does
// not depend on arrival order).  The first version gave a column to a thread and walked ALL rows serially: 4 CTAs for the hidden
// layer's 1024 columns (40 us for 1 MB), 16 CTAs x 19,200 rows for the LSTM bias gradient.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long ld, int rows, int cols,
                   float* __restrict__ partial, unsigned int* __restrict__ tickets, float* __restrict__ out) {
  __shared__ float red[8][64];
  __shared__ unsigned int last;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * lane;
  const int slices = gridDim.y;
  const int per = (rows + slices - 1) / slices;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float a0 = 0.0f, a1 = 0.0f;
  if (c < cols) {                                        // cols and ld are even (bf16 pairs): c + 1 < cols too
#pragma unroll 4
    for (int r = r0 + grp; r < r1; r += 8) {
      const uint32_t h = __ldg(reinterpret_cast<const uint32_t*>(hi + static_cast<long long>(r) * ld + c));
      a0 += __uint_as_float(h << 16);
      a1 += __uint_as_float(h & 0xFFFF0000u);
      if (lo) {
        const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(lo + static_cast<long long>(r) * ld + c));
        a0 += __uint_as_float(l << 16);
        a1 += __uint_as_float(l & 0xFFFF0000u);
      }
    }
  }
  red[grp][2 * lane] = a0;
  red[grp][2 * lane + 1] = a1;
  __syncthreads();
  const int col = blockIdx.x * 64 + threadIdx.x;
  float t = 0.0f;
  if (threadIdx.x < 64) {
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
  }
  if (slices == 1) {
    if (threadIdx.x < 64 && col < cols) out[col] = t;
    return;
  }
  if (threadIdx.x < 64 && col < cols) partial[static_cast<long long>(blockIdx.y) * cols + col] = t;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + blockIdx.x, 1u);
  __syncthreads();
  if (last != static_cast<unsigned int>(slices - 1)) return;
  __threadfence();
  if (threadIdx.x < 64 && col < cols) {
    float s = 0.0f;
    for (int y = 0; y < slices; ++y) s += __ldcg(partial + static_cast<long long>(y) * cols + col);
    out[col] = s;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
}

// segment of a packed row: 0 = plain tensor / MoE gate rows, 1 = MoE expert rows, -1 = padding row
__device__ __forceinline__ int row_segment(long long row, int per, int nmix) {
  if (per <= 0) return 0;
  const int j = static_cast<int>(row % 128);
  const int cpt = 128 / per;
  if (j >= cpt * per) return -1;
  return ((j % per) <= nmix) ? 0 : 1;
}

// g += l2 * w (regulariser gradient, wh/train.py:440-442,459); per-segment sums of g^2 and w^2.
// sums[0..1] = sum g^2 (segment 0 / 1), sums[2..3] = sum w^2.
template <bool VEC>
__global__ void grad_reg_sumsq_kernel(float* __restrict__ g, const float* __restrict__ w, long long rows, int row_len, float l2,
                                      int per, int nmix, float* __restrict__ sums) {
  float sg[2] = {0.0f, 0.0f}, sw[2] = {0.0f, 0.0f};
  const long long total = rows * row_len;
  constexpr int kV = VEC ? 4 : 1;                   // VEC: row_len % 4 == 0 and 16-byte aligned pointers
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * kV; i < total; i += (long long)gridDim.x * blockDim.x * kV) {
    const int seg = per > 0 ? row_segment(i / row_len, per, nmix) : 0;     // one tensor (per == 0): no division
    if (seg < 0) continue;
    float wv[kV], gv[kV];
    if (VEC) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + i), g4 = *reinterpret_cast<const float4*>(g + i);
      wv[0] = w4.x; wv[kV > 1 ? 1 : 0] = w4.y; wv[kV > 2 ? 2 : 0] = w4.z; wv[kV > 3 ? 3 : 0] = w4.w;
      gv[0] = g4.x; gv[kV > 1 ? 1 : 0] = g4.y; gv[kV > 2 ? 2 : 0] = g4.z; gv[kV > 3 ? 3 : 0] = g4.w;
    } else {
      wv[0] = w[i];
      gv[0] = g[i];
    }
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      gv[j] += l2 * wv[j];
      sg[seg] += gv[j] * gv[j];
      sw[seg] += wv[j] * wv[j];
    }
    if (l2 != 0.0f) {
      if (VEC) *reinterpret_cast<float4*>(g + i) = make_float4(gv[0], gv[kV > 1 ? 1 : 0], gv[kV > 2 ? 2 : 0], gv[kV > 3 ? 3 : 0]);
      else g[i] = gv[0];
    }
  }
  __shared__ float red[4][8];
  __shared__ bool is_last;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float a = warp_sum(sg[s]), b = warp_sum(sw[s]);
    if ((threadIdx.x & 31) == 0) { red[s][threadIdx.x >> 5] = a; red[2 + s][threadIdx.x >> 5] = b; }
  }
  __syncthreads();
  // DETERMINISTIC across launches and across data-parallel ranks (replicas must apply bit-identical clip scales): every
  // block parks its four partial sums in sums[8 + 4 * block ..] and the LAST block to finish (ticket at sums[4]) adds them up
  // in block order -- no floating-point atomics.
  if (threadIdx.x < 4) {
    float t = 0.0f;
    for (int wq = 0; wq < (blockDim.x >> 5); ++wq) t += red[threadIdx.x][wq];
    sums[8 + 4 * blockIdx.x + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(reinterpret_cast<unsigned int*>(sums + 4), 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x < 128) {
    __threadfence();
    const int j = threadIdx.x & 3, part = threadIdx.x >> 2;          // 32 partial chains per sum, then a fixed tree
    float t = 0.0f;
    for (unsigned int b = part; b < gridDim.x; b += 32) t += __ldcg(sums + 8 + 4 * b + j);
    // lanes j, j+4, ..., j+28 of a warp hold 8 of the 32 chains of sum j: fixed butterfly, then the 4 warps through shared memory
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) < 4) red[threadIdx.x & 3][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 4) sums[threadIdx.x] = (red[threadIdx.x][0] + red[threadIdx.x][1]) + (red[threadIdx.x][2] + red[threadIdx.x][3]);
}

// per-tensor clip_by_norm (g * c / max(||g||, c)) + TF-1.0 Adam (epsilon outside the sqrt, lr_t carries the bias
// correction) + refresh of the bf16 operand copy.  only_seg >= 0 restricts the update to one segment.
template <bool VEC>
__global__ void clip_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 long long rows, int row_len, const float* __restrict__ sums, float clip, float lr_t, float beta1,
                                 float beta2, float eps, int per, int nmix, int only_seg, __nv_bfloat16* __restrict__ w_bf16) {
  float scale[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float n = sqrtf(sums[s]);
    scale[s] = clip > 0.0f ? clip / fmaxf(n, clip) : 1.0f;
  }
  const long long total = rows * row_len;
  constexpr int kV = VEC ? 4 : 1;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * kV; i < total; i += (long long)gridDim.x * blockDim.x * kV) {
    const int seg = per > 0 ? row_segment(i / row_len, per, nmix) : 0;
    if (seg < 0 || (only_seg >= 0 && seg != only_seg)) continue;
    float wv[4], gv[4], mv[4], vv[4];
    if (VEC) {
      const float4 a = *reinterpret_cast<const float4*>(w + i), b = *reinterpret_cast<const float4*>(g + i);
      const float4 c = *reinterpret_cast<const float4*>(m + i), d = *reinterpret_cast<const float4*>(v + i);
      wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w;
      gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
      mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w;
      vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
      wv[0] = w[i]; gv[0] = g[i]; mv[0] = m[i]; vv[0] = v[i];
    }
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      const float gs = gv[j] * scale[seg];
      mv[j] = beta1 * mv[j] + (1.0f - beta1) * gs;
      vv[j] = beta2 * vv[j] + (1.0f - beta2) * gs * gs;
      wv[j] -= lr_t * mv[j] / (sqrtf(vv[j]) + eps);
    }
    if (VEC) {
      *reinterpret_cast<float4*>(m + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      *reinterpret_cast<float4*>(w + i) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      if (w_bf16)
        *reinterpret_cast<uint2*>(w_bf16 + i) = make_uint2(pack_bf16x2(__float2bfloat16_rn(wv[0]), __float2bfloat16_rn(wv[1])),
                                                            pack_bf16x2(__float2bfloat16_rn(wv[2]), __float2bfloat16_rn(wv[3])));
    } else {
      m[i] = mv[0]; v[i] = vv[0]; w[i] = wv[0];
      if (w_bf16) w_bf16[i] = __float2bfloat16_rn(wv[0]);
    }
  }
}

}  // namespace

namespace yt8m {
// ---------------------------------------------------------------------------------------------
// MoE backward epilogue: recompute the logits tile, read dL/dp, write dL/dlogits (packed column order)
//   p = sum_m g_m s_m,  g = softmax_{M+1}(a),  s = sigmoid(e + bias)
//   dL/de_m = dp * g_m * s_m * (1 - s_m)
//   dL/da_j = dp * g_j * ((j < M ? s_j : 0) - p)
// ---------------------------------------------------------------------------------------------
template <int NMIX>
struct EpiMoeBwd {
  static constexpr int kSmemBytes = 0;
  template <class P> static __device__ __forceinline__ void stage(const P&, int, uint8_t*, int) {}
  static constexpr int kPer = 2 * NMIX + 1;
  static constexpr int kCpt = 128 / kPer;
  struct Params {
    const float* dp;            // [M, vocab]
    long long ld_dp;
    const float* bias_packed;
    __nv_bfloat16* dl_hi;       // [M, ld_dl] packed column order
    __nv_bfloat16* dl_lo;
    long long ld_dl;
    int vocab;
  };
  template <int BLOCK_N>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& /*s*/, int row, int n0, uint32_t taddr,
                                             bool row_valid, bool /*have_acc*/, int /*split*/, uint8_t* /*smem*/) {
    static_assert(BLOCK_N == 128, "MoE epilogue expects 128-column tiles");
    float acc[128];
#pragma unroll
    for (int c = 0; c < 128; c += 32) tmem_ld32(taddr + c, reinterpret_cast<uint32_t*>(acc) + c);
    tmem_ld_wait();
    if (!row_valid) return;
    const int v0 = (n0 / 128) * kCpt;
    const float* bias = p.bias_packed + n0;
    const float* dprow = p.dp + static_cast<long long>(row) * p.ld_dp + v0;
#pragma unroll
    for (int c = 0; c < kCpt; ++c) {
      const int o = c * kPer;
      float mx = acc[o];
#pragma unroll
      for (int m = 1; m <= NMIX; ++m) mx = fmaxf(mx, acc[o + m]);
      float g[NMIX + 1], sg[NMIX], den = 0.0f;
#pragma unroll
      for (int m = 0; m <= NMIX; ++m) { g[m] = __expf(acc[o + m] - mx); den += g[m]; }
      const float inv = 1.0f / den;
      float prob = 0.0f;
#pragma unroll
      for (int m = 0; m < NMIX; ++m) {
        sg[m] = sigmoidf_(acc[o + NMIX + 1 + m] + __ldg(bias + o + NMIX + 1 + m));
        g[m] *= inv;
        prob += g[m] * sg[m];
      }
      g[NMIX] *= inv;
      const float dpv = (v0 + c < p.vocab) ? __ldg(dprow + c) : 0.0f;
#pragma unroll
      for (int m = 0; m <= NMIX; ++m) acc[o + m] = dpv * g[m] * ((m < NMIX ? sg[m] : 0.0f) - prob);
#pragma unroll
      for (int m = 0; m < NMIX; ++m) acc[o + NMIX + 1 + m] = dpv * g[m] * sg[m] * (1.0f - sg[m]);
    }
#pragma unroll
    for (int c = kCpt * kPer; c < 128; ++c) acc[c] = 0.0f;          // padding columns
    __nv_bfloat16* oh = p.dl_hi + static_cast<long long>(row) * p.ld_dl + n0;
    __nv_bfloat16* ol = p.dl_lo + static_cast<long long>(row) * p.ld_dl + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint4 hi, lo;
      pack8_hi_lo(acc + 8 * j, hi, lo);
      reinterpret_cast<uint4*>(oh)[j] = hi;
      reinterpret_cast<uint4*>(ol)[j] = lo;
    }
  }
};
}  // namespace yt8m

namespace {
template <int BLOCK_N, int A_SPLIT, class Epi, bool MN>
int launch_gemm_t(const CUtensorMap& tm_a_hi, const CUtensorMap& tm_a_lo, const CUtensorMap& tm_b, int M, int N, int K,
                  const typename Epi::Params& ep, cudaStream_t stream, int split_k = 1) {
  using S = GemmSmem<BLOCK_N, A_SPLIT, MN>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_SPLIT, Epi, MN>;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal + Epi::kSmemBytes));
    attr_done = true;
  }
  constexpr int kStageK = MN ? 128 : kBlockK;
  GemmShape shape;
  shape.M = M; shape.N = N; shape.K = K; shape.a_f16 = 0;
  const int num_kb = (K + kStageK - 1) / kStageK;
  shape.kb_per_split = (num_kb + split_k - 1) / split_k;
  const int splits = (num_kb + shape.kb_per_split - 1) / shape.kb_per_split;
  const long long tiles = static_cast<long long>((N + BLOCK_N - 1) / BLOCK_N) * ((M + kBlockM - 1) / kBlockM) * splits;
  const int grid = static_cast<int>(std::min<long long>(tiles, 148));                  // persistent CTAs walk the tiles
  kern<<<grid, kGemmThreads, S::kTotal + Epi::kSmemBytes, stream>>>(tm_a_hi, tm_a_lo, tm_b, shape, ep);
  return check_launch("gemm_tcgen05_kernel");
}
}  // namespace

extern "C" {

int yt8m_logistic_bwd_dz(const float* dp, const float* p, int B, int V, yt8m_bf16* dz_hi, yt8m_bf16* dz_lo, long long ld,
                         yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(dp && p && dz_hi && dz_lo, YT8M_E_BADPTR, "yt8m_logistic_bwd_dz: null pointer");
  YT8M_REQUIRE(B > 0 && V > 0 && ld >= V, YT8M_E_BADSHAPE, "yt8m_logistic_bwd_dz: bad shape");
  logistic_bwd_dz_kernel<<<grid_for((long long)B * V, 256), 256, 0, stream>>>(dp, p, B, V, reinterpret_cast<__nv_bfloat16*>(dz_hi),
                                                                              reinterpret_cast<__nv_bfloat16*>(dz_lo), ld);
  return check_launch("logistic_bwd_dz_kernel");
}

// out[M, N] (fp32, row stride ld_out) = A^T . B with A = a_hi (+ a_lo) stored [Kb, lda >= M] and B stored [Kb, ldb >= N],
// i.e. the contraction runs over the ROWS (batch) of both stored matrices: the weight gradient dW^T[out, in] =
// dZ^T . X lands directly in the packed [out, in] layout.  M, N, lda, ldb multiples of 8.
int yt8m_wgrad(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* b, long long ldb, int M, int N,
               int Kb, float* out, long long ld_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(a_hi && b && out, YT8M_E_BADPTR, "yt8m_wgrad: null pointer");
  YT8M_REQUIRE(M > 0 && N > 0 && Kb > 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= M && ldb >= N && ld_out >= N,
               YT8M_E_BADSHAPE, "yt8m_wgrad: bad shape M=%d N=%d Kb=%d lda=%lld ldb=%lld", M, N, Kb, lda, ldb);
  CUtensorMap tm_a_hi, tm_a_lo, tm_b;
  int rc;
  // stored [Kb rows, M cols]: boxes of 128 rows x 64 columns
  if ((rc = make_tmap_bf16_2d(&tm_a_hi, a_hi, Kb, M, lda, 128)) != YT8M_OK) return rc;
  if (a_lo) { if ((rc = make_tmap_bf16_2d(&tm_a_lo, a_lo, Kb, M, lda, 128)) != YT8M_OK) return rc; }
  else tm_a_lo = tm_a_hi;
  if ((rc = make_tmap_bf16_2d(&tm_b, b, Kb, N, ldb, 128)) != YT8M_OK) return rc;
  // few output tiles but a long contraction (e.g. dCw^T[64, 1152] over B*T frame rows): split the contraction rows
  // over CTAs and reduce with fp32 vector atomics into the zeroed output
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128), num_kb = (Kb + 127) / 128;
  int split_k = 1;
  if (tiles < 74 && num_kb >= 8) split_k = std::max(1, std::min(148 / tiles, num_kb / 2));
  EpiLinear::Params ep{};
  ep.out_f32 = out; ep.ld_out = ld_out; ep.act = ACT_NONE; ep.split_k = split_k;
  if (split_k > 1)
    YT8M_CUDA(cudaMemset2DAsync(out, static_cast<size_t>(ld_out) * sizeof(float), 0, static_cast<size_t>(N) * sizeof(float), M, stream));
  return a_lo ? launch_gemm_t<128, 2, EpiLinear, true>(tm_a_hi, tm_a_lo, tm_b, M, N, Kb, ep, stream, split_k)
              : launch_gemm_t<128, 1, EpiLinear, true>(tm_a_hi, tm_a_lo, tm_b, M, N, Kb, ep, stream, split_k);
}

int yt8m_colsum_bf16(const yt8m_bf16* hi, const yt8m_bf16* lo, long long ld, int rows, int cols, float* out,
                     yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(hi && out, YT8M_E_BADPTR, "yt8m_colsum_bf16: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld >= cols && cols % 2 == 0 && ld % 2 == 0, YT8M_E_BADSHAPE,
               "yt8m_colsum_bf16: bad shape (cols and ld even) rows=%d cols=%d ld=%lld", rows, cols, ld);
  const int col_blocks = (cols + 63) / 64;
  const int slices = std::max(1, std::min(32, rows / 128));
  float* partial = nullptr;
  unsigned int* tickets = nullptr;
  if (slices > 1) {
    const size_t ticket_bytes = (static_cast<size_t>(col_blocks) * sizeof(unsigned int) + 255) & ~size_t(255);
    void* scratch = lib_scratch(kScratchColsum, 65536 + static_cast<size_t>(slices) * cols * sizeof(float), 65536, stream);
    YT8M_REQUIRE(ticket_bytes <= 65536, YT8M_E_UNSUPPORTED, "yt8m_colsum_bf16: too many columns (%d)", cols);
    if (!scratch) return YT8M_E_CUDA;
    tickets = static_cast<unsigned int*>(scratch);
    partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + 65536);
  }
  colsum_bf16_kernel<<<dim3(col_blocks, slices), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(hi),
                                                                  reinterpret_cast<const __nv_bfloat16*>(lo), ld, rows, cols, partial,
                                                                  tickets, out);
  return check_launch("colsum_bf16_kernel");
}

int yt8m_moe_bwd_dlogits(const yt8m_bf16* x_hi, const yt8m_bf16* x_lo, long long ldx, const yt8m_bf16* w_packed, long long ldw,
                         const float* bias_packed, const float* dp, long long ld_dp, int B, int D, int vocab, int num_mixtures,
                         yt8m_bf16* dl_hi, yt8m_bf16* dl_lo, long long ld_dl, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x_hi && w_packed && bias_packed && dp && dl_hi && dl_lo, YT8M_E_BADPTR, "yt8m_moe_bwd_dlogits: null pointer");
  const long long rows = yt8m_moe_packed_rows(vocab, num_mixtures);
  YT8M_REQUIRE(rows > 0 && B > 0 && D > 0 && ldx % 8 == 0 && ldw % 8 == 0 && ld_dl >= rows && ld_dl % 8 == 0 && ld_dp >= vocab,
               YT8M_E_BADSHAPE, "yt8m_moe_bwd_dlogits: bad shape");
  const int n = static_cast<int>(rows);
  CUtensorMap tm_a_hi, tm_a_lo, tm_b;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm_a_hi, x_hi, B, D, ldx, kBlockM)) != YT8M_OK) return rc;
  if (x_lo) { if ((rc = make_tmap_bf16_2d(&tm_a_lo, x_lo, B, D, ldx, kBlockM)) != YT8M_OK) return rc; }
  else tm_a_lo = tm_a_hi;
  if ((rc = make_tmap_bf16_2d(&tm_b, w_packed, n, D, ldw, 128)) != YT8M_OK) return rc;
#define YT8M_CASE(NM)                                                                                        \
  case NM: {                                                                                                  \
    EpiMoeBwd<NM>::Params ep;                                                                                 \
    ep.dp = dp; ep.ld_dp = ld_dp; ep.bias_packed = bias_packed; ep.vocab = vocab;                           \
    ep.dl_hi = reinterpret_cast<__nv_bfloat16*>(dl_hi); ep.dl_lo = reinterpret_cast<__nv_bfloat16*>(dl_lo); \
    ep.ld_dl = ld_dl;                                                                                         \
    return x_lo ? launch_gemm_t<128, 2, EpiMoeBwd<NM>, false>(tm_a_hi, tm_a_lo, tm_b, B, n, D, ep, stream)  \
                : launch_gemm_t<128, 1, EpiMoeBwd<NM>, false>(tm_a_hi, tm_a_lo, tm_b, B, n, D, ep, stream); \
  }
  switch (num_mixtures) {
    YT8M_CASE(1)
    YT8M_CASE(2)
    YT8M_CASE(4)
    default:
      set_error("yt8m_moe_bwd_dlogits: num_mixtures=%d unsupported (1, 2, 4)", num_mixtures);
      return YT8M_E_UNSUPPORTED;
  }
#undef YT8M_CASE
}

int yt8m_grad_reg_sumsq(float* grad, const float* param, long long rows, int row_len, float l2, int moe_per, int moe_nmix,
                        float* sums4, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(grad && param && sums4, YT8M_E_BADPTR, "yt8m_grad_reg_sumsq: null pointer");
  YT8M_REQUIRE(rows > 0 && row_len > 0, YT8M_E_BADSHAPE, "yt8m_grad_reg_sumsq: bad shape");
  YT8M_CUDA(cudaMemsetAsync(sums4, 0, 8 * sizeof(float), stream));       // the four sums + the block ticket
  const bool vec = row_len % 4 == 0 && aligned16(grad) && aligned16(param);
  // at most (YT8M_SUMS_FLOATS - 8) / 4 blocks: each parks four partials in the caller's buffer
  constexpr int kMaxBlocks = (YT8M_SUMS_FLOATS - 8) / 4;
  if (vec) grad_reg_sumsq_kernel<true><<<min(grid_for(rows * row_len / 4, 2048), kMaxBlocks), 256, 0, stream>>>(grad, param, rows, row_len, l2, moe_per, moe_nmix, sums4);
  else grad_reg_sumsq_kernel<false><<<min(grid_for(rows * row_len, 2048), kMaxBlocks), 256, 0, stream>>>(grad, param, rows, row_len, l2, moe_per, moe_nmix, sums4);
  return check_launch("grad_reg_sumsq_kernel");
}

int yt8m_clip_adam_step(float* param, const float* grad, float* m, float* v, long long rows, int row_len, const float* sums4,
                        float clip, float lr_t, float beta1, float beta2, float eps, int moe_per, int moe_nmix, int only_segment,
                        yt8m_bf16* param_bf16, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(param && grad && m && v && sums4, YT8M_E_BADPTR, "yt8m_clip_adam_step: null pointer");
  YT8M_REQUIRE(rows > 0 && row_len > 0, YT8M_E_BADSHAPE, "yt8m_clip_adam_step: bad shape");
  __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(param_bf16);
  const bool vec = row_len % 4 == 0 && aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v) &&
                   (!pb || (reinterpret_cast<uintptr_t>(pb) & 7u) == 0);
  if (vec)
    clip_adam_kernel<true><<<grid_for(rows * row_len / 4, 2048), 256, 0, stream>>>(param, grad, m, v, rows, row_len, sums4, clip, lr_t, beta1,
                                                                                   beta2, eps, moe_per, moe_nmix, only_segment, pb);
  else
    clip_adam_kernel<false><<<grid_for(rows * row_len, 2048), 256, 0, stream>>>(param, grad, m, v, rows, row_len, sums4, clip, lr_t, beta1,
                                                                                beta2, eps, moe_per, moe_nmix, only_segment, pb);
  return check_launch("clip_adam_kernel");
}

}  // extern "C"
