// yt8m_b200 -- C-ABI entry points built on the tcgen05 GEMM main loop: dense layer, MoE head, LSTM.
#include "yt8m_gemm.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {

constexpr int kNumSms = 148;

template <int BLOCK_N, int A_SPLIT, class Epi, int MT = 1, int CTAS = 1>
int launch_gemm(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* w, long long ldw, int M,
                int N_rows_w, int N, int K, int split_k, const typename Epi::Params& ep, cudaStream_t stream,
                int a_f16 = 0, unsigned long long hint_a = kEvictNormal, unsigned long long hint_w = kEvictNormal, bool pdl = false) {
  using S = GemmSmem<BLOCK_N, A_SPLIT, false, MT, CTAS>;
  CUtensorMap tm_a_hi, tm_a_lo, tm_b;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm_a_hi, a_hi, M, K, lda, kBlockM)) != YT8M_OK) return rc;
  if (A_SPLIT == 2) {
    if ((rc = make_tmap_bf16_2d(&tm_a_lo, a_lo, M, K, lda, kBlockM)) != YT8M_OK) return rc;
  } else {
    tm_a_lo = tm_a_hi;
  }
  if ((rc = make_tmap_bf16_2d(&tm_b, w, N_rows_w, K, ldw, BLOCK_N)) != YT8M_OK) return rc;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_SPLIT, Epi, false, MT, CTAS>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal + Epi::kSmemBytes));
    attr_done = true;
  }
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  GemmShape shape;
  shape.M = M; shape.N = N; shape.K = K; shape.a_f16 = a_f16;
  if (!(host_debug_flags() & 131072)) { shape.hint_a = hint_a; shape.hint_w = hint_w; }      // flag: A/B without the hints
  shape.timeline = host_debug_timeline();
  shape.kb_per_split = (num_kb + split_k - 1) / split_k;
  const int splits = (num_kb + shape.kb_per_split - 1) / shape.kb_per_split;
  const long long tiles = static_cast<long long>((N + BLOCK_N - 1) / BLOCK_N) * ((M + MT * kBlockM - 1) / (MT * kBlockM)) * splits;
  const int grid = static_cast<int>(std::min<long long>(tiles, kNumSms * CTAS));       // persistent CTAs walk the tiles
  if (pdl) {
    // the kernel's prologue overlaps the tail of the previous kernel on the stream (griddepcontrol.wait inside)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = S::kTotal + Epi::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    YT8M_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_a_hi, tm_a_lo, tm_b, shape, ep));
  } else {
    kern<<<grid, kGemmThreads, S::kTotal + Epi::kSmemBytes, stream>>>(tm_a_hi, tm_a_lo, tm_b, shape, ep);
  }
  return check_launch("gemm_tcgen05_kernel");
}

// ----- split-K finalize: ws fp32 [M, N] -> affine + activation -> outputs -------------------------
__global__ void linear_finalize_kernel(const float* __restrict__ ws, long long M, int N, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int act, float* out_f32,
                                       __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, long long ld_out, int out_f16) {
  const long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int c = static_cast<int>(i - r * N);
    float v = ws[i];
    if (scale) v *= scale[c];
    if (shift) v += shift[c];
    v = apply_act(v, act);
    const long long o = r * ld_out + c;
    if (out_f32) out_f32[o] = v;
    if (out_hi) {
      if (out_f16) {
        reinterpret_cast<__half*>(out_hi)[o] = __float2half_rn(v);
      } else {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        out_hi[o] = h;
        if (out_lo) out_lo[o] = l;
      }
    }
  }
}

// ----- weight packing ------------------------------------------------------------------------------
__global__ void pack_transpose_kernel(const float* __restrict__ w_kn, int K, int N, __nv_bfloat16* __restrict__ out,
                                      long long ldw) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w_kn[(long long)k * N + n] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < ldw) out[(long long)n * ldw + k] = __float2bfloat16_rn(k < K ? tile[threadIdx.x][i] : 0.0f);
  }
}

__global__ void moe_pack_kernel(const float* __restrict__ gate_w, const float* __restrict__ expert_w,
                                const float* __restrict__ expert_b, int D, int V, int M, __nv_bfloat16* __restrict__ wp,
                                long long ldw, float* __restrict__ bias_packed) {
  const int r = blockIdx.x;
  const int per = 2 * M + 1, cpt = 128 / per;
  const int tile = r / 128, j = r % 128;
  const int c = j / per, within = j % per;
  const int v = tile * cpt + c;
  const bool valid = (c < cpt) && (v < V);
  const float* src = nullptr;
  long long ncols = 0, col = 0;
  float bias = 0.0f;
  if (valid) {
    if (within <= M) { src = gate_w; ncols = (long long)V * (M + 1); col = (long long)v * (M + 1) + within; }
    else { src = expert_w; ncols = (long long)V * M; col = (long long)v * M + (within - M - 1); bias = expert_b[col]; }
  }
  for (int k = threadIdx.x; k < ldw; k += blockDim.x)
    wp[(long long)r * ldw + k] = __float2bfloat16_rn((valid && k < D) ? src[(long long)k * ncols + col] : 0.0f);
  if (threadIdx.x == 0) bias_packed[r] = bias;
}

__global__ void lstm_pack_kernel(const float* __restrict__ w_tf, const float* __restrict__ b_tf, int in_plus_h, int H,
                                 __nv_bfloat16* __restrict__ wp, float* __restrict__ bp) {
  const int r = blockIdx.x;               // packed row 4u+g
  const int u = r >> 2, g = r & 3;
  const long long col = (long long)g * H + u;
  for (int k = threadIdx.x; k < in_plus_h; k += blockDim.x)
    wp[(long long)r * in_plus_h + k] = __float2bfloat16_rn(w_tf[(long long)k * 4 * H + col]);
  if (threadIdx.x == 0) bp[r] = b_tf[col];
}

__global__ void group_max_kernel(const float* __restrict__ in, long long groups, int heads, int cols, float* __restrict__ out) {
  const long long total = groups * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / cols;
    const int c = static_cast<int>(i - g * cols);
    float m = in[(g * heads) * cols + c];
    for (int a = 1; a < heads; ++a) m = fmaxf(m, in[(g * heads + a) * cols + c]);
    out[i] = m;
  }
}

int pick_split_k(int M, int N, int K, int block_n, int mt) {
  const int tiles = ((M + mt * kBlockM - 1) / (mt * kBlockM)) * ((N + block_n - 1) / block_n);
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  if (tiles >= kNumSms / 2 || num_kb < 16) return 1;
  int s = std::min(kNumSms / tiles, num_kb / 8);        // never spill into a second wave
  return std::max(s, 1);
}

}  // namespace

extern "C" {

size_t yt8m_linear_workspace_bytes(int M, int N, int K) {
  (void)K;
  return static_cast<size_t>(M) * N * sizeof(float);
}

int yt8m_linear_fwd(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* w, long long ldw, int M,
                    int N, int K, const float* col_scale, const float* col_shift, int act, int a_fmt, int out_fmt,
                    float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, long long ld_out, void* workspace,
                    size_t workspace_bytes, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(a_hi && w, YT8M_E_BADPTR, "yt8m_linear_fwd: null operand");
  YT8M_REQUIRE(M > 0 && N > 0 && K > 0, YT8M_E_BADSHAPE, "yt8m_linear_fwd: bad shape M=%d N=%d K=%d", M, N, K);
  YT8M_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, YT8M_E_BADSHAPE,
               "yt8m_linear_fwd: lda=%lld ldw=%lld must be multiples of 8 and >= K=%d", lda, ldw, K);
  YT8M_REQUIRE(out_f32 || out_hi, YT8M_E_BADPTR, "yt8m_linear_fwd: no output");
  YT8M_REQUIRE(ld_out >= N, YT8M_E_BADSHAPE, "yt8m_linear_fwd: ld_out < N");
  YT8M_REQUIRE((a_fmt == YT8M_FMT_BF16 || a_fmt == YT8M_FMT_F16) && (out_fmt == YT8M_FMT_BF16 || out_fmt == YT8M_FMT_F16),
               YT8M_E_UNSUPPORTED, "yt8m_linear_fwd: operand format must be YT8M_FMT_BF16 or YT8M_FMT_F16");
  YT8M_REQUIRE(!(a_fmt == YT8M_FMT_F16 && a_lo) && !(out_fmt == YT8M_FMT_F16 && out_lo), YT8M_E_UNSUPPORTED,
               "yt8m_linear_fwd: an fp16 operand is a single tensor (no lo half)");
  const int a_f16 = a_fmt == YT8M_FMT_F16;
  int block_n = N <= 32 ? 32 : (N >= 512 && M > 128 ? 256 : 128);
  if ((host_debug_flags() & 16384) && block_n == 256) block_n = 128;                     // experiment: narrower tiles, deeper ring
  // two accumulators per CTA share every W tile -- unless that would leave only two pipeline stages (hi/lo A with
  // 256-wide tiles: measured slower, the ring is latency-bound) or the grid cannot fill the SMs anyway (then
  // split-K takes over and pairing M tiles would only double the number of fp32 atomics per output element)
  const long long tiles1 = static_cast<long long>((M + kBlockM - 1) / kBlockM) * ((N + block_n - 1) / block_n);
  int mt = (M > kBlockM && !(a_lo && block_n == 256) && tiles1 >= 2 * 148) ? 2 : 1;
  if ((host_debug_flags() & 256) && M > kBlockM && !(a_lo && block_n == 256)) mt = 2;      // experiment: force pairing
  if (host_debug_flags() & 512) mt = 1;
  int split_k = pick_split_k(M, N, K, block_n, mt);
  if (split_k > 1 && (!workspace || workspace_bytes < yt8m_linear_workspace_bytes(M, N, K))) split_k = 1;

  EpiLinear::Params ep;
  ep.col_scale = col_scale; ep.col_shift = col_shift; ep.act = act; ep.split_k = split_k;
  ep.out_f16 = out_fmt == YT8M_FMT_F16;
  if (split_k > 1) {
    YT8M_CUDA(cudaMemsetAsync(workspace, 0, static_cast<size_t>(M) * N * sizeof(float), stream));
    ep.out_f32 = static_cast<float*>(workspace); ep.out_hi = nullptr; ep.out_lo = nullptr; ep.ld_out = N;
  } else {
    ep.out_f32 = out_f32; ep.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
    ep.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo); ep.ld_out = ld_out;
  }
  int rc;
  // a weight matrix larger than half of L2 (126 MB) that every step streams once: do not let it displace what is re-read
  const unsigned long long hw = (static_cast<long long>(N) * K * 2 > (48ll << 20)) ? kEvictFirst : kEvictNormal;
  const unsigned long long ha = kEvictNormal;
#define YT8M_DISPATCH(BN)                                                                                              \
  rc = mt == 2 ? (a_lo ? launch_gemm<BN, 2, EpiLinear, 2>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, a_f16, ha, hw)    \
                       : launch_gemm<BN, 1, EpiLinear, 2>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, a_f16, ha, hw))   \
               : (a_lo ? launch_gemm<BN, 2, EpiLinear, 1>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, a_f16, ha, hw)    \
                       : launch_gemm<BN, 1, EpiLinear, 1>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, a_f16, ha, hw))
  if (block_n == 32) { YT8M_DISPATCH(32); }
  else if (block_n == 256) { YT8M_DISPATCH(256); }
  else { YT8M_DISPATCH(128); }
#undef YT8M_DISPATCH
  if (rc != YT8M_OK) return rc;
  if (split_k > 1) {
    const long long total = static_cast<long long>(M) * N;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, kNumSms * 8));
    linear_finalize_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), M, N, col_scale, col_shift,
                                                       act, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                       reinterpret_cast<__nv_bfloat16*>(out_lo), ld_out, ep.out_f16);
    return check_launch("linear_finalize_kernel");
  }
  return YT8M_OK;
}

}  // extern "C"

// acc[M, N] (fp32, row stride ld_acc) += A[M, K] . W[N, K]^T, the contraction split over CTAs and reduced with fp32 vector
// atomics straight into acc: no workspace, no memset, no finalize pass.  The reverse LSTM recurrence calls it once per frame
// (dh_{t-1} += dG_t . Wh: M = batch, N = H, K = 4H): with yt8m_linear_fwd's split-K that was a memset + GEMM + finalize
// triple per step (profiles/r02f_lstm_train_profile.txt: 598 x (1.7 + 7.9 + 2.4) us).
int yt8m::linear_accumulate(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* w, long long ldw, int M,
                            int N, int K, float* acc, long long ld_acc, cudaStream_t stream, bool pdl) {
  YT8M_REQUIRE(a_hi && w && acc, YT8M_E_BADPTR, "linear_accumulate: null pointer");
  YT8M_REQUIRE(M > 0 && N > 0 && K > 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K && ld_acc >= N, YT8M_E_BADSHAPE,
               "linear_accumulate: bad shape M=%d N=%d K=%d", M, N, K);
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + 127) / 128);
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int split_k = std::max(1, std::min(kNumSms / std::max(tiles, 1), num_kb / 4));
  EpiLinear::Params ep{};
  ep.act = ACT_NONE;
  ep.split_k = std::max(split_k, 2);                    // > 1 selects the atomic epilogue, also for a single split
  ep.out_f32 = acc;
  ep.ld_out = ld_acc;
  return a_lo ? launch_gemm<128, 2, EpiLinear, 1>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, 0, kEvictNormal, kEvictNormal, pdl)
              : launch_gemm<128, 1, EpiLinear, 1>(a_hi, a_lo, lda, w, ldw, M, N, N, K, split_k, ep, stream, 0, kEvictNormal, kEvictNormal, pdl);
}

extern "C" {

int yt8m_pack_transpose_bf16(const float* w_kn, int K, int N, yt8m_bf16* w_packed, long long ldw, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(w_kn && w_packed, YT8M_E_BADPTR, "yt8m_pack_transpose_bf16: null pointer");
  YT8M_REQUIRE(K > 0 && N > 0 && ldw >= K && ldw % 8 == 0, YT8M_E_BADSHAPE, "yt8m_pack_transpose_bf16: bad shape");
  dim3 grid((static_cast<int>(ldw) + 31) / 32, (N + 31) / 32);
  pack_transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(w_kn, K, N, reinterpret_cast<__nv_bfloat16*>(w_packed), ldw);
  return check_launch("pack_transpose_kernel");
}

long long yt8m_moe_packed_rows(int vocab, int num_mixtures) {
  if (vocab <= 0 || num_mixtures <= 0 || 2 * num_mixtures + 1 > 128) return -1;
  const int cpt = 128 / (2 * num_mixtures + 1);
  return 128LL * ((vocab + cpt - 1) / cpt);
}

int yt8m_moe_pack_weights(const float* gate_w, const float* expert_w, const float* expert_b, int D, int vocab,
                          int num_mixtures, yt8m_bf16* w_packed, long long ldw, float* bias_packed,
                          yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(gate_w && expert_w && expert_b && w_packed && bias_packed, YT8M_E_BADPTR, "yt8m_moe_pack_weights: null pointer");
  const long long rows = yt8m_moe_packed_rows(vocab, num_mixtures);
  YT8M_REQUIRE(rows > 0 && D > 0 && ldw >= D && ldw % 8 == 0, YT8M_E_BADSHAPE, "yt8m_moe_pack_weights: bad shape");
  moe_pack_kernel<<<static_cast<int>(rows), 256, 0, stream>>>(gate_w, expert_w, expert_b, D, vocab, num_mixtures,
                                                              reinterpret_cast<__nv_bfloat16*>(w_packed), ldw, bias_packed);
  return check_launch("moe_pack_kernel");
}

int yt8m_moe_fwd(const yt8m_bf16* x_hi, const yt8m_bf16* x_lo, long long ldx, const yt8m_bf16* w_packed, long long ldw,
                 const float* bias_packed, int B, int D, int vocab, int num_mixtures, int x_fmt, float* out,
                 long long ld_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x_hi && w_packed && bias_packed && out, YT8M_E_BADPTR, "yt8m_moe_fwd: null pointer");
  const long long rows = yt8m_moe_packed_rows(vocab, num_mixtures);
  YT8M_REQUIRE(rows > 0 && B > 0 && D > 0, YT8M_E_BADSHAPE, "yt8m_moe_fwd: bad shape B=%d D=%d V=%d M=%d", B, D, vocab,
               num_mixtures);
  YT8M_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= D && ldw >= D && ld_out >= vocab, YT8M_E_BADSHAPE,
               "yt8m_moe_fwd: strides must be multiples of 8 and >= D");
  YT8M_REQUIRE(x_fmt == YT8M_FMT_BF16 || (x_fmt == YT8M_FMT_F16 && !x_lo), YT8M_E_UNSUPPORTED,
               "yt8m_moe_fwd: x_fmt must be YT8M_FMT_BF16, or YT8M_FMT_F16 without a lo half");
  const int x_f16 = x_fmt == YT8M_FMT_F16;
  const int n = static_cast<int>(rows);
  // the packed head (50 MB for MoE-2 on 1024-d) fits L2 beside the streams of the other kernels: keep it for the next step
  const unsigned long long hw_moe = (rows * static_cast<long long>(D) * 2 <= (64ll << 20)) ? kEvictLast : kEvictNormal;
#define YT8M_MOE_CASE(NM)                                                                                   \
  case NM: {                                                                                                \
    EpiMoe<NM>::Params ep;                                                                                  \
    ep.out = out; ep.ld_out = ld_out; ep.bias_packed = bias_packed; ep.vocab = vocab;                      \
    if (B > 128 && (D >= 2048 || B > 256 || (host_debug_flags() & 32768)))   /* two accumulators per CTA pay off once the K loop is long */    \
      return x_lo ? launch_gemm<128, 2, EpiMoe<NM>, 2>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe) \
                  : launch_gemm<128, 1, EpiMoe<NM>, 2>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe); \
    if (D <= 2048 && !(host_debug_flags() & 1024))   /* short K loop: two co-resident CTAs per SM overlap epilogue and loads */ \
      return x_lo ? launch_gemm<128, 2, EpiMoe<NM>, 1, 2>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe) \
                  : launch_gemm<128, 1, EpiMoe<NM>, 1, 2>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe); \
    return x_lo ? launch_gemm<128, 2, EpiMoe<NM>, 1>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe) \
                : launch_gemm<128, 1, EpiMoe<NM>, 1>(x_hi, x_lo, ldx, w_packed, ldw, B, n, n, D, 1, ep, stream, x_f16, kEvictLast, hw_moe); \
  }
  switch (num_mixtures) {
    YT8M_MOE_CASE(1)
    YT8M_MOE_CASE(2)
    YT8M_MOE_CASE(3)
    YT8M_MOE_CASE(4)
    YT8M_MOE_CASE(8)
    default:
      set_error("yt8m_moe_fwd: num_mixtures=%d unsupported (1,2,3,4,8)", num_mixtures);
      return YT8M_E_UNSUPPORTED;
  }
#undef YT8M_MOE_CASE
}

int yt8m_group_max_rows(const float* in, long long groups, int heads, int cols, float* out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(in && out, YT8M_E_BADPTR, "yt8m_group_max_rows: null pointer");
  YT8M_REQUIRE(groups > 0 && heads > 0 && cols > 0, YT8M_E_BADSHAPE, "yt8m_group_max_rows: bad shape");
  const long long total = groups * cols;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, kNumSms * 8));
  group_max_kernel<<<blocks, 256, 0, stream>>>(in, groups, heads, cols, out);
  return check_launch("group_max_kernel");
}

// ------------------------------------------- LSTM -------------------------------------------------

int yt8m_lstm_pack_weights(const float* w_tf, const float* b_tf, int in_dim, int H, yt8m_bf16* w_packed, float* b_packed,
                           yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(w_tf && b_tf && w_packed && b_packed, YT8M_E_BADPTR, "yt8m_lstm_pack_weights: null pointer");
  YT8M_REQUIRE(in_dim > 0 && H > 0 && in_dim % 8 == 0 && H % 32 == 0, YT8M_E_BADSHAPE,
               "yt8m_lstm_pack_weights: in_dim %% 8 and H %% 32 must be 0");
  lstm_pack_kernel<<<4 * H, 256, 0, stream>>>(w_tf, b_tf, in_dim + H, H, reinterpret_cast<__nv_bfloat16*>(w_packed), b_packed);
  return check_launch("lstm_pack_kernel");
}

namespace {
struct LstmWs {
  float* xw;                     // [B*T, 4H]
  // persistent-recurrence path (yt8m_lstm_rec.cu): output sequences of two consecutive layers, barrier counters
  __nv_bfloat16* seq_hi[2];      // [B, T, H]
  __nv_bfloat16* seq_lo[2];
  unsigned int* counters;        // [L][ceil(B / chunk)]
  float* c[8][2];
  float* h[8][2];
  __nv_bfloat16* a_hi[8][2];     // layer 0: [B, H]; layer l>0: [B, 2H] = [h_{l-1} | h_l]
  __nv_bfloat16* a_lo[8][2];
  size_t total;
};
LstmWs carve_lstm_ws(void* base, int B, int T, int H, int L) {
  LstmWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  w.xw = static_cast<float*>(take(static_cast<size_t>(B) * T * 4 * H * sizeof(float)));
  if (lstm_rec_supported(H)) {
    for (int p = 0; p < 2; ++p) {
      w.seq_hi[p] = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * T * H * 2));
      w.seq_lo[p] = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * T * H * 2));
    }
    const int chunk = lstm_rec_batch_chunk();
    w.counters = static_cast<unsigned int*>(take(sizeof(unsigned int) * L * ((B + chunk - 1) / chunk)));
  }
  for (int l = 0; l < L; ++l)
    for (int p = 0; p < 2; ++p) {
      w.c[l][p] = static_cast<float*>(take(static_cast<size_t>(B) * H * sizeof(float)));
      w.h[l][p] = static_cast<float*>(take(static_cast<size_t>(B) * H * sizeof(float)));
      const size_t width = (l == 0) ? H : 2 * H;
      w.a_hi[l][p] = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * width * 2));
      w.a_lo[l][p] = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * width * 2));
    }
  w.total = off;
  return w;
}
struct LstmStatePtrs { const float* c[8]; const float* h[8]; };
__global__ void lstm_gather_state_kernel(const LstmStatePtrs sp, int B, int H, int L, float* out) {
  // out[b, l*2H + {0..H-1}] = c_l[b], out[b, l*2H + H + ...] = h_l[b]
  const long long total = (long long)B * L * 2 * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = static_cast<int>(i / (L * 2 * H));
    const int r = static_cast<int>(i % (L * 2 * H));
    const int l = r / (2 * H), which = (r % (2 * H)) / H, u = r % H;
    out[i] = (which == 0 ? sp.c[l] : sp.h[l])[(long long)b * H + u];
  }
}
}  // namespace

size_t yt8m_lstm_workspace_bytes(int B, int T, int D, int H, int L) {
  (void)D;
  if (B <= 0 || T <= 0 || H <= 0 || L <= 0 || L > 8) return 0;
  return carve_lstm_ws(nullptr, B, T, H, L).total;
}

// seq_hi_all / seq_lo_all (nullable): per-layer [B, T, H] buffers that RETAIN every layer's output sequence (training)
static int lstm_fwd_impl(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                         const yt8m_bf16* const* w_packed, const float* const* b_packed, float forget_bias, float* state_out,
                         float* out_seq, yt8m_bf16* out_seq_bf, yt8m_bf16* const* seq_hi_all, yt8m_bf16* const* seq_lo_all,
                         void* workspace, size_t workspace_bytes, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && num_frames && w_packed && b_packed && state_out && workspace, YT8M_E_BADPTR, "yt8m_lstm_fwd: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && D % 8 == 0 && H % 32 == 0 && L >= 1 && L <= 8, YT8M_E_BADSHAPE,
               "yt8m_lstm_fwd: bad shape B=%d T=%d D=%d H=%d L=%d", B, T, D, H, L);
  YT8M_REQUIRE(workspace_bytes >= yt8m_lstm_workspace_bytes(B, T, D, H, L), YT8M_E_BADSHAPE,
               "yt8m_lstm_fwd: workspace too small");
  LstmWs ws = carve_lstm_ws(workspace, B, T, H, L);
  bool rec_refused = false;
  if (!(host_debug_flags() & 8192) && lstm_rec_available(H)) {
    // one persistent launch per layer (and per 64 videos); the input projection of EVERY layer is one big GEMM
    const int chunk = lstm_rec_batch_chunk();
    const int n_chunks = (B + chunk - 1) / chunk;
    for (int l = 0; l < L && !rec_refused; ++l) {
      const bool top = (l == L - 1);
      int rc;
      if (l == 0)
        rc = yt8m_linear_fwd(x, nullptr, D, w_packed[0], D + H, B * T, 4 * H, D, nullptr, b_packed[0], YT8M_ACT_NONE,
                             YT8M_FMT_BF16, YT8M_FMT_BF16, ws.xw, nullptr, nullptr, 4 * H, nullptr, 0, stream_);
      else
        rc = yt8m_linear_fwd(seq_hi_all ? seq_hi_all[l - 1] : reinterpret_cast<const yt8m_bf16*>(ws.seq_hi[(l - 1) & 1]),
                             seq_lo_all ? seq_lo_all[l - 1] : reinterpret_cast<const yt8m_bf16*>(ws.seq_lo[(l - 1) & 1]), H,
                             w_packed[l], 2 * H, B * T, 4 * H, H, nullptr, b_packed[l], YT8M_ACT_NONE, YT8M_FMT_BF16,
                             YT8M_FMT_BF16, ws.xw, nullptr, nullptr, 4 * H, nullptr, 0, stream_);
      if (rc != YT8M_OK) return rc;
      yt8m_bf16* hi = seq_hi_all ? seq_hi_all[l] : reinterpret_cast<yt8m_bf16*>(ws.seq_hi[l & 1]);
      yt8m_bf16* lo = seq_lo_all ? seq_lo_all[l] : reinterpret_cast<yt8m_bf16*>(ws.seq_lo[l & 1]);
      if (top && out_seq_bf && !seq_hi_all) hi = out_seq_bf;   // the bf16 output sequence IS the hi half of h_seq
      const yt8m_bf16* w_rec = w_packed[l] + (l == 0 ? D : H);
      const long long ldw = (l == 0) ? (D + H) : 2 * H;
      rc = launch_lstm_rec(ws.xw, num_frames, B, T, H, w_rec, ldw, forget_bias, hi, lo,
                           top ? out_seq : nullptr, state_out + static_cast<long long>(l) * 2 * H,
                           state_out + static_cast<long long>(l) * 2 * H + H, static_cast<long long>(L) * 2 * H,
                           ws.counters + l * n_chunks, stream);
      if (rc == YT8M_E_UNSUPPORTED && l == 0 && !seq_hi_all) { rec_refused = true; break; }   // cooperative launch refused: per-step path
      if (rc != YT8M_OK) return rc;
    }
    if (!rec_refused) return YT8M_OK;
  }
  YT8M_REQUIRE(!seq_hi_all, YT8M_E_UNSUPPORTED,
               "yt8m_lstm_fwd_train: needs the persistent recurrence (H in {256, 512, 768, 1024} and an idle GPU), H=%d", H);
  // zero initial state (c, h fp32 and the bf16 operand buffers), both ping-pong halves
  YT8M_CUDA(cudaMemsetAsync(ws.c[0][0], 0, ws.total - (reinterpret_cast<char*>(ws.c[0][0]) - static_cast<char*>(workspace)), stream));

  // hoisted input projection of layer 0: xw = x . Wx0^T + b0  (packed column order), one big GEMM
  int rc = yt8m_linear_fwd(x, nullptr, D, w_packed[0], D + H, B * T, 4 * H, D, nullptr, b_packed[0], YT8M_ACT_NONE,
                           YT8M_FMT_BF16, YT8M_FMT_BF16, ws.xw, nullptr, nullptr, 4 * H, nullptr, 0, stream_);
  if (rc != YT8M_OK) return rc;

  for (int t = 0; t < T; ++t) {
    const int rd = t & 1, wr = rd ^ 1;
    for (int l = 0; l < L; ++l) {
      EpiLstm::Params ep{};
      ep.c_in = ws.c[l][rd]; ep.h_in = ws.h[l][rd]; ep.c_out = ws.c[l][wr]; ep.h_out = ws.h[l][wr];
      ep.num_frames = num_frames; ep.t = t; ep.hidden = H; ep.forget_bias = forget_bias;
      const bool top = (l == L - 1);
      if (top && out_seq) { ep.out_seq = out_seq + static_cast<long long>(t) * H; }
      if (top && out_seq_bf) { ep.out_seq_bf = reinterpret_cast<__nv_bfloat16*>(out_seq_bf) + static_cast<long long>(t) * H; }
      ep.ld_seq = static_cast<long long>(T) * H;
      // next-layer operand: left half of layer l+1's input buffer read at this same t
      if (!top) { ep.a1_hi = ws.a_hi[l + 1][rd]; ep.a1_lo = ws.a_lo[l + 1][rd]; ep.ld_a1 = 2 * H; }
      const yt8m_bf16 *a_hi, *a_lo;
      const yt8m_bf16* w;
      long long lda;
      int K;
      if (l == 0) {
        ep.xw = ws.xw + static_cast<long long>(t) * 4 * H; ep.ld_xw = static_cast<long long>(T) * 4 * H;
        ep.bias_packed = nullptr;
        ep.a0_hi = ws.a_hi[0][wr]; ep.a0_lo = ws.a_lo[0][wr]; ep.ld_a0 = H;
        a_hi = reinterpret_cast<const yt8m_bf16*>(ws.a_hi[0][rd]); a_lo = reinterpret_cast<const yt8m_bf16*>(ws.a_lo[0][rd]);
        lda = H; K = H; w = w_packed[0] + D;                       // recurrent columns of the packed matrix
      } else {
        ep.xw = nullptr; ep.bias_packed = b_packed[l];
        ep.a0_hi = ws.a_hi[l][wr] + H; ep.a0_lo = ws.a_lo[l][wr] + H; ep.ld_a0 = 2 * H;   // right half = own h
        a_hi = reinterpret_cast<const yt8m_bf16*>(ws.a_hi[l][rd]); a_lo = reinterpret_cast<const yt8m_bf16*>(ws.a_lo[l][rd]);
        lda = 2 * H; K = 2 * H; w = w_packed[l];
      }
      const long long ldw = (l == 0) ? (D + H) : 2 * H;
      rc = launch_gemm<128, 2, EpiLstm>(a_hi, a_lo, lda, w, ldw, B, 4 * H, 4 * H, K, 1, ep, stream);
      if (rc != YT8M_OK) return rc;
    }
  }
  // final state lives in the buffers written by the last step
  const int fin = T & 1;
  LstmStatePtrs sp{};
  for (int l = 0; l < L; ++l) { sp.c[l] = ws.c[l][fin]; sp.h[l] = ws.h[l][fin]; }
  const long long total = static_cast<long long>(B) * L * 2 * H;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, kNumSms * 4));
  lstm_gather_state_kernel<<<blocks, 256, 0, stream>>>(sp, B, H, L, state_out);
  return check_launch("lstm_gather_state_kernel");
}

int yt8m_lstm_fwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                  const yt8m_bf16* const* w_packed, const float* const* b_packed, float forget_bias, float* state_out,
                  float* out_seq, yt8m_bf16* out_seq_bf, void* workspace, size_t workspace_bytes, yt8m_stream_t stream_) {
  return lstm_fwd_impl(x, num_frames, B, T, D, H, L, w_packed, b_packed, forget_bias, state_out, out_seq, out_seq_bf, nullptr,
                       nullptr, workspace, workspace_bytes, stream_);
}

int yt8m_lstm_fwd_train(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                        const yt8m_bf16* const* w_packed, const float* const* b_packed, float forget_bias, float* state_out,
                        float* out_seq, yt8m_bf16* const* seq_hi, yt8m_bf16* const* seq_lo, void* workspace,
                        size_t workspace_bytes, yt8m_stream_t stream_) {
  YT8M_REQUIRE(seq_hi && seq_lo, YT8M_E_BADPTR, "yt8m_lstm_fwd_train: null sequence buffers");
  for (int l = 0; l < L && l < 8; ++l)
    YT8M_REQUIRE(seq_hi[l] && seq_lo[l] && aligned16(seq_hi[l]) && aligned16(seq_lo[l]), YT8M_E_BADPTR,
                 "yt8m_lstm_fwd_train: sequence buffer of layer %d is null or not 16-byte aligned", l);
  return lstm_fwd_impl(x, num_frames, B, T, D, H, L, w_packed, b_packed, forget_bias, state_out, out_seq, nullptr, seq_hi, seq_lo,
                       workspace, workspace_bytes, stream_);
}

}  // extern "C"
