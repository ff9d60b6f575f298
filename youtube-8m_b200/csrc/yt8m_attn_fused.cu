// yt8m_b200 -- fused multi-head attention pooling over the raw frames (zt/frame_level_models.py:4372-4398; BASELINE.json configs[4]):
//   logits[t, a] = x[t, :] . W[:, a]   (the mean-pooled half of W and the bias add a per-video constant: softmax over t ignores it)
//   w[t, a] = softmax_t(logits)[t, a] * mask[t], renormalised over t        (mask: t < num_frames, or "frame row is non-zero")
//   out[a, :] = sum_t w[t, a] * x[t, :]
// ONE kernel, one CTA per video, the frames leave HBM once: phase A (a warp per frame) computes the A logits of every frame and
// keeps them in shared memory (T x A floats); phase B streams the same frames again -- out of L2, they were read microseconds
// ago by the same SM -- and accumulates the A weighted sums, every thread owning four feature columns.  SURVEY.md §8(d): HBM
// bound, 691,200 B per 300-frame video.  The predecessor ran a tensor-core GEMM over all B*T rows (N = 8 of a 32-wide tile),
// wrote B*T x 8 logits to HBM and read the frames a second time in a separate kernel: 0.50 ms at B = 256 (profiles/
// r02a_bench_c5.json); the arithmetic here is 11 MFLOP per video -- CUDA cores are plenty.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {

constexpr int kMaxA = 8;

template <int A>
__global__ void __launch_bounds__(1024, 1)
attn_pool_fused_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, long long ldw,
                       const int* __restrict__ num_frames, int T, int D, int G, float* __restrict__ out,
                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* wl = reinterpret_cast<float*>(smem_raw);               // [T][A] logits -> weights
  float* inv_s = wl + static_cast<size_t>(T) * A;               // [A] 1 / sum (+ A floats of padding: 16-byte alignment below)
  // W^T staged as [D][A] bf16: one 16-byte vector per feature (A = 8)
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(inv_s + 2 * A);
  static_assert(A == 8, "one uint4 of weights per feature");
  uint8_t* fold_raw = reinterpret_cast<uint8_t*>(wt + static_cast<size_t>(D) * A);     // [A][D] floats: phase B's group fold (16-byte aligned: D % 8 == 0)
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const __nv_bfloat16* xb = x + static_cast<long long>(b) * T * D;
  const int nf = num_frames ? min(max(__ldg(num_frames + b), 0), T) : T;
  for (int i = tid; i < D * A; i += blockDim.x) {
    const int d = i / A, a = i - d * A;
    // K-major packed [A, ldw] -> [D][A]; the 16-byte slot of feature d inside its 128-byte chunk (8 features) is XOR-swizzled
    // with the chunk index: consecutive lanes read consecutive chunks in phase A, without bank conflicts
    const int c = d >> 3, slot = (d & 7) ^ (c & 7);
    wt[(c * 8 + slot) * A + a] = w[static_cast<long long>(a) * ldw + d];
  }
  __syncthreads();
  // ---- phase A: a warp per frame; lane owns the 16-byte chunks (8 features) lane, lane + 32, ...; the chunks of a frame are
  //      requested in batches of kBatch BEFORE any of them is used: the pass is latency bound (one warp, one frame at a time),
  //      so the number of loads in flight is what sets its speed (a load-use-load loop ran at ~12 us per frame) ----
  constexpr int kBatch = 5;                                      // 5 x 32 lanes x 8 features = 1280 >= 1152
  const int nchunks = D / 8;
  for (int t = warp; t < nf; t += nwarps) {
    const uint4* row = reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * D);
    float acc[A];
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a] = 0.0f;
    bool nz = false;
    for (int c0 = 0; c0 < nchunks; c0 += 32 * kBatch) {
      uint4 u[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int c = c0 + lane + 32 * j;
        u[j] = c < nchunks ? __ldg(row + c) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int c = c0 + lane + 32 * j;
        if (c < nchunks) {
          const uint32_t xw[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
          nz |= ((xw[0] | xw[1] | xw[2] | xw[3]) & 0x7FFF7FFFu) != 0u;
          const uint4* wrow = reinterpret_cast<const uint4*>(wt + static_cast<size_t>(c) * 8 * A);     // 8 features x 8 heads
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x0 = __uint_as_float(xw[i] << 16), x1 = __uint_as_float(xw[i] & 0xFFFF0000u);
            const uint4 wa = wrow[(2 * i) ^ (c & 7)], wb = wrow[(2 * i + 1) ^ (c & 7)];
            const uint32_t wa4[4] = {wa.x, wa.y, wa.z, wa.w}, wb4[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              acc[2 * k] += x0 * __uint_as_float(wa4[k] << 16) + x1 * __uint_as_float(wb4[k] << 16);
              acc[2 * k + 1] += x0 * __uint_as_float(wa4[k] & 0xFFFF0000u) + x1 * __uint_as_float(wb4[k] & 0xFFFF0000u);
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a] = warp_sum(acc[a]);
    nz = __any_sync(0xffffffffu, nz);
    if (lane < A) {
      float v = acc[0];
#pragma unroll
      for (int a = 1; a < A; ++a) v = (lane == a) ? acc[a] : v;
      wl[t * A + lane] = (num_frames || nz) ? v : -INFINITY;     // non-zero-frame mask (zt/frame_level_models.py:4372-4375)
    }
  }
  __syncthreads();
  // ---- softmax over the valid frames: warp a handles head a ----
  if (warp < A) {
    const int a = warp;
    float mx = -INFINITY;
    for (int t = lane; t < nf; t += 32) mx = fmaxf(mx, wl[t * A + a]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int t = lane; t < nf; t += 32) {
      const float l = wl[t * A + a];
      const float e = (l != -INFINITY) ? __expf(l - mx) : 0.0f;
      wl[t * A + a] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) inv_s[a] = 1.0f / sum;                        // sum == 0 only for an all-masked video: NaN like 0/0 in TF
  }
  __syncthreads();
  // ---- phase B: a thread owns feature columns 4 j .. 4 j + 3 of every G-th frame (G = thread groups of D / 4 threads: 3 for
  //      1024 threads at D = 1152); the groups' partial sums are folded into group 0 through shared memory, in group order ----
  const int tpg = D / 4;                                   // threads per group
  const int grp = tid / tpg, j = tid - grp * tpg;
  const int c0 = 4 * j;
  const bool active = grp < G;
  float acc[A][4];
#pragma unroll
  for (int a = 0; a < A; ++a) { acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.0f; }
  if (active) {
#pragma unroll 4
    for (int t = grp; t < nf; t += G) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(xb + static_cast<long long>(t) * D + c0));
      const float x0 = __uint_as_float(u.x << 16), x1 = __uint_as_float(u.x & 0xFFFF0000u);
      const float x2 = __uint_as_float(u.y << 16), x3 = __uint_as_float(u.y & 0xFFFF0000u);
      const float4 p0 = *reinterpret_cast<const float4*>(wl + t * A), p1 = *reinterpret_cast<const float4*>(wl + t * A + 4);   // broadcast reads
      const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const float pa = p[a];
        acc[a][0] += pa * x0; acc[a][1] += pa * x1; acc[a][2] += pa * x2; acc[a][3] += pa * x3;
      }
    }
  }
  float4* fold = reinterpret_cast<float4*>(fold_raw);      // [A][D / 4] float4
  for (int g = 1; g < G; ++g) {
    __syncthreads();
    if (grp == g) {
#pragma unroll
      for (int a = 0; a < A; ++a) fold[a * tpg + j] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const float4 v = fold[a * tpg + j];
        acc[a][0] += v.x; acc[a][1] += v.y; acc[a][2] += v.z; acc[a][3] += v.w;
      }
    }
  }
  if (grp == 0) {
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const float s = inv_s[a];
      const float v0 = acc[a][0] * s, v1 = acc[a][1] * s, v2 = acc[a][2] * s, v3 = acc[a][3] * s;
      const long long o = (static_cast<long long>(b) * A + a) * D + c0;
      if (out) *reinterpret_cast<float4*>(out + o) = make_float4(v0, v1, v2, v3);
      if (out_hi) {
        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(v0, h0, l0); split_bf16(v1, h1, l1); split_bf16(v2, h2, l2); split_bf16(v3, h3, l3);
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
      }
    }
  }
}

}  // namespace

extern "C" int yt8m_attn_pool_fused(const yt8m_bf16* x, const yt8m_bf16* w_packed, long long ldw, const int* num_frames, int B, int T,
                                    int D, int A, float* out, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && w_packed && (out || out_hi), YT8M_E_BADPTR, "yt8m_attn_pool_fused: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && D > 0 && D % 8 == 0 && D <= 4096 && ldw >= D, YT8M_E_BADSHAPE,
               "yt8m_attn_pool_fused: bad shape B=%d T=%d D=%d", B, T, D);
  YT8M_REQUIRE(A == kMaxA, YT8M_E_UNSUPPORTED, "yt8m_attn_pool_fused: built for %d heads (moe_num_extend / lstm_attentions default), got %d",
               kMaxA, A);
  // phase A wants many warps (a frame each); phase B folds G groups of D / 4 threads (G = 3 at D = 1152): as many threads as
  // the kernel may have (1024), but at least one full group; the fold buffer only when it fits
  int threads = 1024;
  if (D / 4 > threads) threads = ((D / 4 + 31) / 32) * 32;
  int G = std::max(1, std::min(threads / (D / 4), 4));
  const size_t base = (static_cast<size_t>(T) * A + 2 * A + 4) * sizeof(float) + static_cast<size_t>(D) * A * 2 + 16;
  if (G > 1 && base + static_cast<size_t>(D) * A * sizeof(float) > 200 * 1024) G = 1;
  const size_t smem = base + (G > 1 ? static_cast<size_t>(D) * A * sizeof(float) : 0);
  YT8M_REQUIRE(smem <= 200 * 1024, YT8M_E_UNSUPPORTED, "yt8m_attn_pool_fused: T * A + D * A too large for shared memory");
  auto kern = attn_pool_fused_kernel<kMaxA>;
  static bool attr_done = false;
  if (!attr_done) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  kern<<<B, threads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w_packed), ldw,
                                     num_frames, T, D, G, out, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                     reinterpret_cast<__nv_bfloat16*>(out_lo));
  return check_launch("attn_pool_fused_kernel");
}
