// yt8m_b200 -- backward of the NetVLAD pooling layer (definition: oracle/yt8m_oracle.py:netvlad_pool; not part of
// the reference) and the small pointwise backward pieces of the frame-level training step.
//
// Forward (per video b, frames t < n_b):  z = scale * (x . Cw) + shift;  a = softmax_K(z);  asum[k] = sum_t a[t,k]
//   V[d,k] = sum_t a[t,k] x[t,d] - asum[k] C2[d,k];   U[:,k] = V[:,k] * rs[k], rs = rsqrt(max(||V[:,k]||^2, 1e-12))
//   Y = U * gs, gs = rsqrt(max(||U||_F^2, 1e-12))                                    (the fused forward kernel)
// Backward, given dY:
//   (1) yt8m_netvlad_bwd_norm:   dV, dasum[k] = -sum_d dV[d,k] C2[d,k]              (two normalisations, residual)
//   (2) yt8m_netvlad_bwd_dcw2:   dC2[d,k] = -sum_b asum[b,k] dV[b,d,k]
//   (3) yt8m_netvlad_bwd_da:     da[t,k] = sum_d x[t,d] dV[d,k]                     (batched [T x D] . [D x K])
//   (4) yt8m_netvlad_bwd_softmax: g = da + dasum; dz = a * (g - sum_k a g), masked; dshift; dz * scale as bf16 hi/lo
//   (5) dCw^T[K, D] = (dz*scale)^T . X  through yt8m_wgrad (tcgen05, split over the B*T contraction rows)
// (3) + (4) here are the generic path (K = 32, or D not a multiple of 64): a SIMT tile kernel for da and a row kernel for the
// softmax backward, fed with z recomputed by yt8m_linear_fwd.  For K in {64, 128} the trainers use
// yt8m_netvlad_bwd_assign_fused (yt8m_netvlad_bwd_tc.cu): one tcgen05 kernel that recomputes the logits on chip, forms da on the
// tensor cores from dV as bf16 hi + lo (emitted by (1)) and does the softmax backward in its epilogue -- 498 us -> 56 us at
// BASELINE config 2 shapes.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {

constexpr int kNormThreads = 256;

// (1) one CTA per video.  stats[b] = {asum[K], ss[K], total}.  Two passes over the video's D x K slab; a thread owns FOUR
// adjacent clusters (16-byte loads, a warp covers 512 contiguous bytes) and keeps four rows in flight -- the first version
// (one float per thread, one row in flight) ran at 1.2 TB/s (profiles/r02d_train_step_launches.txt: 190 us for 226 MB).
template <int KC>
__global__ void __launch_bounds__(kNormThreads)
netvlad_bwd_norm_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ stats,
                        const float* __restrict__ cw2, int D, float* __restrict__ dv, float* __restrict__ dasum,
                        __nv_bfloat16* __restrict__ dv_hi, __nv_bfloat16* __restrict__ dv_lo) {
  constexpr int kTpr = KC / 4;                          // threads per row
  constexpr int kRows = kNormThreads / kTpr;            // rows per sweep
  __shared__ float4 red_a[kRows][kTpr], red_b[kRows][kTpr];
  __shared__ float coef_a[KC], coef_b[KC], col_r[KC], col_y2[KC];
  __shared__ float r_tot;
  const int b = blockIdx.x;
  const int c4 = threadIdx.x % kTpr, grp = threadIdx.x / kTpr;
  const long long base = static_cast<long long>(b) * D * KC;
  const float4* dy4 = reinterpret_cast<const float4*>(dy + base) + c4;
  const float4* y4 = reinterpret_cast<const float4*>(y + base) + c4;
  const float* st = stats + static_cast<long long>(b) * (2 * KC + 1);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f), y2 = r;
#pragma unroll 4
  for (int d = grp; d < D; d += kRows) {
    const float4 yy = __ldg(y4 + static_cast<long long>(d) * kTpr), dd = __ldg(dy4 + static_cast<long long>(d) * kTpr);
    r.x += dd.x * yy.x; r.y += dd.y * yy.y; r.z += dd.z * yy.z; r.w += dd.w * yy.w;
    y2.x += yy.x * yy.x; y2.y += yy.y * yy.y; y2.z += yy.z * yy.z; y2.w += yy.w * yy.w;
  }
  red_a[grp][c4] = r;
  red_b[grp][c4] = y2;
  __syncthreads();
  if (threadIdx.x < KC) {                               // fixed order: deterministic
    const int k = threadIdx.x;
    float sr = 0.0f, sy = 0.0f;
    for (int g = 0; g < kRows; ++g) {
      sr += reinterpret_cast<const float*>(&red_a[g][0])[k];
      sy += reinterpret_cast<const float*>(&red_b[g][0])[k];
    }
    col_r[k] = sr;
    col_y2[k] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int j = 0; j < KC; ++j) t += col_r[j];
    r_tot = t;
  }
  __syncthreads();
  if (threadIdx.x < KC) {
    const int k = threadIdx.x;
    const float ss = st[KC + k], total = st[2 * KC];
    const bool col_live = ss > 1e-12f, all_live = total > 1e-12f;
    const float rs = rsqrtf(fmaxf(ss, 1e-12f)), gs = rsqrtf(fmaxf(total, 1e-12f));
    const float R = all_live ? r_tot : 0.0f;                       // clamped norm: the scale is a constant
    // dU = gs (dY - R Y);  dV_k = rs (dU_k - (<dU_k, Y_k> / gs) Y_k)  with <dU_k, Y_k> = gs (r_k - R y2_k)
    const float q = col_live ? (col_r[k] - R * col_y2[k]) : 0.0f;
    coef_a[k] = rs * gs;                                           // multiplies dY
    coef_b[k] = -rs * (gs * R + q / gs);                           // multiplies Y
  }
  __syncthreads();
  const float4 ca = *reinterpret_cast<const float4*>(coef_a + 4 * c4), cb = *reinterpret_cast<const float4*>(coef_b + 4 * c4);
  const float4* c24 = reinterpret_cast<const float4*>(cw2) + c4;
  float4* dv4 = reinterpret_cast<float4*>(dv + base) + c4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int d = grp; d < D; d += kRows) {
    const long long o = static_cast<long long>(d) * kTpr;
    const float4 yy = __ldg(y4 + o), dd = __ldg(dy4 + o), cc = __ldg(c24 + o);
    float4 g;
    g.x = ca.x * dd.x + cb.x * yy.x; g.y = ca.y * dd.y + cb.y * yy.y; g.z = ca.z * dd.z + cb.z * yy.z; g.w = ca.w * dd.w + cb.w * yy.w;
    dv4[o] = g;
    if (dv_hi) {                                          // tensor-core operands of the assignment backward (yt8m_netvlad_bwd_tc.cu)
      __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
      split_bf16(g.x, h0, l0); split_bf16(g.y, h1, l1); split_bf16(g.z, h2, l2); split_bf16(g.w, h3, l3);
      const long long e = base + (o + c4) * 4;
      *reinterpret_cast<uint2*>(dv_hi + e) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
      *reinterpret_cast<uint2*>(dv_lo + e) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
    }
    s.x += g.x * cc.x; s.y += g.y * cc.y; s.z += g.z * cc.z; s.w += g.w * cc.w;
  }
  __syncthreads();
  red_a[grp][c4] = s;
  __syncthreads();
  if (threadIdx.x < KC) {
    const int k = threadIdx.x;
    float t = 0.0f;
    for (int g = 0; g < kRows; ++g) t += reinterpret_cast<const float*>(&red_a[g][0])[k];
    dasum[static_cast<long long>(b) * KC + k] = -t;
  }
}

// (2) dC2[e] = -sum_b asum[b, e % K] * dV[b, e].  A thread owns four adjacent elements; the batch is cut into kDcGroups
// slices (blockIdx.y) whose partial sums are combined in a fixed order by the last slice to finish (a ticket per column
// block: no atomics on the data, the result does not depend on the arrival order).
constexpr int kDcGroups = 8;
__global__ void __launch_bounds__(256)
netvlad_bwd_dcw2_kernel(const float* __restrict__ dv, const float* __restrict__ stats, int B, long long n, int KC,
                        float* __restrict__ partial, unsigned int* __restrict__ tickets, float* __restrict__ dcw2) {
  const long long e4 = blockIdx.x * 256LL + threadIdx.x;             // float4 index
  const int g = blockIdx.y;
  const int per = (B + kDcGroups - 1) / kDcGroups;
  const int b0 = g * per, b1 = min(B, b0 + per);
  const bool live = e4 * 4 < n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const int k = static_cast<int>((e4 * 4) % KC);
    const float4* src = reinterpret_cast<const float4*>(dv) + e4;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) {
      const float4 v = __ldg(src + static_cast<long long>(b) * (n / 4));
      const float* ap = stats + static_cast<long long>(b) * (2 * KC + 1) + k;      // rows of 2K + 1 floats: not 16-byte aligned
      const float4 a = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
      acc.x -= a.x * v.x; acc.y -= a.y * v.y; acc.z -= a.z * v.z; acc.w -= a.w * v.w;
    }
    reinterpret_cast<float4*>(partial)[static_cast<long long>(g) * (n / 4) + e4] = acc;
  }
  __threadfence();
  __shared__ unsigned int last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + blockIdx.x, 1u);
  __syncthreads();
  if (last != kDcGroups - 1) return;
  __threadfence();
  if (live) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int gg = 0; gg < kDcGroups; ++gg) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial) + static_cast<long long>(gg) * (n / 4) + e4);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(dcw2)[e4] = t;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;                     // self-cleaning: the next launch finds zeros
}

// (3) da[b, t, :] = x[b, t, :] . dV[b]    CTA = (video, 64-frame tile): 64 x KC outputs, 256 threads, 4 x (KC/16) each.
template <int KC>
__global__ void __launch_bounds__(256)
netvlad_bwd_da_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dv, int T, int D, float* __restrict__ da) {
  constexpr int kTD = 32;                                 // D-chunk
  constexpr int kNC = KC / 16;                            // clusters per thread
  __shared__ float xs[kTD][64 + 1];                       // [d][t]
  __shared__ float vs[kTD][KC];                           // [d][k]
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // ty: frame quad (4 frames), tx: cluster group
  const __nv_bfloat16* xb = x + (static_cast<long long>(b) * T) * D;
  const float* vb = dv + static_cast<long long>(b) * D * KC;
  float acc[4][kNC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < kNC; ++j) acc[i][j] = 0.0f;
  for (int d0 = 0; d0 < D; d0 += kTD) {
    // x tile: 64 frames x 32 dims (bf16), 8 per thread
    {
      const int f = threadIdx.x / 4, c = (threadIdx.x % 4) * 8;
      const int t = t0 + f;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (t < T) u = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(t) * D + d0 + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xs[c + 2 * j][f] = __uint_as_float(w[j] << 16);
        xs[c + 2 * j + 1][f] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
    for (int i = threadIdx.x; i < kTD * KC / 4; i += 256) {
      const int d = i / (KC / 4), c4 = i % (KC / 4);
      *reinterpret_cast<float4*>(&vs[d][4 * c4]) = *reinterpret_cast<const float4*>(vb + static_cast<long long>(d0 + d) * KC + 4 * c4);
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < kTD; ++d) {
      float xv[4], vv[kNC];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[d][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < kNC; ++j) vv[j] = vs[d][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < kNC; ++j) acc[i][j] += xv[i] * vv[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= T) continue;
#pragma unroll
    for (int j = 0; j < kNC; ++j) da[(static_cast<long long>(b) * T + t) * KC + tx + 16 * j] = acc[i][j];
  }
}

// (4) one warp per frame row
template <int KC>
__global__ void __launch_bounds__(256)
netvlad_bwd_softmax_kernel(const float* __restrict__ z, const float* __restrict__ da, const float* __restrict__ dasum,
                           const int* __restrict__ num_frames, int B, int T, const float* __restrict__ scale,
                           __nv_bfloat16* __restrict__ dzs_hi, __nv_bfloat16* __restrict__ dzs_lo, float* __restrict__ dshift) {
  constexpr int kPer = KC / 32;
  __shared__ float sh[8][KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) part[j] = 0.0f;
  const long long rows = static_cast<long long>(B) * T;
  for (long long row = blockIdx.x * 8LL + warp; row < rows; row += gridDim.x * 8LL) {
    const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long long>(b) * T);
    const bool live = t < min(max(num_frames[b], 0), T);
    float zz[kPer], g[kPer];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      zz[j] = z[row * KC + lane + 32 * j];
      g[j] = da[row * KC + lane + 32 * j] + dasum[static_cast<long long>(b) * KC + lane + 32 * j];
      mx = fmaxf(mx, zz[j]);
    }
    mx = warp_max(mx);
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { zz[j] = __expf(zz[j] - mx); den += zz[j]; }
    den = warp_sum(den);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < kPer; ++j) { zz[j] /= den; s += zz[j] * g[j]; }
    s = warp_sum(s);
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const float dz = live ? zz[j] * (g[j] - s) : 0.0f;
      part[j] += dz;
      const float v = scale ? dz * scale[lane + 32 * j] : dz;
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      dzs_hi[row * KC + lane + 32 * j] = h;
      dzs_lo[row * KC + lane + 32 * j] = l;
    }
  }
  if (dshift) {
#pragma unroll
    for (int j = 0; j < kPer; ++j) sh[warp][lane + 32 * j] = part[j];
    __syncthreads();
    for (int k = threadIdx.x; k < KC; k += 256) {
      float t = 0.0f;
      for (int w = 0; w < 8; ++w) t += sh[w][k];
      atomicAdd(dshift + k, t);
    }
  }
}

// activation backward of a dense layer:  d_pre = dy * act'(y) * col_scale   (y = post-activation output) as bf16 hi/lo
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long rows, int cols, int act,
                               const float* __restrict__ col_scale, __nv_bfloat16* __restrict__ out_hi,
                               __nv_bfloat16* __restrict__ out_lo, long long ld_out) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float yy = y[i];
    float d = dy[i];
    switch (act) {
      case YT8M_ACT_RELU: d = yy > 0.0f ? d : 0.0f; break;
      case YT8M_ACT_RELU6: d = (yy > 0.0f && yy < 6.0f) ? d : 0.0f; break;
      case YT8M_ACT_SIGMOID: d *= yy * (1.0f - yy); break;
      case YT8M_ACT_TANH: d *= 1.0f - yy * yy; break;
      default: break;
    }
    if (col_scale) d *= col_scale[c];
    __nv_bfloat16 h, l;
    split_bf16(d, h, l);
    out_hi[r * ld_out + c] = h;
    if (out_lo) out_lo[r * ld_out + c] = l;
  }
}

}  // namespace

extern "C" {

int yt8m_netvlad_bwd_norm(const float* dy, const float* y, const float* stats, const float* cw2, int B, int D, int K, float* dv,
                          float* dasum, float* dcw2, yt8m_bf16* dv_hi_, yt8m_bf16* dv_lo_, yt8m_stream_t stream_) {
  __nv_bfloat16* dv_hi = reinterpret_cast<__nv_bfloat16*>(dv_hi_);
  __nv_bfloat16* dv_lo = reinterpret_cast<__nv_bfloat16*>(dv_lo_);
  YT8M_REQUIRE((dv_hi == nullptr) == (dv_lo == nullptr), YT8M_E_BADPTR, "yt8m_netvlad_bwd_norm: dv_hi and dv_lo come together");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(dy && y && stats && cw2 && dv && dasum, YT8M_E_BADPTR, "yt8m_netvlad_bwd_norm: null pointer");
  YT8M_REQUIRE(B > 0 && D > 0, YT8M_E_BADSHAPE, "yt8m_netvlad_bwd_norm: bad shape B=%d D=%d", B, D);
  switch (K) {
    case 32: netvlad_bwd_norm_kernel<32><<<B, kNormThreads, 0, stream>>>(dy, y, stats, cw2, D, dv, dasum, dv_hi, dv_lo); break;
    case 64: netvlad_bwd_norm_kernel<64><<<B, kNormThreads, 0, stream>>>(dy, y, stats, cw2, D, dv, dasum, dv_hi, dv_lo); break;
    case 128: netvlad_bwd_norm_kernel<128><<<B, kNormThreads, 0, stream>>>(dy, y, stats, cw2, D, dv, dasum, dv_hi, dv_lo); break;
    default:
      set_error("yt8m_netvlad_bwd_norm: K=%d unsupported (32, 64, 128)", K);
      return YT8M_E_UNSUPPORTED;
  }
  int rc = check_launch("netvlad_bwd_norm_kernel");
  if (rc != YT8M_OK || !dcw2) return rc;
  const long long n = static_cast<long long>(D) * K;
  const int blocks = static_cast<int>((n / 4 + 255) / 256);
  // library-owned scratch (per device): kDcGroups partial sums + one ticket per column block
  const size_t ticket_bytes = 4096;
  YT8M_REQUIRE(blocks * sizeof(unsigned int) <= ticket_bytes, YT8M_E_UNSUPPORTED, "yt8m_netvlad_bwd_norm: D * K too large");
  void* scratch = lib_scratch(kScratchDcw2, ticket_bytes + static_cast<size_t>(kDcGroups) * n * sizeof(float), ticket_bytes, stream);
  if (!scratch) return YT8M_E_CUDA;
  unsigned int* tickets = static_cast<unsigned int*>(scratch);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(scratch) + ticket_bytes);
  netvlad_bwd_dcw2_kernel<<<dim3(blocks, kDcGroups), 256, 0, stream>>>(dv, stats, B, n, K, partial, tickets, dcw2);
  return check_launch("netvlad_bwd_dcw2_kernel");
}

int yt8m_netvlad_bwd_assign(const yt8m_bf16* x, const int* num_frames, const float* z, const float* dv, const float* dasum,
                            const float* scale, int B, int T, int D, int K, float* da_ws, yt8m_bf16* dzs_hi, yt8m_bf16* dzs_lo,
                            float* dshift, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && num_frames && z && dv && dasum && da_ws && dzs_hi && dzs_lo, YT8M_E_BADPTR,
               "yt8m_netvlad_bwd_assign: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && D > 0 && D % 32 == 0, YT8M_E_BADSHAPE, "yt8m_netvlad_bwd_assign: bad shape B=%d T=%d D=%d", B, T, D);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dzs_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(dzs_lo);
  const dim3 grid((T + 63) / 64, B);
  const int sm_blocks = static_cast<int>(std::min<long long>((static_cast<long long>(B) * T + 7) / 8, 148 * 8));
  if (dshift) YT8M_CUDA(cudaMemsetAsync(dshift, 0, K * sizeof(float), stream));
#define YT8M_CASE(KC)                                                                                              \
  case KC:                                                                                                          \
    netvlad_bwd_da_kernel<KC><<<grid, 256, 0, stream>>>(xb, dv, T, D, da_ws);                                       \
    if (int rc = check_launch("netvlad_bwd_da_kernel")) return rc;                                                 \
    netvlad_bwd_softmax_kernel<KC><<<sm_blocks, 256, 0, stream>>>(z, da_ws, dasum, num_frames, B, T, scale, hi, lo, dshift); \
    return check_launch("netvlad_bwd_softmax_kernel");
  switch (K) {
    YT8M_CASE(32)
    YT8M_CASE(64)
    YT8M_CASE(128)
    default:
      set_error("yt8m_netvlad_bwd_assign: K=%d unsupported (32, 64, 128)", K);
      return YT8M_E_UNSUPPORTED;
  }
#undef YT8M_CASE
}

int yt8m_act_bwd(const float* dy, const float* y, long long rows, int cols, int act, const float* col_scale, yt8m_bf16* out_hi,
                 yt8m_bf16* out_lo, long long ld_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(dy && y && out_hi, YT8M_E_BADPTR, "yt8m_act_bwd: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld_out >= cols, YT8M_E_BADSHAPE, "yt8m_act_bwd: bad shape");
  YT8M_REQUIRE(act >= YT8M_ACT_NONE && act <= YT8M_ACT_TANH, YT8M_E_UNSUPPORTED, "yt8m_act_bwd: unknown activation %d", act);
  const long long total = rows * cols;
  act_bwd_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8)), 256, 0, stream>>>(
      dy, y, rows, cols, act, col_scale, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), ld_out);
  return check_launch("act_bwd_kernel");
}

}  // extern "C"
