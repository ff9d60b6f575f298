// yt8m_b200 -- training-mode batch normalisation (slim.batch_norm(is_training=True), TF 1.0 non-fused form) for the layers of
// wh/all_frame_models/dbof_model.py:64-108: batch statistics over the rows of a [rows, C] matrix (biased variance), the folded
// affine + activation, the moving-average update (decay 0.999, wh/train.py:449-456 runs them as UPDATE_OPS), and the backward
// pass through the statistics.  All bandwidth-bound column kernels: a CTA owns 32 columns, lanes = columns (128-byte rows).
#include "yt8m_common.cuh"
#include "yt8m_host.h"

using namespace yt8m;

namespace {

constexpr int kColsPerCta = 32;
constexpr int kRowWarps = 16;               // warps of a CTA walk the rows in turn

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return static_cast<float>(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float act_grad(float y, int act) {
  // derivative of the activation expressed through its OUTPUT y (ReLU / ReLU6: 1 inside the open interval)
  switch (act) {
    case 1: return y > 0.0f ? 1.0f : 0.0f;
    case 2: return (y > 0.0f && y < 6.0f) ? 1.0f : 0.0f;
    case 3: return y * (1.0f - y);
    case 4: return 1.0f - y * y;
    default: return 1.0f;
  }
}
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.0f);
    case 2: return fminf(fmaxf(v, 0.0f), 6.0f);
    case 3: return 1.0f / (1.0f + __expf(-v));
    case 4: { const float e = __expf(-2.0f * fabsf(v)); return copysignf((1.0f - e) / (1.0f + e), v); }
    default: return v;
  }
}

// two column sums over the rows, deterministic: warp w of the CTA takes rows w, w + 16, ...; fixed-order tree over the warps
template <class F>
__device__ __forceinline__ void col_reduce2(long long rows, int col, bool col_ok, F f, float& s0, float& s1) {
  __shared__ float red[2][kRowWarps][kColsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a = 0.0f, b = 0.0f;
  if (col_ok)
    for (long long r = warp; r < rows; r += kRowWarps) f(r, col, a, b);
  red[0][warp][lane] = a;
  red[1][warp][lane] = b;
  __syncthreads();
  if (warp == 0) {
    float ta = 0.0f, tb = 0.0f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) { ta += red[0][w][lane]; tb += red[1][w][lane]; }
    s0 = ta; s1 = tb;
  }
}

// mean / biased variance of every column (two passes: the second one about the mean, as tf.nn.moments does)
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32) bn_stats_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                                  float* __restrict__ mean, float* __restrict__ var) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * kColsPerCta + lane;
  const bool ok = col < cols;
  float s = 0.0f, unused = 0.0f;
  col_reduce2(rows, col, ok, [&](long long r, int c, float& a, float&) { a += ldf(x + r * ld + c); }, s, unused);
  __shared__ float mean_s[kColsPerCta];
  if (warp == 0) mean_s[lane] = s / static_cast<float>(rows);
  __syncthreads();
  const float mu = mean_s[lane];
  float q = 0.0f;
  col_reduce2(rows, col, ok, [&](long long r, int c, float& a, float&) { const float d = ldf(x + r * ld + c) - mu; a += d * d; }, q, unused);
  if (warp == 0 && ok) {
    mean[col] = mu;
    var[col] = q / static_cast<float>(rows);
  }
}

// scale = gamma * rsqrt(var + eps), shift = beta - mean * scale; moving statistics <- decay * moving + (1 - decay) * batch
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int cols, float* __restrict__ scale, float* __restrict__ shift,
                               float* __restrict__ moving_mean, float* __restrict__ moving_var, float decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const float s = (gamma ? gamma[c] : 1.0f) * rsqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = (beta ? beta[c] : 0.0f) - mean[c] * s;
  if (moving_mean) moving_mean[c] = moving_mean[c] * decay + mean[c] * (1.0f - decay);
  if (moving_var) moving_var[c] = moving_var[c] * decay + var[c] * (1.0f - decay);
}

template <typename T>
__global__ void col_affine_act_kernel(const T* __restrict__ x, long long rows, int cols, long long ld, const float* __restrict__ scale,
                                      const float* __restrict__ shift, int act, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                      __nv_bfloat16* __restrict__ out_lo, long long ld_out) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float v = ldf(x + r * ld + c);
    if (scale) v *= scale[c];
    if (shift) v += shift[c];
    v = act_apply(v, act);
    if (out) out[r * ld_out + c] = v;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      out_hi[r * ld_out + c] = h;
      if (out_lo) out_lo[r * ld_out + c] = l;
    }
  }
}

// g = dy * act'(y);  sum_g[c] = sum_r g,  sum_gx[c] = sum_r g * xhat   (xhat = (x - mean) * rstd)
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32) bn_bwd_reduce_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ y,
                                                                       long long ld_y, const T* __restrict__ x, long long ld_x,
                                                                       const float* __restrict__ mean, const float* __restrict__ var,
                                                                       float eps, int act, long long rows, int cols,
                                                                       float* __restrict__ sum_g, float* __restrict__ sum_gx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * kColsPerCta + lane;
  const bool ok = col < cols;
  const float mu = ok ? mean[col] : 0.0f, rstd = ok ? rsqrtf(var[col] + eps) : 0.0f;
  float s0 = 0.0f, s1 = 0.0f;
  col_reduce2(rows, col, ok, [&](long long r, int c, float& a, float& b) {
    float g = dy[r * ld_dy + c];
    if (y) g *= act_grad(y[r * ld_y + c], act);
    a += g;
    b += g * (ldf(x + r * ld_x + c) - mu) * rstd;
  }, s0, s1);
  if (warp == 0 && ok) {
    sum_g[col] = s0;
    sum_gx[col] = s1;
  }
}

// dx = gamma * rstd * (g - sum_g / rows - xhat * sum_gx / rows)
template <typename T>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ y, long long ld_y,
                                    const T* __restrict__ x, long long ld_x, const float* __restrict__ mean, const float* __restrict__ var,
                                    float eps, const float* __restrict__ gamma, int act, const float* __restrict__ sum_g,
                                    const float* __restrict__ sum_gx, long long rows, int cols, float* __restrict__ dx,
                                    __nv_bfloat16* __restrict__ dx_hi, __nv_bfloat16* __restrict__ dx_lo, long long ld_dx) {
  const long long total = rows * cols;
  const float inv_rows = 1.0f / static_cast<float>(rows);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float rstd = rsqrtf(var[c] + eps);
    float g = dy[r * ld_dy + c];
    if (y) g *= act_grad(y[r * ld_y + c], act);
    const float xhat = (ldf(x + r * ld_x + c) - mean[c]) * rstd;
    const float v = (gamma ? gamma[c] : 1.0f) * rstd * (g - sum_g[c] * inv_rows - xhat * sum_gx[c] * inv_rows);
    if (dx) dx[r * ld_dx + c] = v;
    if (dx_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      dx_hi[r * ld_dx + c] = h;
      if (dx_lo) dx_lo[r * ld_dx + c] = l;
    }
  }
}

int grid_for(long long total, int per_block) {
  long long b = (total + per_block - 1) / per_block;
  const long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int yt8m_bn_stats(const void* x, int src_dtype, long long rows, int cols, long long ld, float* mean, float* var, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && mean && var, YT8M_E_BADPTR, "yt8m_bn_stats: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld >= cols, YT8M_E_BADSHAPE, "yt8m_bn_stats: bad shape rows=%lld cols=%d ld=%lld", rows, cols, ld);
  const int blocks = (cols + kColsPerCta - 1) / kColsPerCta;
  if (src_dtype == YT8M_SRC_F32)
    bn_stats_kernel<float><<<blocks, kRowWarps * 32, 0, stream>>>(static_cast<const float*>(x), rows, cols, ld, mean, var);
  else if (src_dtype == YT8M_SRC_BF16)
    bn_stats_kernel<__nv_bfloat16><<<blocks, kRowWarps * 32, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, ld, mean, var);
  else { set_error("yt8m_bn_stats: src_dtype %d unsupported", src_dtype); return YT8M_E_UNSUPPORTED; }
  return check_launch("bn_stats_kernel");
}

int yt8m_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int cols, float* scale,
                 float* shift, float* moving_mean, float* moving_var, float decay, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(mean && var && scale && shift, YT8M_E_BADPTR, "yt8m_bn_fold: null pointer");
  YT8M_REQUIRE(cols > 0, YT8M_E_BADSHAPE, "yt8m_bn_fold: bad shape");
  bn_fold_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(gamma, beta, mean, var, eps, cols, scale, shift, moving_mean, moving_var, decay);
  return check_launch("bn_fold_kernel");
}

int yt8m_col_affine_act(const void* x, int src_dtype, long long rows, int cols, long long ld, const float* scale, const float* shift,
                        int act, float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, long long ld_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && (out_f32 || out_hi), YT8M_E_BADPTR, "yt8m_col_affine_act: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld >= cols && ld_out >= cols, YT8M_E_BADSHAPE, "yt8m_col_affine_act: bad shape");
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  const int blocks = grid_for(rows * cols, 256);
  if (src_dtype == YT8M_SRC_F32)
    col_affine_act_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), rows, cols, ld, scale, shift, act, out_f32, oh, ol, ld_out);
  else if (src_dtype == YT8M_SRC_BF16)
    col_affine_act_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, ld, scale, shift, act, out_f32, oh, ol, ld_out);
  else { set_error("yt8m_col_affine_act: src_dtype %d unsupported", src_dtype); return YT8M_E_UNSUPPORTED; }
  return check_launch("col_affine_act_kernel");
}

int yt8m_bn_bwd(const float* dy, long long ld_dy, const float* y, long long ld_y, const void* x, int src_dtype, long long ld_x,
                const float* mean, const float* var, float eps, const float* gamma, int act, long long rows, int cols,
                float* dgamma, float* dbeta, float* dx_f32, yt8m_bf16* dx_hi, yt8m_bf16* dx_lo, long long ld_dx, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(dy && x && mean && var && dgamma && dbeta, YT8M_E_BADPTR, "yt8m_bn_bwd: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld_dy >= cols && ld_x >= cols && (!y || ld_y >= cols), YT8M_E_BADSHAPE, "yt8m_bn_bwd: bad shape");
  YT8M_REQUIRE(src_dtype == YT8M_SRC_F32 || src_dtype == YT8M_SRC_BF16, YT8M_E_UNSUPPORTED, "yt8m_bn_bwd: src_dtype %d unsupported", src_dtype);
  const bool want_dx = dx_f32 || dx_hi;
  YT8M_REQUIRE(!want_dx || ld_dx >= cols, YT8M_E_BADSHAPE, "yt8m_bn_bwd: ld_dx");
  const int blocks = (cols + kColsPerCta - 1) / kColsPerCta;
  // dbeta = sum g, dgamma = sum g * xhat: the two column sums ARE the parameter gradients
  if (src_dtype == YT8M_SRC_F32)
    bn_bwd_reduce_kernel<float><<<blocks, kRowWarps * 32, 0, stream>>>(dy, ld_dy, y, ld_y, static_cast<const float*>(x), ld_x, mean, var, eps, act,
                                                                       rows, cols, dbeta, dgamma);
  else
    bn_bwd_reduce_kernel<__nv_bfloat16><<<blocks, kRowWarps * 32, 0, stream>>>(dy, ld_dy, y, ld_y, static_cast<const __nv_bfloat16*>(x), ld_x, mean,
                                                                               var, eps, act, rows, cols, dbeta, dgamma);
  int rc = check_launch("bn_bwd_reduce_kernel");
  if (rc != YT8M_OK || !want_dx) return rc;
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(dx_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(dx_lo);
  const int b2 = grid_for(rows * cols, 256);
  if (src_dtype == YT8M_SRC_F32)
    bn_bwd_apply_kernel<float><<<b2, 256, 0, stream>>>(dy, ld_dy, y, ld_y, static_cast<const float*>(x), ld_x, mean, var, eps, gamma, act, dbeta,
                                                       dgamma, rows, cols, dx_f32, oh, ol, ld_dx);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<b2, 256, 0, stream>>>(dy, ld_dy, y, ld_y, static_cast<const __nv_bfloat16*>(x), ld_x, mean, var, eps, gamma,
                                                               act, dbeta, dgamma, rows, cols, dx_f32, oh, ol, ld_dx);
  return check_launch("bn_bwd_apply_kernel");
}

}  // extern "C"
