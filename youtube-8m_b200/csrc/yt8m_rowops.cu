// yt8m_b200 -- bandwidth-bound row kernels: frame-row L2-normalise / de-quantise, attention pooling,
// context gate, bf16 hi/lo split, cross-entropy, per-row top-k.  Coalesced 16-byte accesses, one warp per
// row or one CTA per (video, column slice); grids sized against the 148 SMs.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {
constexpr int kNumSms = 148;

// ------------------------------------------------------------------------------------------------
// l2norm rows (wh/all_feature_transform/default_transformer.py:5-8) with optional uint8 de-quantise
// (wh/utils.py:23-38).  One warp per row; two passes over the row (second hits L1).
// ------------------------------------------------------------------------------------------------
template <int SRC>
__device__ __forceinline__ void load8(const void* base, long long elem_off, float* v) {
  if (SRC == YT8M_SRC_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else if (SRC == YT8M_SRC_BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __uint_as_float(w[j] << 16);
      v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(base) + elem_off);
    const uint32_t w[2] = {u.x, u.y};
    // Dequantize(max=2, min=-2): q * (4/255) + (4/512 - 2)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = static_cast<float>((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * (4.0f / 255.0f) + (4.0f / 512.0f - 2.0f);
  }
}

// row_offsets != NULL: RAGGED source -- video b's frames are rows [row_offsets[b], row_offsets[b] + num_frames[b]) of x
// (the reader's packed batch: padding never crosses PCIe); the output stays the padded [B, frames_per_video, dim].
template <int SRC>
__global__ void l2norm_rows_kernel(const void* __restrict__ x, long long rows, int dim, int normalize,
                                   const int* __restrict__ num_frames, int frames_per_video,
                                   const long long* __restrict__ row_offsets,
                                   __nv_bfloat16* __restrict__ out_bf, float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int chunks = dim >> 3;
  for (long long r = warp_global; r < rows; r += nwarps) {
    bool pad = false;
    const long long base = r * dim;
    long long src = base;
    if (num_frames) {
      const long long b = r / frames_per_video;
      const long long t = r - b * frames_per_video;
      pad = t >= num_frames[b];
      if (row_offsets) src = (row_offsets[b] + t) * dim;
    }
    float scale = 1.0f;
    if (normalize && !pad) {
      float ss = 0.0f;
      for (int c = lane; c < chunks; c += 32) {
        float v[8];
        load8<SRC>(x, src + c * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
      }
      ss = warp_sum(ss);
      scale = rsqrtf(fmaxf(ss, 1e-12f));
    }
    for (int c = lane; c < chunks; c += 32) {
      float v[8];
      if (pad) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
      } else {
        load8<SRC>(x, src + c * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= scale;
      }
      if (out_bf) {
        uint4 hi, lo;
        pack8_hi_lo(v, hi, lo);
        *reinterpret_cast<uint4*>(out_bf + base + c * 8) = hi;
      }
      if (out_f32) {
        float4* o = reinterpret_cast<float4*>(out_f32 + base + c * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention pooling: one CTA per (video, 256-column slice).  The T x A weights are built in shared
// memory (softmax over T or sigmoid gate, masked, renormalised), then every thread streams its two
// feature columns over the frames with bf16x2 loads (128 B per warp request) and keeps A fp32
// accumulators per column in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnMaxA = 16;
constexpr int kAttnThreads = 128;

template <int A_MAX>
__global__ void __launch_bounds__(kAttnThreads)
attn_pool_kernel(const float* __restrict__ logits, long long ld_logits, const __nv_bfloat16* __restrict__ feats,
                 const int* __restrict__ num_frames, int T, int A, int F, int mode, float* __restrict__ out,
                 __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ float sw[];            // [T][A] weights, then [A] scratch x2
  float* red = sw + (size_t)T * A;         // [A] max / [A] sum
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const __nv_bfloat16* fb = feats + (long long)b * T * F;
  const float* lb = logits + (long long)b * T * ld_logits;
  const int nf = num_frames ? min(max(num_frames[b], 0), T) : T;

  // --- frame mask: sequence mask, or "row has a non-zero entry" (zt/frame_level_models.py:4372-4375)
  // stored temporarily in sw[t*A] sign: we keep a separate pass to stay simple.
  for (int i = tid; i < T * A; i += kAttnThreads) {
    const int t = i / A, a = i - t * A;
    sw[i] = lb[(long long)t * ld_logits + a];
  }
  __syncthreads();
  if (!num_frames) {
    // one warp per frame: any non-zero bf16 in the row?
    const int warp = tid >> 5, lane = tid & 31;
    for (int t = warp; t < T; t += kAttnThreads / 32) {
      bool nz = false;
      const uint4* row = reinterpret_cast<const uint4*>(fb + (long long)t * F);
      for (int c = lane; c < F / 8; c += 32) {
        const uint4 u = row[c];
        nz |= ((u.x | u.y | u.z | u.w) & 0x7FFF7FFFu) != 0u;
      }
      nz = __any_sync(0xffffffffu, nz);
      if (!nz && lane < A) sw[t * A + lane] = -INFINITY;        // masked frame (A <= 16 < 32 lanes)
    }
    __syncthreads();
  }
  // --- weights
  if (tid < A) {
    const int a = tid;
    if (mode == 0) {
      // softmax over all T, times mask, renormalised over T == softmax over the unmasked frames
      float mx = -INFINITY;
      for (int t = 0; t < nf; ++t) mx = fmaxf(mx, sw[t * A + a]);
      float sum = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float l = sw[t * A + a];
        const float e = (t < nf && l != -INFINITY) ? __expf(l - mx) : 0.0f;
        sw[t * A + a] = e;
        sum += e;
      }
      red[a] = 1.0f / sum;                 // sum == 0 only for an all-masked video: 0 * inf = NaN like 0/0 in TF
    } else {
      float sum = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float l = sw[t * A + a];
        const float g = (t < nf && l != -INFINITY) ? sigmoidf_(l) : 0.0f;
        sw[t * A + a] = g;
        sum += g;
      }
      red[a] = 1.0f / (sum + 1e-8f);
    }
  }
  __syncthreads();
  // --- pooled features: this thread owns columns c0, c0+1
  const int c0 = blockIdx.y * (2 * kAttnThreads) + 2 * tid;
  if (c0 >= F) return;
  float acc0[A_MAX], acc1[A_MAX];
#pragma unroll
  for (int a = 0; a < A_MAX; ++a) { acc0[a] = 0.0f; acc1[a] = 0.0f; }
  const int t_end = num_frames ? nf : T;
#pragma unroll 4
  for (int t = 0; t < t_end; ++t) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(fb + (long long)t * F + c0);
    const float x0 = __uint_as_float(u << 16), x1 = __uint_as_float(u & 0xFFFF0000u);
#pragma unroll
    for (int a = 0; a < A_MAX; ++a) {
      if (a < A) {
        const float w = sw[t * A + a];
        acc0[a] += w * x0;
        acc1[a] += w * x1;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < A_MAX; ++a) {
    if (a < A) {
      const float inv = red[a];
      const float v0 = acc0[a] * inv, v1 = acc1[a] * inv;
      const long long o = ((long long)b * A + a) * F + c0;
      if (out) { out[o] = v0; out[o + 1] = v1; }
      if (out_hi) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v0, h0, l0);
        split_bf16(v1, h1, l1);
        *reinterpret_cast<uint32_t*>(out_hi + o) = pack_bf16x2(h0, h1);
        if (out_lo) *reinterpret_cast<uint32_t*>(out_lo + o) = pack_bf16x2(l0, l1);
      }
    }
  }
}

__global__ void context_gate_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ scale,
                                    const float* __restrict__ shift, long long rows, int cols, float* __restrict__ out,
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % cols);
    float z = g[i];
    if (scale) z *= scale[c];
    if (shift) z += shift[c];
    const float y = x[i] * sigmoidf_(z);
    if (out) out[i] = y;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(y, h, l);
      out_hi[i] = h;
      if (out_lo) out_lo[i] = l;
    }
  }
}

// y = x * scale[c] + shift[c] -> bf16 hi/lo (+fp32): exact application of an inference-mode batch-norm
// to a GEMM operand (wh/all_frame_models/dbof_model.py:64-70 input_bn).
template <typename TIn>
__global__ void col_affine_kernel(const TIn* __restrict__ x, long long rows, int cols, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                  __nv_bfloat16* __restrict__ out_lo) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % cols);
    float v = static_cast<float>(x[i]);
    if (scale) v *= scale[c];
    if (shift) v += shift[c];
    if (out) out[i] = v;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      out_hi[i] = h;
      if (out_lo) out_lo[i] = l;
    }
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ x, long long rows, int cols, long long ld_in,
                                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ld_out) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    __nv_bfloat16 h, l;
    split_bf16(x[r * ld_in + c], h, l);
    out_hi[r * ld_out + c] = h;
    if (out_lo) out_lo[r * ld_out + c] = l;
  }
}

// CrossEntropyLoss, wh/losses.py:114-130 (epsilon = 10e-6)
__global__ void xent_kernel(const float* __restrict__ pred, const float* __restrict__ labels, long long total, float inv_b,
                            float* __restrict__ loss_out, float* __restrict__ dpred, float grad_scale) {
  const float eps = 10e-6f;
  float local = 0.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float p = pred[i], y = labels[i];
    local -= y * logf(p + eps) + (1.0f - y) * logf(1.0f - p + eps);
    if (dpred) dpred[i] = -(y / (p + eps) - (1.0f - y) / (1.0f - p + eps)) * inv_b * grad_scale;
  }
  local = warp_sum(local);
  __shared__ float part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = local;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? part[lane] : 0.0f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss_out, v * inv_b);
  }
}

// per-row top-k, one warp per row; k selection passes over the row (ties -> lower index first)
__global__ void topk_rows_kernel(const float* __restrict__ x, long long rows, int cols, int k, int* __restrict__ idx_out,
                                 float* __restrict__ val_out) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + row * cols;
  float last_v = INFINITY;
  int last_i = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7FFFFFFF;
    for (int c = lane; c < cols; c += 32) {
      const float v = xr[c];
      const bool after = (v < last_v) || (v == last_v && c > last_i);   // strictly after the previous pick
      if (after && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      idx_out[row * k + j] = (bi == 0x7FFFFFFF) ? -1 : bi;
      val_out[row * k + j] = bv;
    }
    last_v = bv;
    last_i = bi;
  }
}

int grid_for(long long total, int per_block) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((total + per_block - 1) / per_block, kNumSms * 16)));
}
}  // namespace

extern "C" {

int yt8m_l2norm_rows_fwd(const void* x, int src_dtype, long long rows, int dim, int normalize, const int* num_frames,
                         int frames_per_video, yt8m_bf16* out_bf16, float* out_f32, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && (out_bf16 || out_f32), YT8M_E_BADPTR, "yt8m_l2norm_rows_fwd: null pointer");
  YT8M_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0, YT8M_E_BADSHAPE, "yt8m_l2norm_rows_fwd: dim must be a multiple of 8");
  YT8M_REQUIRE(!num_frames || frames_per_video > 0, YT8M_E_BADSHAPE, "yt8m_l2norm_rows_fwd: frames_per_video");
  if (rows == 0) return YT8M_OK;
  const int threads = 256;
  const int blocks = grid_for(rows, threads / 32);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  switch (src_dtype) {
    case YT8M_SRC_F32: l2norm_rows_kernel<YT8M_SRC_F32><<<blocks, threads, 0, stream>>>(x, rows, dim, normalize, num_frames, frames_per_video, nullptr, ob, out_f32); break;
    case YT8M_SRC_BF16: l2norm_rows_kernel<YT8M_SRC_BF16><<<blocks, threads, 0, stream>>>(x, rows, dim, normalize, num_frames, frames_per_video, nullptr, ob, out_f32); break;
    case YT8M_SRC_U8: l2norm_rows_kernel<YT8M_SRC_U8><<<blocks, threads, 0, stream>>>(x, rows, dim, normalize, num_frames, frames_per_video, nullptr, ob, out_f32); break;
    default: set_error("yt8m_l2norm_rows_fwd: unknown src_dtype %d", src_dtype); return YT8M_E_UNSUPPORTED;
  }
  return check_launch("l2norm_rows_kernel");
}

int yt8m_frames_unpack_u8(const uint8_t* packed, const long long* row_offsets, const int* num_frames, int B, int T, int dim,
                          int normalize, yt8m_bf16* out_bf16, float* out_f32, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(packed && row_offsets && num_frames && (out_bf16 || out_f32), YT8M_E_BADPTR, "yt8m_frames_unpack_u8: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && dim > 0 && dim % 8 == 0, YT8M_E_BADSHAPE, "yt8m_frames_unpack_u8: bad shape B=%d T=%d dim=%d", B, T, dim);
  YT8M_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 7u) == 0, YT8M_E_BADPTR, "yt8m_frames_unpack_u8: packed must be 8-byte aligned");
  const long long rows = static_cast<long long>(B) * T;
  const int threads = 256;
  const int blocks = grid_for(rows, threads / 32);
  l2norm_rows_kernel<YT8M_SRC_U8><<<blocks, threads, 0, stream>>>(packed, rows, dim, normalize, num_frames, T, row_offsets,
                                                                   reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32);
  return check_launch("l2norm_rows_kernel");
}

int yt8m_attn_pool_fwd(const float* logits, long long ld_logits, const yt8m_bf16* feats, const int* num_frames, int B, int T,
                       int A, int F, int mode, float* out, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(logits && feats && (out || out_hi), YT8M_E_BADPTR, "yt8m_attn_pool_fwd: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && A > 0 && A <= kAttnMaxA && F % 8 == 0 && ld_logits >= A && (mode == 0 || mode == 1),
               YT8M_E_BADSHAPE, "yt8m_attn_pool_fwd: bad shape B=%d T=%d A=%d F=%d mode=%d", B, T, A, F, mode);
  const size_t smem = (static_cast<size_t>(T) * A + 2 * A) * sizeof(float);
  YT8M_REQUIRE(smem <= 200 * 1024, YT8M_E_UNSUPPORTED, "yt8m_attn_pool_fwd: T*A too large for shared memory");
  auto kern = attn_pool_kernel<kAttnMaxA>;
  if (smem > 48 * 1024) YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(B, (F + 2 * kAttnThreads - 1) / (2 * kAttnThreads));
  kern<<<grid, kAttnThreads, smem, stream>>>(logits, ld_logits, reinterpret_cast<const __nv_bfloat16*>(feats), num_frames, T, A,
                                            F, mode, out, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                            reinterpret_cast<__nv_bfloat16*>(out_lo));
  return check_launch("attn_pool_kernel");
}

int yt8m_context_gate_fwd(const float* x, const float* g, const float* scale, const float* shift, long long rows, int cols,
                          float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && g && (out_f32 || out_hi), YT8M_E_BADPTR, "yt8m_context_gate_fwd: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0, YT8M_E_BADSHAPE, "yt8m_context_gate_fwd: bad shape");
  context_gate_kernel<<<grid_for(rows * cols, 256), 256, 0, stream>>>(x, g, scale, shift, rows, cols, out_f32,
                                                                      reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                                      reinterpret_cast<__nv_bfloat16*>(out_lo));
  return check_launch("context_gate_kernel");
}

int yt8m_col_affine(const void* x, int src_dtype, long long rows, int cols, const float* scale, const float* shift,
                    float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && (out_f32 || out_hi), YT8M_E_BADPTR, "yt8m_col_affine: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0, YT8M_E_BADSHAPE, "yt8m_col_affine: bad shape");
  const int blocks = grid_for(rows * cols, 256);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  if (src_dtype == YT8M_SRC_F32)
    col_affine_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), rows, cols, scale, shift, out_f32, oh, ol);
  else if (src_dtype == YT8M_SRC_BF16)
    col_affine_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, scale, shift, out_f32, oh, ol);
  else { set_error("yt8m_col_affine: src_dtype %d unsupported", src_dtype); return YT8M_E_UNSUPPORTED; }
  return check_launch("col_affine_kernel");
}

int yt8m_split_bf16(const float* x, long long rows, int cols, long long ld_in, yt8m_bf16* out_hi, yt8m_bf16* out_lo,
                    long long ld_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && out_hi, YT8M_E_BADPTR, "yt8m_split_bf16: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && ld_in >= cols && ld_out >= cols, YT8M_E_BADSHAPE, "yt8m_split_bf16: bad shape");
  split_bf16_kernel<<<grid_for(rows * cols, 256), 256, 0, stream>>>(x, rows, cols, ld_in, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                                    reinterpret_cast<__nv_bfloat16*>(out_lo), ld_out);
  return check_launch("split_bf16_kernel");
}

int yt8m_xent_fwd_bwd(const float* pred, const float* labels, int B, int V, float* loss_out, float* dpred, float grad_scale,
                      yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(pred && labels && loss_out, YT8M_E_BADPTR, "yt8m_xent_fwd_bwd: null pointer");
  YT8M_REQUIRE(B > 0 && V > 0, YT8M_E_BADSHAPE, "yt8m_xent_fwd_bwd: bad shape");
  YT8M_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), stream));
  const long long total = static_cast<long long>(B) * V;
  xent_kernel<<<grid_for(total, 256 * 4), 256, 0, stream>>>(pred, labels, total, 1.0f / B, loss_out, dpred, grad_scale);
  return check_launch("xent_kernel");
}

int yt8m_topk_rows(const float* x, long long rows, int cols, int k, int* idx_out, float* val_out, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && idx_out && val_out, YT8M_E_BADPTR, "yt8m_topk_rows: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && k > 0 && k <= 32 && k <= cols, YT8M_E_BADSHAPE, "yt8m_topk_rows: need 0 < k <= min(32, cols)");
  const int threads = 128;
  const int blocks = static_cast<int>((rows * 32 + threads - 1) / threads);
  topk_rows_kernel<<<blocks, threads, 0, stream>>>(x, rows, cols, k, idx_out, val_out);
  return check_launch("topk_rows_kernel");
}

}  // extern "C"
