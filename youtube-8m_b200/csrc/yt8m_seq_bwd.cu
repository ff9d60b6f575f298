// yt8m_b200 -- backward of the sequence poolers (sm_100a): LSTM back-propagation through time, attention
// pooling, context gating.  Together with yt8m_train.cu / yt8m_netvlad_bwd.cu this closes the train.py step
// (wh/train.py:440-466: tf.gradients over the model graph) for every pooler of SURVEY.md §8(a).
//
// LSTM (wh/all_frame_models/lstm_model.py:30-47; BasicLSTMCell / dynamic_rnn semantics in oracle/yt8m_oracle.py):
// the forward keeps only every layer's output sequence h (bf16 hi/lo).  The backward of a layer
//   1. rebuilds the concatenated operand  A_t = [in_t | h_{t-1}]  (h shifted by one frame, zero at t = 0) and
//      recomputes ALL gate pre-activations with ONE tensor-core GEMM  G = A . W^T + b  over the B*T rows -- the
//      recurrence is not replayed, because h_{t-1} is known;
//   2. a forward scan (elementwise, one thread per cell) turns G into the gate activations in place and stores c_t;
//   3. the reverse recurrence: per step one elementwise kernel (cell backward -> dG_t as bf16 hi/lo) and one
//      tensor-core GEMM  dh_{t-1} = dG_t . Wh  (K = 4H, split-K);
//   4. the weight gradient  dW^T = dG^T . A  is ONE MN-major GEMM over all B*T rows, the bias gradient a column
//      sum, and the gradient handed to the layer below  dIn = dG . Wx  one more GEMM.
// Rows with t >= num_frames[b] are frozen in the forward: they contribute no dG and the state gradient passes
// through them unchanged.
#include "yt8m_common.cuh"
#include "yt8m_host.h"

#include <algorithm>

using namespace yt8m;

namespace {

constexpr int kNumSms = 148;

// ------------------------------------------------ LSTM ------------------------------------------------------------

// G[b, t, 4u + {0,1,2,3}] (pre-activations i, j, f, o) -> activations in place; c_seq[b, t, u] = c_t
__global__ void lstm_bwd_scan_kernel(float* __restrict__ G, float* __restrict__ c_seq, const int* __restrict__ num_frames,
                                     int B, int T, int H, float forget_bias) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * H) return;
  const int b = static_cast<int>(idx / H), u = static_cast<int>(idx % H);
  const int nf = min(max(num_frames[b], 0), T);
  float4* g4 = reinterpret_cast<float4*>(G) + static_cast<long long>(b) * T * H + u;
  float* cs = c_seq + static_cast<long long>(b) * T * H + u;
  float c = 0.0f;
  for (int t = 0; t < nf; ++t) {
    const float4 g = g4[static_cast<long long>(t) * H];
    float4 a;
    a.x = sigmoidf_(g.x);
    a.y = tanhf_(g.y);
    a.z = sigmoidf_(g.z + forget_bias);
    a.w = sigmoidf_(g.w);
    c = c * a.z + a.x * a.y;
    g4[static_cast<long long>(t) * H] = a;
    cs[static_cast<long long>(t) * H] = c;
  }
}

struct LstmBwdStep {
  const float* acts;        // [B, T, 4H] gate activations (unit-major)
  const float* c_seq;       // [B, T, H]
  float* dh_acc;            // [B, H]: dG_{t+1} . Wh, accumulated by split-K atomics; the reader zeroes what it consumed
  const float* dstate_c;    // [B] rows of stride ld_state: dL/dc_final of this layer
  const float* dstate_h;
  long long ld_state;
  const float* dout;        // nullable [B, T, H]: gradient of this layer's output sequence
  float* dc;                // [B, H] carried cell-state gradient
  __nv_bfloat16* dg_hi;     // [B, T, 4H]
  __nv_bfloat16* dg_lo;
  const int* num_frames;
  int B, T, H, t;
};

__global__ void lstm_bwd_step_kernel(const LstmBwdStep p) {
  // chained with the per-frame GEMM by programmatic dependent launch: let the GEMM of this frame set itself up now, and do not
  // read dh_acc before the GEMM of the frame above has completed (no-ops for a plain launch)
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(p.B) * p.H) return;
  const int b = static_cast<int>(idx / p.H), u = static_cast<int>(idx % p.H);
  const int nf = min(max(p.num_frames[b], 0), p.T);
  const int t = p.t;
  if (t >= nf) return;                       // frozen row: dG stays zero (memset), state gradient passes through
  const bool last_live = (t + 1 >= nf);      // the step above is frozen or does not exist: take dL/dstate
  const long long so = static_cast<long long>(b) * p.ld_state + u;
  float dh = last_live ? (p.dstate_h ? p.dstate_h[so] : 0.0f) : p.dh_acc[idx];
  // the next frame's product accumulates into this buffer: leave zeros behind (rows that are frozen or take dL/dstate never
  // received anything but zeros: their dG above is zero)
  if (!last_live) p.dh_acc[idx] = 0.0f;
  const float dc_in = last_live ? (p.dstate_c ? p.dstate_c[so] : 0.0f) : p.dc[idx];
  const long long row = static_cast<long long>(b) * p.T + t;
  if (p.dout) dh += p.dout[row * p.H + u];
  const float4 a = reinterpret_cast<const float4*>(p.acts)[row * p.H + u];
  const float c_t = p.c_seq[row * p.H + u];
  const float c_prev = t > 0 ? p.c_seq[(row - 1) * p.H + u] : 0.0f;
  const float tc = tanhf_(c_t);
  const float d_o = dh * tc * a.w * (1.0f - a.w);
  const float dct = dc_in + dh * a.w * (1.0f - tc * tc);
  const float d_i = dct * a.y * a.x * (1.0f - a.x);
  const float d_j = dct * a.x * (1.0f - a.y * a.y);
  const float d_f = dct * c_prev * a.z * (1.0f - a.z);
  p.dc[idx] = dct * a.z;
  __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
  split_bf16(d_i, h0, l0);
  split_bf16(d_j, h1, l1);
  split_bf16(d_f, h2, l2);
  split_bf16(d_o, h3, l3);
  const long long o = (row * p.H + u) * 4;
  *reinterpret_cast<uint2*>(p.dg_hi + o) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
  *reinterpret_cast<uint2*>(p.dg_lo + o) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
}

struct LstmBwdWs {
  float* G;                   // [B*T, 4H]
  float* c_seq;               // [B*T, H]
  __nv_bfloat16* dg_hi;       // [B*T, 4H]
  __nv_bfloat16* dg_lo;
  __nv_bfloat16* acat_hi;     // [B*T, in_max + H]
  __nv_bfloat16* acat_lo;
  float* dY[2];               // [B*T, H] gradient of a layer's output sequence (ping-pong between layers)
  float* dh_acc;              // [B, H]
  float* dc;                  // [B, H]
  void* splitk;
  size_t splitk_bytes;
  size_t total;
};
LstmBwdWs carve_bwd_ws(void* base, int B, int T, int D, int H, int L) {
  LstmBwdWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const size_t rows = static_cast<size_t>(B) * T;
  const size_t in_max = static_cast<size_t>(std::max(D, H));
  w.G = static_cast<float*>(take(rows * 4 * H * sizeof(float)));
  w.c_seq = static_cast<float*>(take(rows * H * sizeof(float)));
  w.dg_hi = static_cast<__nv_bfloat16*>(take(rows * 4 * H * 2));
  w.dg_lo = static_cast<__nv_bfloat16*>(take(rows * 4 * H * 2));
  w.acat_hi = static_cast<__nv_bfloat16*>(take(rows * (in_max + H) * 2));
  w.acat_lo = static_cast<__nv_bfloat16*>(take(rows * (in_max + H) * 2));
  for (int i = 0; i < 2; ++i) w.dY[i] = L > 1 ? static_cast<float*>(take(rows * H * sizeof(float))) : nullptr;
  w.dh_acc = static_cast<float*>(take(static_cast<size_t>(B) * H * sizeof(float)));
  w.dc = static_cast<float*>(take(static_cast<size_t>(B) * H * sizeof(float)));
  w.splitk_bytes = yt8m_linear_workspace_bytes(B, H, 4 * H);
  w.splitk = take(w.splitk_bytes);
  w.total = off;
  return w;
}

// ------------------------------------------- attention pooling ----------------------------------------------------
constexpr int kAttnMaxA = 16;
constexpr int kAttnBwdThreads = 256;

// one CTA per video.  Shared memory: w[T][A] (normalised weights), dw[T][A], dout[A][F], inv[A], r[A].
template <int A_MAX>
__global__ void __launch_bounds__(kAttnBwdThreads)
attn_pool_bwd_kernel(const float* __restrict__ logits, long long ld_logits, const __nv_bfloat16* __restrict__ feats,
                     const int* __restrict__ num_frames, int T, int A, int F, int mode, const float* __restrict__ dout,
                     float* __restrict__ dlogits, long long ld_dl, float* __restrict__ dfeats) {
  extern __shared__ float sm[];
  const size_t ta_pad = (static_cast<size_t>(T) * A + 3) & ~size_t(3);      // keeps dout 16-byte aligned
  float* sw = sm;                               // [T][A]
  float* sdw = sw + ta_pad;                     // [T][A]
  float* sdo = sdw + ta_pad;                    // [A][F]
  float* inv = sdo + static_cast<size_t>(A) * F;// [A]
  float* rr = inv + A;                          // [A]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* fb = feats + static_cast<long long>(b) * T * F;
  const float* lb = logits + static_cast<long long>(b) * T * ld_logits;
  const int nf = num_frames ? min(max(num_frames[b], 0), T) : T;

  for (int i = tid; i < T * A; i += kAttnBwdThreads) {
    const int t = i / A, a = i - t * A;
    sw[i] = lb[static_cast<long long>(t) * ld_logits + a];
    sdw[i] = 0.0f;
  }
  for (int i = tid; i < A * F; i += kAttnBwdThreads) sdo[i] = dout[static_cast<long long>(b) * A * F + i];
  __syncthreads();
  if (!num_frames) {
    // mask = "frame row has a non-zero entry" (zt/frame_level_models.py:4372-4375)
    for (int t = warp; t < T; t += kAttnBwdThreads / 32) {
      bool nz = false;
      const uint4* row = reinterpret_cast<const uint4*>(fb + static_cast<long long>(t) * F);
      for (int c = lane; c < F / 8; c += 32) {
        const uint4 u = row[c];
        nz |= ((u.x | u.y | u.z | u.w) & 0x7FFF7FFFu) != 0u;
      }
      nz = __any_sync(0xffffffffu, nz);
      if (!nz && lane < A) sw[t * A + lane] = -INFINITY;
    }
    __syncthreads();
  }
  if (tid < A) {
    const int a = tid;
    if (mode == 0) {
      float mx = -INFINITY;
      for (int t = 0; t < nf; ++t) mx = fmaxf(mx, sw[t * A + a]);
      float sum = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float l = sw[t * A + a];
        const float e = (t < nf && l != -INFINITY) ? __expf(l - mx) : 0.0f;
        sw[t * A + a] = e;
        sum += e;
      }
      inv[a] = 1.0f / sum;
    } else {
      float sum = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float l = sw[t * A + a];
        const float g = (t < nf && l != -INFINITY) ? sigmoidf_(l) : 0.0f;
        sw[t * A + a] = g;
        sum += g;
      }
      inv[a] = 1.0f / (sum + 1e-8f);
    }
  }
  __syncthreads();
  // dw[t, a] = dout[a, :] . feats[t, :];   dfeats[t, :] = sum_a w[t, a] * dout[a, :]      (one warp per frame)
  for (int t = warp; t < T; t += kAttnBwdThreads / 32) {
    // padded / masked frames hold raw weight 0: they get no gradient and their feature row is not read
    float wt[A_MAX];
    bool any = false;
#pragma unroll
    for (int a = 0; a < A_MAX; ++a) {
      wt[a] = a < A ? sw[t * A + a] * inv[a] : 0.0f;
      any |= wt[a] != 0.0f;
    }
    float acc[A_MAX];
#pragma unroll
    for (int a = 0; a < A_MAX; ++a) acc[a] = 0.0f;
    for (int c = lane; c < F / 8; c += 32) {
      float x[8];
      if (any) {
        const uint4 u = *reinterpret_cast<const uint4*>(fb + static_cast<long long>(t) * F + 8 * c);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          x[2 * j] = __uint_as_float(w4[j] << 16);
          x[2 * j + 1] = __uint_as_float(w4[j] & 0xFFFF0000u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.0f;
      }
      float df[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) df[j] = 0.0f;
#pragma unroll
      for (int a = 0; a < A_MAX; ++a) {
        if (a < A) {
          const float4 d0 = *reinterpret_cast<const float4*>(sdo + a * F + 8 * c);
          const float4 d1 = *reinterpret_cast<const float4*>(sdo + a * F + 8 * c + 4);
          const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[a] += d[j] * x[j];
            df[j] += wt[a] * d[j];
          }
        }
      }
      if (dfeats) {
        float4* o = reinterpret_cast<float4*>(dfeats + (static_cast<long long>(b) * T + t) * F + 8 * c);
        o[0] = make_float4(df[0], df[1], df[2], df[3]);
        o[1] = make_float4(df[4], df[5], df[6], df[7]);
      }
    }
#pragma unroll
    for (int a = 0; a < A_MAX; ++a) {
      if (a < A) {
        const float s = warp_sum(acc[a]);
        if (lane == 0) sdw[t * A + a] = s;
      }
    }
  }
  __syncthreads();
  if (tid < A) {
    float r = 0.0f;
    for (int t = 0; t < T; ++t) r += sw[t * A + tid] * inv[tid] * sdw[t * A + tid];
    rr[tid] = r;
  }
  __syncthreads();
  for (int i = tid; i < T * A; i += kAttnBwdThreads) {
    const int t = i / A, a = i - t * A;
    const float raw = sw[i];                    // e (mode 0) or masked sigmoid (mode 1)
    float dl;
    if (mode == 0) {
      dl = raw * inv[a] * (sdw[i] - rr[a]);
    } else {
      dl = (sdw[i] - rr[a]) * inv[a] * raw * (1.0f - raw);
    }
    dlogits[(static_cast<long long>(b) * T + t) * ld_dl + a] = dl;
  }
}

// ------------------------------------------- context gating -------------------------------------------------------
__global__ void context_gate_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                                        const float* __restrict__ scale, const float* __restrict__ shift, long long rows,
                                        int cols, float* __restrict__ dx, float* __restrict__ dg, __nv_bfloat16* __restrict__ dg_hi,
                                        __nv_bfloat16* __restrict__ dg_lo, long long ld_dg) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float z = g[i];
    const float sc = scale ? scale[c] : 1.0f;
    z = z * sc + (shift ? shift[c] : 0.0f);
    const float s = sigmoidf_(z);
    const float d = dy[i];
    if (dx) dx[i] = d * s;
    const float dgv = d * x[i] * s * (1.0f - s) * sc;
    if (dg) dg[i] = dgv;
    if (dg_hi) {
      __nv_bfloat16 h, l;
      split_bf16(dgv, h, l);
      dg_hi[r * ld_dg + c] = h;
      if (dg_lo) dg_lo[r * ld_dg + c] = l;
    }
  }
}

// backward of the max over `heads` consecutive rows: the gradient goes to the first head that attains the maximum
__global__ void group_max_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, long long groups, int heads,
                                     int cols, float* __restrict__ din) {
  const long long total = groups * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long g = i / cols;
    const int c = static_cast<int>(i - g * cols);
    int best = 0;
    float m = in[(g * heads) * cols + c];
    for (int a = 1; a < heads; ++a) {
      const float v = in[(g * heads + a) * cols + c];
      if (v > m) { m = v; best = a; }
    }
    const float d = dout[i];
    for (int a = 0; a < heads; ++a) din[(g * heads + a) * cols + c] = a == best ? d : 0.0f;
  }
}

// backward of y = x * rsqrt(max(sum x^2, 1e-12)) (tf.nn.l2_normalize): one warp per row
__global__ void l2norm_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int dim,
                                       float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const float* xr = x + r * dim;
    const float* dr = dy + r * dim;
    float ss = 0.0f, dot = 0.0f;
    for (int c = lane; c < dim; c += 32) {
      const float v = xr[c];
      ss += v * v;
      dot += v * dr[c];
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float inv = rsqrtf(fmaxf(ss, 1e-12f));
    // d/dx [x * inv(ss)]: inv is constant below the epsilon clamp
    const float k = ss > 1e-12f ? dot * inv * inv * inv : 0.0f;
    for (int c = lane; c < dim; c += 32) dx[r * dim + c] = dr[c] * inv - xr[c] * k;
  }
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] += x[i];
}

}  // namespace

extern "C" {

int yt8m_group_max_rows_bwd(const float* in, const float* dout, long long groups, int heads, int cols, float* din,
                            yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(in && dout && din, YT8M_E_BADPTR, "yt8m_group_max_rows_bwd: null pointer");
  YT8M_REQUIRE(groups > 0 && heads > 0 && cols > 0, YT8M_E_BADSHAPE, "yt8m_group_max_rows_bwd: bad shape");
  const long long total = groups * cols;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, kNumSms * 8));
  group_max_bwd_kernel<<<blocks, 256, 0, stream>>>(in, dout, groups, heads, cols, din);
  return check_launch("group_max_bwd_kernel");
}

int yt8m_l2norm_rows_bwd(const float* x, const float* dy, long long rows, int dim, float* dx, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && dy && dx, YT8M_E_BADPTR, "yt8m_l2norm_rows_bwd: null pointer");
  YT8M_REQUIRE(rows >= 0 && dim > 0, YT8M_E_BADSHAPE, "yt8m_l2norm_rows_bwd: bad shape");
  if (rows == 0) return YT8M_OK;
  const int blocks = static_cast<int>(std::min<long long>((rows + 7) / 8, kNumSms * 16));
  l2norm_rows_bwd_kernel<<<blocks, 256, 0, stream>>>(x, dy, rows, dim, dx);
  return check_launch("l2norm_rows_bwd_kernel");
}

int yt8m_add_inplace(float* y, const float* x, long long n, yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(y && x, YT8M_E_BADPTR, "yt8m_add_inplace: null pointer");
  YT8M_REQUIRE(n >= 0, YT8M_E_BADSHAPE, "yt8m_add_inplace: n < 0");
  if (n == 0) return YT8M_OK;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, kNumSms * 16));
  add_inplace_kernel<<<blocks, 256, 0, stream>>>(y, x, n);
  return check_launch("add_inplace_kernel");
}

size_t yt8m_lstm_bwd_workspace_bytes(int B, int T, int D, int H, int L) {
  if (B <= 0 || T <= 0 || D <= 0 || H <= 0 || L <= 0 || L > 8) return 0;
  return carve_bwd_ws(nullptr, B, T, D, H, L).total;
}

int yt8m_lstm_bwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                  const yt8m_bf16* const* w_packed, const float* const* b_packed, const yt8m_bf16* const* wt_packed,
                  float forget_bias, const yt8m_bf16* const* seq_hi, const yt8m_bf16* const* seq_lo, const float* dstate,
                  const float* dout_seq, float* const* dw, float* const* db, void* workspace, size_t workspace_bytes,
                  yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(x && num_frames && w_packed && b_packed && wt_packed && seq_hi && seq_lo && dw && db && workspace, YT8M_E_BADPTR,
               "yt8m_lstm_bwd: null pointer");
  YT8M_REQUIRE(dstate || dout_seq, YT8M_E_BADPTR, "yt8m_lstm_bwd: neither a state gradient nor an output gradient");
  YT8M_REQUIRE(B > 0 && T > 0 && D > 0 && D % 8 == 0 && H % 32 == 0 && L >= 1 && L <= 8, YT8M_E_BADSHAPE,
               "yt8m_lstm_bwd: bad shape B=%d T=%d D=%d H=%d L=%d", B, T, D, H, L);
  YT8M_REQUIRE(workspace_bytes >= yt8m_lstm_bwd_workspace_bytes(B, T, D, H, L), YT8M_E_BADSHAPE,
               "yt8m_lstm_bwd: workspace too small");
  LstmBwdWs ws = carve_bwd_ws(workspace, B, T, D, H, L);
  const long long rows = static_cast<long long>(B) * T;
  const long long ld_state = static_cast<long long>(L) * 2 * H;
  const int cell_blocks = static_cast<int>((static_cast<long long>(B) * H + 255) / 256);
  int rc;
  for (int l = L - 1; l >= 0; --l) {
    const int in = l == 0 ? D : H;
    const long long ldc = in + H;
    const yt8m_bf16* in_hi = l == 0 ? x : seq_hi[l - 1];
    const yt8m_bf16* in_lo = l == 0 ? nullptr : seq_lo[l - 1];
    // 1. A = [in_t | h_{t-1}]
    YT8M_CUDA(cudaMemcpy2DAsync(ws.acat_hi, ldc * 2, in_hi, static_cast<size_t>(in) * 2, static_cast<size_t>(in) * 2, rows,
                                cudaMemcpyDeviceToDevice, stream));
    if (in_lo)
      YT8M_CUDA(cudaMemcpy2DAsync(ws.acat_lo, ldc * 2, in_lo, static_cast<size_t>(in) * 2, static_cast<size_t>(in) * 2, rows,
                                  cudaMemcpyDeviceToDevice, stream));
    else
      YT8M_CUDA(cudaMemset2DAsync(ws.acat_lo, ldc * 2, 0, static_cast<size_t>(in) * 2, rows, stream));
    if (rows > 1) {
      YT8M_CUDA(cudaMemcpy2DAsync(ws.acat_hi + ldc + in, ldc * 2, seq_hi[l], static_cast<size_t>(H) * 2, static_cast<size_t>(H) * 2,
                                  rows - 1, cudaMemcpyDeviceToDevice, stream));
      YT8M_CUDA(cudaMemcpy2DAsync(ws.acat_lo + ldc + in, ldc * 2, seq_lo[l], static_cast<size_t>(H) * 2, static_cast<size_t>(H) * 2,
                                  rows - 1, cudaMemcpyDeviceToDevice, stream));
    }
    YT8M_CUDA(cudaMemset2DAsync(ws.acat_hi + in, static_cast<size_t>(T) * ldc * 2, 0, static_cast<size_t>(H) * 2, B, stream));
    YT8M_CUDA(cudaMemset2DAsync(ws.acat_lo + in, static_cast<size_t>(T) * ldc * 2, 0, static_cast<size_t>(H) * 2, B, stream));
    //    G = A . W^T + b  (all frames at once)
    rc = yt8m_linear_fwd(reinterpret_cast<const yt8m_bf16*>(ws.acat_hi), reinterpret_cast<const yt8m_bf16*>(ws.acat_lo), ldc,
                         w_packed[l], ldc, static_cast<int>(rows), 4 * H, static_cast<int>(ldc), nullptr, b_packed[l],
                         YT8M_ACT_NONE, YT8M_FMT_BF16, YT8M_FMT_BF16, ws.G, nullptr, nullptr, 4 * H, nullptr, 0, stream_);
    if (rc != YT8M_OK) return rc;
    // 2. activations + cell states
    lstm_bwd_scan_kernel<<<cell_blocks, 256, 0, stream>>>(ws.G, ws.c_seq, num_frames, B, T, H, forget_bias);
    if ((rc = check_launch("lstm_bwd_scan_kernel")) != YT8M_OK) return rc;
    // 3. reverse recurrence
    YT8M_CUDA(cudaMemsetAsync(ws.dg_hi, 0, static_cast<size_t>(rows) * 4 * H * 2, stream));
    YT8M_CUDA(cudaMemsetAsync(ws.dg_lo, 0, static_cast<size_t>(rows) * 4 * H * 2, stream));
    YT8M_CUDA(cudaMemsetAsync(ws.dh_acc, 0, static_cast<size_t>(B) * H * sizeof(float), stream));
    LstmBwdStep sp{};
    sp.acts = ws.G; sp.c_seq = ws.c_seq; sp.dh_acc = ws.dh_acc;
    sp.dstate_c = dstate ? dstate + static_cast<long long>(l) * 2 * H : nullptr;
    sp.dstate_h = dstate ? dstate + static_cast<long long>(l) * 2 * H + H : nullptr;
    sp.ld_state = ld_state;
    sp.dout = (l == L - 1) ? dout_seq : ws.dY[l & 1];
    sp.dc = ws.dc; sp.dg_hi = ws.dg_hi; sp.dg_lo = ws.dg_lo; sp.num_frames = num_frames;
    sp.B = B; sp.T = T; sp.H = H;
    const yt8m_bf16* wt_rec = wt_packed[l] + static_cast<long long>(in) * 4 * H;      // rows [in, in+H) of W^T: Wh
    for (int t = T - 1; t >= 0; --t) {
      sp.t = t;
      if (t == T - 1) {
        lstm_bwd_step_kernel<<<cell_blocks, 256, 0, stream>>>(sp);         // follows memsets: a plain launch
      } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cell_blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        YT8M_CUDA(cudaLaunchKernelEx(&cfg, lstm_bwd_step_kernel, sp));
      }
      if ((rc = check_launch("lstm_bwd_step_kernel")) != YT8M_OK) return rc;
      if (t == 0) break;
      rc = linear_accumulate(reinterpret_cast<const yt8m_bf16*>(ws.dg_hi) + static_cast<long long>(t) * 4 * H,
                             reinterpret_cast<const yt8m_bf16*>(ws.dg_lo) + static_cast<long long>(t) * 4 * H,
                             static_cast<long long>(T) * 4 * H, wt_rec, 4 * H, B, H, 4 * H, ws.dh_acc, H, stream, /*pdl=*/true);
      if (rc != YT8M_OK) return rc;
    }
    // 4. parameter gradients over all frames, and the gradient of the layer below
    rc = yt8m_wgrad(reinterpret_cast<const yt8m_bf16*>(ws.dg_hi), reinterpret_cast<const yt8m_bf16*>(ws.dg_lo), 4 * H,
                    reinterpret_cast<const yt8m_bf16*>(ws.acat_hi), ldc, 4 * H, static_cast<int>(ldc), static_cast<int>(rows), dw[l],
                    ldc, stream_);
    if (rc != YT8M_OK) return rc;
    rc = yt8m_colsum_bf16(reinterpret_cast<const yt8m_bf16*>(ws.dg_hi), reinterpret_cast<const yt8m_bf16*>(ws.dg_lo), 4 * H,
                          static_cast<int>(rows), 4 * H, db[l], stream_);
    if (rc != YT8M_OK) return rc;
    if (l > 0) {
      rc = yt8m_linear_fwd(reinterpret_cast<const yt8m_bf16*>(ws.dg_hi), reinterpret_cast<const yt8m_bf16*>(ws.dg_lo), 4 * H,
                           wt_packed[l], 4 * H, static_cast<int>(rows), in, 4 * H, nullptr, nullptr, YT8M_ACT_NONE, YT8M_FMT_BF16,
                           YT8M_FMT_BF16, ws.dY[(l - 1) & 1], nullptr, nullptr, in, nullptr, 0, stream_);
      if (rc != YT8M_OK) return rc;
    }
  }
  return YT8M_OK;
}

int yt8m_attn_pool_bwd(const float* logits, long long ld_logits, const yt8m_bf16* feats, const int* num_frames, int B, int T,
                       int A, int F, int mode, const float* dout, float* dlogits, long long ld_dl, float* dfeats,
                       yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(logits && feats && dout && dlogits, YT8M_E_BADPTR, "yt8m_attn_pool_bwd: null pointer");
  YT8M_REQUIRE(B > 0 && T > 0 && A > 0 && A <= kAttnMaxA && F > 0 && F % 8 == 0 && (mode == 0 || mode == 1) && ld_logits >= A &&
                   ld_dl >= A,
               YT8M_E_BADSHAPE, "yt8m_attn_pool_bwd: bad shape B=%d T=%d A=%d F=%d mode=%d", B, T, A, F, mode);
  const size_t smem = (2 * ((static_cast<size_t>(T) * A + 3) & ~size_t(3)) + static_cast<size_t>(A) * F + 2 * A) * sizeof(float);
  YT8M_REQUIRE(smem <= 200 * 1024, YT8M_E_UNSUPPORTED, "yt8m_attn_pool_bwd: T*A + A*F too large for shared memory");
  auto kern = attn_pool_bwd_kernel<kAttnMaxA>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    YT8M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  kern<<<B, kAttnBwdThreads, smem, stream>>>(logits, ld_logits, reinterpret_cast<const __nv_bfloat16*>(feats), num_frames, T, A, F,
                                             mode, dout, dlogits, ld_dl, dfeats);
  return check_launch("attn_pool_bwd_kernel");
}

int yt8m_context_gate_bwd(const float* dy, const float* x, const float* g, const float* scale, const float* shift, long long rows,
                          int cols, float* dx, float* dg, yt8m_bf16* dg_hi, yt8m_bf16* dg_lo, long long ld_dg,
                          yt8m_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  YT8M_REQUIRE(dy && x && g && (dx || dg || dg_hi), YT8M_E_BADPTR, "yt8m_context_gate_bwd: null pointer");
  YT8M_REQUIRE(rows > 0 && cols > 0 && (!dg_hi || ld_dg >= cols), YT8M_E_BADSHAPE, "yt8m_context_gate_bwd: bad shape");
  const long long total = rows * cols;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, kNumSms * 16));
  context_gate_bwd_kernel<<<blocks, 256, 0, stream>>>(dy, x, g, scale, shift, rows, cols, dx, dg,
                                                      reinterpret_cast<__nv_bfloat16*>(dg_hi),
                                                      reinterpret_cast<__nv_bfloat16*>(dg_lo), ld_dg);
  return check_launch("context_gate_bwd_kernel");
}

}  // extern "C"
