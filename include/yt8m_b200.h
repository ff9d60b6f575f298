/* yt8m_b200 -- C ABI of libyt8m_b200.so: the B200 (sm_100a) hot path of wangheda/youtube-8m.
 *
 * The reference has no FFI: its hot path is TensorFlow-1.0 graph ops built inside the model plugins'
 * create_model() (youtube-8m-wangheda/models.py:17-21).  Each entry point below replaces the group of
 * TF ops cited beside it ("wh/" = youtube-8m-wangheda/, "zt/" = youtube-8m-zhangteng/); the Python
 * plugins in youtube-8m_b200/ bind them with ctypes (see INTEGRATION.md for the stub a maintainer of the
 * reference would add).
 *
 * Conventions
 *   - every function returns 0 (YT8M_OK) or a negative YT8M_E_* code; yt8m_last_error() returns a
 *     thread-local, human-readable message for the last failure on the calling thread;
 *   - all tensor arguments are CALLER-OWNED DEVICE pointers (row-major, explicit sizes / row strides in
 *     elements); kernels are enqueued on `stream` and the call returns without synchronising;
 *   - no allocation, no global state: scratch memory is passed in (`workspace`, size from the matching
 *     *_workspace_bytes query), so calls are re-entrant (one stream per rank);
 *   - "bf16" pointers are uint16_t-sized IEEE bfloat16; activations that must keep fp32-level precision
 *     through the tensor cores are passed as a hi/lo bf16 pair (x ~= hi + lo); `lo` may be NULL;
 *   - weights are PACKED ONCE (yt8m_*_pack_*) from the reference/TensorFlow [in, out] layout into the
 *     K-contiguous [out, in] bf16 layout the tcgen05 kernels consume.
 */
#ifndef YT8M_B200_H_
#define YT8M_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YT8M_OK 0
#define YT8M_E_BADSHAPE (-1)
#define YT8M_E_BADPTR (-2)
#define YT8M_E_CUDA (-3)
#define YT8M_E_UNSUPPORTED (-4)

typedef void* yt8m_stream_t; /* cudaStream_t */
typedef uint16_t yt8m_bf16;

/* activation codes for yt8m_linear_fwd (slim.fully_connected activation_fn) */
#define YT8M_ACT_NONE 0
#define YT8M_ACT_RELU 1
#define YT8M_ACT_RELU6 2
#define YT8M_ACT_SIGMOID 3
#define YT8M_ACT_TANH 4

/* 16-bit operand formats of a GEMM (activation AND weight operand: the tensor core takes one 16-bit format per
 * instruction; fp16 x bf16 faults).
 *   YT8M_FMT_BF16: bf16 weights; activation bf16, optionally as a hi + lo pair (lo = bf16(v - hi): ~16 significant
 *                  bits, two MMAs per weight tile)
 *   YT8M_FMT_F16 : IEEE fp16 in the same 2-byte storage, 11 significant bits, ONE MMA per tile, no lo tensor -- for
 *                  bounded activations (L2-normalised descriptors, ReLU6 outputs).  The weight operand is then the
 *                  fp16 conversion of the packed bf16 weights (exact for |w| >= 2^-17, < 2^-24 absolute below). */
#define YT8M_FMT_BF16 0
#define YT8M_FMT_F16 1

/* library version (major*10000 + minor*100 + patch) */
int yt8m_version(void);
const char* yt8m_last_error(void);
/* number of CUDA kernels this library has launched in this process (successful launches only) */
long long yt8m_launch_count(void);

/* ---- frame-row transforms ----------------------------------------------------------------------
 * wh/all_feature_transform/default_transformer.py:5-8  (tf.nn.l2_normalize over the feature dim) and,
 * for src_dtype = U8, wh/utils.py:23-38 Dequantize(max=2, min=-2) fused in front of it
 * (reader path wh/readers.py:178-186).  out = x * rsqrt(max(sum x^2, 1e-12)); all-zero (padding)
 * rows stay zero.  src_dtype: 0 = fp32, 1 = bf16, 2 = uint8 (de-quantised first; rows at or beyond
 * num_frames[b] are written as zeros like the reader's padding when num_frames != NULL).
 * `normalize` = 0 skips the L2 normalisation (plain convert / de-quantise). */
#define YT8M_SRC_F32 0
#define YT8M_SRC_BF16 1
#define YT8M_SRC_U8 2
int yt8m_l2norm_rows_fwd(const void* x, int src_dtype, long long rows, int dim, int normalize,
                         const int* num_frames, int frames_per_video, yt8m_bf16* out_bf16, float* out_f32,
                         yt8m_stream_t stream);

/* RAGGED ingest of a frame-level batch (wh/readers.py:159-186 resize_axis + Dequantize + default_transformer.py:5-8):
 * `packed` holds only the real frames, uint8 [sum_b num_frames[b], dim], video b starting at row row_offsets[b]
 * (the zero padding of the reference's reader never crosses PCIe); the output is the padded, de-quantised,
 * L2-normalised [B, T, dim] the poolers consume (rows t >= num_frames[b] are zeros).  num_frames[b] <= T. */
int yt8m_frames_unpack_u8(const uint8_t* packed, const long long* row_offsets, const int* num_frames, int B, int T,
                          int dim, int normalize, yt8m_bf16* out_bf16, float* out_f32, yt8m_stream_t stream);

/* backward of the row L2 normalisation y = x * rsqrt(max(sum x^2, 1e-12)) (tf.nn.l2_normalize, e.g.
 * wh/all_video_models/deep_combine_chain_model.py:41): dx = dy / ||x|| - x (x . dy) / ||x||^3; x, dy, dx fp32 [rows, dim]. */
int yt8m_l2norm_rows_bwd(const float* x, const float* dy, long long rows, int dim, float* dx, yt8m_stream_t stream);

/* ---- dense layer (slim.fully_connected / tf.matmul + bias / folded batch-norm + activation) ----
 * out[M, N] = act((A[M, K] . W[N, K]^T) * col_scale[N] + col_shift[N]);  A = a_hi (+ a_lo).
 * Replaces e.g. wh/all_video_models/logistic_model.py:23-25, wh/all_frame_models/dbof_model.py:76-115,
 * wh/all_video_models/deep_combine_chain_model.py:31-36.  lda and ldw must be multiples of 8 elements (16-byte TMA
 * strides) and >= K; K itself is free (out-of-range columns are zero-filled by the tensor map).
 * Outputs (any subset non-NULL): fp32, bf16 hi, bf16 lo, all with row stride ld_out.
 * workspace: optional split-K scratch (>= yt8m_linear_workspace_bytes) -- used when the tile grid
 * alone cannot fill the GPU. */
size_t yt8m_linear_workspace_bytes(int M, int N, int K);
int yt8m_linear_fwd(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* w,
                    long long ldw, int M, int N, int K, const float* col_scale, const float* col_shift,
                    int act, int a_fmt, int out_fmt, float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo,
                    long long ld_out, void* workspace, size_t workspace_bytes, yt8m_stream_t stream);

/* fp32 [K, N] (TF layout) -> bf16 [N, ldw] (K contiguous, zero padded to ldw) */
int yt8m_pack_transpose_bf16(const float* w_kn, int K, int N, yt8m_bf16* w_packed, long long ldw,
                             yt8m_stream_t stream);

/* ---- MoeModel head: wh/all_video_models/moe_model.py:38-65 -------------------------------------
 * p[b, v] = sum_{m<M} softmax_{M+1}(x . Wg)[v, m] * sigmoid(x . We + be)[v, m]
 * Packed weight: rows grouped in tiles of 128; a tile holds CPT = floor(128 / (2M+1)) classes,
 * class-major, each class = [gate_0..gate_M, expert_0..expert_{M-1}], zero padded to 128 rows.
 * yt8m_moe_packed_rows() = 128 * ceil(V / CPT).  num_mixtures in {1, 2, 3, 4, 8}. */
long long yt8m_moe_packed_rows(int vocab, int num_mixtures);
int yt8m_moe_pack_weights(const float* gate_w, const float* expert_w, const float* expert_b, int D, int vocab,
                          int num_mixtures, yt8m_bf16* w_packed, long long ldw, float* bias_packed,
                          yt8m_stream_t stream);
int yt8m_moe_fwd(const yt8m_bf16* x_hi, const yt8m_bf16* x_lo, long long ldx, const yt8m_bf16* w_packed,
                 long long ldw, const float* bias_packed, int B, int D, int vocab, int num_mixtures, int x_fmt,
                 float* out, long long ld_out, yt8m_stream_t stream);

/* max over groups of `heads` consecutive rows: out[b, :] = max_a in[b*heads + a, :]
 * (wh/all_frame_models/lstm_attention_max_pooling_model.py:65-66, zt/video_level_models.py:2327-2328) */
int yt8m_group_max_rows(const float* in, long long groups, int heads, int cols, float* out, yt8m_stream_t stream);

/* backward of yt8m_group_max_rows: din[b*heads + a, :] = dout[b, :] where head a is the first to attain the maximum, else 0 */
int yt8m_group_max_rows_bwd(const float* in, const float* dout, long long groups, int heads, int cols, float* din,
                            yt8m_stream_t stream);

/* ---- LstmModel / LstmMemoryModel: wh/all_frame_models/lstm_model.py:30-47 ------------------------
 * MultiRNNCell[BasicLSTMCell(H, forget_bias)] x L under dynamic_rnn(sequence_length = num_frames).
 * Per layer one packed matrix [4H, in_l + H] (in_0 = D, in_l = H): row 4u+g holds gate g (i, j, f, o) of
 * unit u, columns = [x | h] as in TF's [in+H, 4H] kernel; b_packed in the same row order.
 * state_out: [B, L*2*H] = [c0, h0, c1, h1, ...] (lstm_model.py state_is_tuple=False layout).
 * out_seq / out_seq_bf (nullable): top-layer outputs [B, T, H] (zero for t >= num_frames[b]). */
int yt8m_lstm_pack_weights(const float* w_tf, const float* b_tf, int in_dim, int H, yt8m_bf16* w_packed,
                           float* b_packed, yt8m_stream_t stream);
size_t yt8m_lstm_workspace_bytes(int B, int T, int D, int H, int L);
int yt8m_lstm_fwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                  const yt8m_bf16* const* w_packed, const float* const* b_packed, float forget_bias,
                  float* state_out, float* out_seq, yt8m_bf16* out_seq_bf, void* workspace,
                  size_t workspace_bytes, yt8m_stream_t stream);

/* Training: yt8m_lstm_fwd_train is yt8m_lstm_fwd that RETAINS every layer's output sequence in caller buffers
 * seq_hi[l] / seq_lo[l] (bf16 hi/lo, [B, T, H], zero for t >= num_frames[b]); it needs the persistent recurrence
 * (H in {256, 512, 768, 1024}).  yt8m_lstm_bwd is back-propagation through time of the same stack (the tf.gradients
 * of wh/train.py:440-442 through wh/all_frame_models/lstm_model.py:30-47):
 *   wt_packed[l]: bf16 [in_l + H, 4H] = the transpose of w_packed[l] (yt8m_pack_transpose_bf16 of the fp32 master);
 *   dstate (nullable): dL/d state_out [B, L*2*H];  dout_seq (nullable): dL/d out_seq [B, T, H] of the top layer;
 *   dw[l]: fp32 [4H, in_l + H], db[l]: fp32 [4H] -- gradients in the packed layout of the forward weights.
 * Gate pre-activations are recomputed from the retained sequences with one GEMM per layer; the reverse recurrence
 * runs one cell-backward kernel and one tensor-core GEMM (dh_{t-1} = dG_t . Wh) per step. */
int yt8m_lstm_fwd_train(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                        const yt8m_bf16* const* w_packed, const float* const* b_packed, float forget_bias,
                        float* state_out, float* out_seq, yt8m_bf16* const* seq_hi, yt8m_bf16* const* seq_lo,
                        void* workspace, size_t workspace_bytes, yt8m_stream_t stream);
size_t yt8m_lstm_bwd_workspace_bytes(int B, int T, int D, int H, int L);
int yt8m_lstm_bwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int H, int L,
                  const yt8m_bf16* const* w_packed, const float* const* b_packed, const yt8m_bf16* const* wt_packed,
                  float forget_bias, const yt8m_bf16* const* seq_hi, const yt8m_bf16* const* seq_lo, const float* dstate,
                  const float* dout_seq, float* const* dw, float* const* db, void* workspace, size_t workspace_bytes,
                  yt8m_stream_t stream);

/* ---- attention pooling over frames --------------------------------------------------------------
 * out[b, a, :] = sum_t w[b, t, a] * feats[b, t, :]
 *   mode 0 (softmax over T, masked, renormalised):
 *        wh/all_frame_models/lstm_attention_max_pooling_model.py:58-63, zt/frame_level_models.py:4393-4397
 *   mode 1 (sigmoid gate, masked, / (sum + 1e-8)): wh/all_frame_models/lstm_multi_attention_model.py:66-78
 * logits: [B, T, ld_logits>=A] fp32 (already includes the bias).  Mask: t < num_frames[b] when
 * num_frames != NULL, else "frame row has a non-zero entry" (zt/frame_level_models.py:4372-4375).
 * feats: bf16 [B, T, F]. out: fp32 [B, A, F] (+ optional bf16 hi/lo copies as MoE operands). */
int yt8m_attn_pool_fwd(const float* logits, long long ld_logits, const yt8m_bf16* feats, const int* num_frames,
                       int B, int T, int A, int F, int mode, float* out, yt8m_bf16* out_hi, yt8m_bf16* out_lo,
                       yt8m_stream_t stream);

/* The LSTM-free attention pooler of zt/frame_level_models.py:4372-4398 in ONE kernel (BASELINE.json configs[4]): logits
 * x[b, t, :] . W[:, a] (w_packed: bf16 [A, ldw >= D], K-major as from yt8m_pack_transpose_bf16 of W[:D]; the mean-pooled half of
 * Attention/W and Attention/b shift every frame of a video alike and cancel in the softmax over t), masked softmax over the
 * frames (t < num_frames[b], or -- num_frames NULL -- "frame row has a non-zero entry"), weighted sum of the frames.  The frames
 * leave HBM once (the pooling pass re-reads them from L2).  A = 8 heads; D % 8 == 0.  out: fp32 [B, A, D] (+ bf16 hi / lo). */
int yt8m_attn_pool_fused(const yt8m_bf16* x, const yt8m_bf16* w_packed, long long ldw, const int* num_frames, int B, int T, int D,
                         int A, float* out, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream);

/* backward of yt8m_attn_pool_fwd (same logits / feats / mask / mode): dout fp32 [B, A, F] ->
 * dlogits fp32 [B, T, ld_dl >= A] (zero for masked frames) and, when dfeats != NULL, dfeats fp32 [B, T, F]
 * (needed when feats are LSTM outputs: lstm_attention_max_pooling_model.py:63). */
int yt8m_attn_pool_bwd(const float* logits, long long ld_logits, const yt8m_bf16* feats, const int* num_frames, int B,
                       int T, int A, int F, int mode, const float* dout, float* dlogits, long long ld_dl, float* dfeats,
                       yt8m_stream_t stream);

/* ---- NetVLAD (not in the reference; definition in oracle/yt8m_oracle.py:netvlad_pool) -----------
 * FUSED soft-assignment GEMM + masked softmax over K + residual aggregation GEMM + intra-norm +
 * final L2 norm.  x: bf16 [B, T, D]; cw_packed: bf16 [K, D]; scale/shift: [K] (folded BN or bias);
 * cw2: fp32 [D, K]; cw2_hi / cw2_lo (nullable): its bf16 hi/lo split [D, K] -- with them, K = 64 and a dense
 * output (ld_out = D*K) the residual runs on the tensor cores and the descriptor is written through TMA.
 * out: [B, D*K] (D-major, K-minor).  K in {32, 64, 128}, D % 128 == 0, T <= 384.
 * out_fmt = YT8M_FMT_F16: out_hi receives fp16 (out_lo must be NULL).
 * With K = 64, a single 16-bit output (no out_lo, no out_f32), a dense output (ld_out = D*K), D <= 1152 and T > 64 the
 * ONE-PASS cluster kernel runs (csrc/yt8m_netvlad_v4.cu): every frame is read from HBM once, only the
 * ceil(num_frames[b] / 32) tiles that hold real frames are streamed, and videos are scheduled longest first; frame rows
 * inside the last streamed tile but at or beyond num_frames[b] must be finite (the reader's zero padding).
 * out_hi (+ out_lo) double as the stash of the un-normalised descriptor, which the final rescale reads back: out_f32
 * therefore carries the precision of the stash (bf16 hi alone: 8 bits; hi + lo: ~16 bits; fp16: 11 bits).
 * stats (nullable): fp32 [B, 2K+1] = {a_sum[K], ||V[:,k]||^2 [K], ||U||_F^2} saved for yt8m_netvlad_bwd_norm. */
int yt8m_netvlad_fwd(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K,
                     const yt8m_bf16* cw_packed, const float* scale, const float* shift, const float* cw2,
                     const yt8m_bf16* cw2_hi, const yt8m_bf16* cw2_lo, float* out_f32, yt8m_bf16* out_hi,
                     yt8m_bf16* out_lo, long long ld_out, int out_fmt, float* stats, yt8m_stream_t stream);

/* The same layer with BLOCKED ("tiled") layouts for the two tensors its epilogue touches element by element, so that every
 * global access of the kernel is a contiguous 512-byte warp access straight from the accumulator registers (csrc/
 * yt8m_netvlad_v5.cu: four-CTA cluster per video, every frame read from HBM once):
 *   descriptor  element (d, k) of a video at  ((d / 32) * (K / 8) + k / 8) * 256 + (d % 32) * 8 + k % 8      [16-bit values]
 *   cw2_tiled   element (d, k)             at  ((d / 32) * (K / 4) + k / 4) * 128 + (d % 32) * 4 + k % 4      [fp32]
 * The descriptor is a permutation of the row-major [D, K] flattening: the layer that consumes it (the hidden FC) applies the
 * same permutation to its weight ROWS once, at packing time (the flatten order of a VLAD descriptor is a free choice).
 * K = 64, D % 64 == 0, 256 <= D <= 1280; out_fmt: YT8M_FMT_BF16 (hi only) or YT8M_FMT_F16; stats as for yt8m_netvlad_fwd.
 * workspace (>= yt8m_netvlad_tiled_workspace_bytes, may be NULL): scratch for the bf16 assignment [B, T, K].  With it (and
 * D <= 1152) the layer runs as TWO streaming kernels (csrc/yt8m_netvlad_v6.cu: assignment GEMM + softmax with the centres
 * resident in shared memory, then aggregation + normalisation by a four-CTA cluster per video); without it, as the one-pass
 * cluster kernel (csrc/yt8m_netvlad_v5.cu), whose per-tile chain of dependent steps is ~2x slower. */
int yt8m_netvlad_tiled_supported(int T, int D, int K);
size_t yt8m_netvlad_tiled_workspace_bytes(int B, int T, int D, int K);
int yt8m_netvlad_fwd_tiled(const yt8m_bf16* x, const int* num_frames, int B, int T, int D, int K, const yt8m_bf16* cw_packed,
                           const float* scale, const float* shift, const float* cw2_tiled, yt8m_bf16* out_tiled, int out_fmt,
                           float* stats, void* workspace, size_t workspace_bytes, yt8m_stream_t stream);

/* debug only: device buffer (>= 128 u64, or NULL to disable) that receives globaltimer stamps of the
 * NetVLAD kernel's phases for CTA 0 (tools/netvlad_timeline.py decodes them) */
int yt8m_debug_set_timeline(unsigned long long* dev_buf);
/* debug only: ablation switches of the NetVLAD kernel (results become wrong); 0 = normal operation */
int yt8m_debug_set_flags(int flags);

/* ---- elementwise glue -----------------------------------------------------------------------------
 * y = x * sigmoid(g * scale + shift)  (context gating; g = x . Wg from yt8m_linear_fwd) */
int yt8m_context_gate_fwd(const float* x, const float* g, const float* scale, const float* shift, long long rows,
                          int cols, float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream);

/* backward of yt8m_context_gate_fwd: dx = dy * sigmoid(z) (the direct path; nullable), dg = dy * x * sigmoid'(z) * scale
 * as fp32 (nullable) and/or bf16 hi/lo with row stride ld_dg (operands of the gate layer's dgrad / wgrad GEMMs). */
int yt8m_context_gate_bwd(const float* dy, const float* x, const float* g, const float* scale, const float* shift,
                          long long rows, int cols, float* dx, float* dg, yt8m_bf16* dg_hi, yt8m_bf16* dg_lo, long long ld_dg,
                          yt8m_stream_t stream);

/* y[i] += x[i]: where two gradient paths meet (e.g. the direct and the gate path of context gating) */
int yt8m_add_inplace(float* y, const float* x, long long n, yt8m_stream_t stream);

/* y = x * scale[c] + shift[c] on a contiguous [rows, cols] fp32 / bf16 matrix -> fp32 and/or bf16 hi (+lo):
 * inference-mode slim.batch_norm applied to a GEMM operand (wh/all_frame_models/dbof_model.py:64-70). */
int yt8m_col_affine(const void* x, int src_dtype, long long rows, int cols, const float* scale, const float* shift,
                    float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, yt8m_stream_t stream);

/* fp32 [rows, cols] -> bf16 hi (+lo) with destination row stride ld_out and column offset already applied
 * by the caller (used to build concatenated operands, e.g. chain_moe_model.py:16) */
int yt8m_split_bf16(const float* x, long long rows, int cols, long long ld_in, yt8m_bf16* out_hi,
                    yt8m_bf16* out_lo, long long ld_out, yt8m_stream_t stream);

/* CrossEntropyLoss (wh/losses.py:114-130): loss = mean_b sum_v -[y log(p+1e-5) + (1-y) log(1-p+1e-5)];
 * loss_out: 1 float (accumulated; zeroed by the call); dpred (nullable): dLoss/dp * grad_scale. */
int yt8m_xent_fwd_bwd(const float* pred, const float* labels, int B, int V, float* loss_out, float* dpred,
                      float grad_scale, yt8m_stream_t stream);

/* ---- training step (wh/train.py:440-466, wh/utils.py:164-174) ---------------------------------------
 * Master weights / gradients / Adam moments are kept in the PACKED [out, in] layout of the forward kernels.
 *
 * yt8m_logistic_bwd_dz: dz = dp * p * (1 - p) as bf16 hi/lo (sigmoid backward, logistic_model.py:23-25).
 * yt8m_wgrad: out[M, N] = A^T . B, A = a_hi (+ a_lo) stored [Kb, lda >= M], B stored [Kb, ldb >= N]: the
 *   contraction runs over the batch rows of both stored matrices (MN-major tcgen05 operands, no transposes),
 *   e.g. dW^T[V, D] = dZ^T . X.
 * yt8m_colsum_bf16: out[n] = sum_rows (hi + lo)[r, n]   (bias gradients).
 * yt8m_moe_bwd_dlogits: recomputes the MoE logits tile and writes dL/dlogits (packed column order, bf16
 *   hi/lo) from dL/dp -- backward of wh/all_video_models/moe_model.py:54-64.  num_mixtures in {1, 2, 4}.
 * yt8m_grad_reg_sumsq: grad += l2 * param; sums4[0..3] = {sum g^2 seg0, seg1, sum w^2 seg0, seg1}; for MoE packed
 *   weights (moe_per = 2M+1, moe_nmix = M) segment 0 = gate rows, 1 = expert rows (two tensors in the
 *   reference, hence two norms); moe_per = 0 -> one segment.  sums4 must hold YT8M_SUMS_FLOATS floats: beyond the four
 *   results it is the scratch of a fixed-order two-stage reduction (no floating-point atomics), so the sums -- and with them
 *   the clip scales -- are bit-identical from launch to launch and on every data-parallel rank.
 * yt8m_clip_adam_step: g *= clip / max(||g||_seg, clip)  (tf.clip_by_norm per tensor), then TF-1.0 Adam
 *   (theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller), and the
 *   bf16 operand copy is refreshed in place.  only_segment >= 0 restricts the update (MoE bias rows). */
#define YT8M_SUMS_FLOATS 4104 /* 8 + 4 * 1024 partial sums */
int yt8m_logistic_bwd_dz(const float* dp, const float* p, int B, int V, yt8m_bf16* dz_hi, yt8m_bf16* dz_lo, long long ld,
                         yt8m_stream_t stream);
int yt8m_wgrad(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* b, long long ldb, int M, int N,
               int Kb, float* out, long long ld_out, yt8m_stream_t stream);
int yt8m_colsum_bf16(const yt8m_bf16* hi, const yt8m_bf16* lo, long long ld, int rows, int cols, float* out,
                     yt8m_stream_t stream);
int yt8m_moe_bwd_dlogits(const yt8m_bf16* x_hi, const yt8m_bf16* x_lo, long long ldx, const yt8m_bf16* w_packed, long long ldw,
                         const float* bias_packed, const float* dp, long long ld_dp, int B, int D, int vocab, int num_mixtures,
                         yt8m_bf16* dl_hi, yt8m_bf16* dl_lo, long long ld_dl, yt8m_stream_t stream);
int yt8m_grad_reg_sumsq(float* grad, const float* param, long long rows, int row_len, float l2, int moe_per, int moe_nmix,
                        float* sums4, yt8m_stream_t stream);
int yt8m_clip_adam_step(float* param, const float* grad, float* m, float* v, long long rows, int row_len, const float* sums4,
                        float clip, float lr_t, float beta1, float beta2, float eps, int moe_per, int moe_nmix, int only_segment,
                        yt8m_bf16* param_bf16, yt8m_stream_t stream);

/* ---- backward of the NetVLAD layer and of a dense layer's activation (frame-level training step) ----------
 * Definition: oracle/yt8m_oracle.py:netvlad_pool (not in the reference).  All tensors fp32 unless noted.
 * yt8m_netvlad_bwd_norm:   dy, y [B, D*K] (y = the forward's fp32 output), stats from the forward, cw2 [D, K] ->
 *   dv [B, D*K] = dL/dV (V = un-normalised descriptor), dasum [B, K] = dL/da_sum, dcw2 [D, K] (nullable) = dL/dcw2;
 *   dv_hi / dv_lo (nullable, together): the same dv as bf16 hi + lo [B, D, K], the tensor-core operand of
 *   yt8m_netvlad_bwd_assign_fused.  dcw2 uses a library-owned scratch buffer per device: calls that write dcw2 must be
 *   ordered on one stream per device.
 * yt8m_netvlad_bwd_assign: x bf16 [B, T, D], z [B*T, K] = scale * (x . Cw) + shift (recomputed with
 *   yt8m_linear_fwd), dv, dasum -> dzs bf16 hi/lo [B*T, K] = dL/d(x . Cw) (i.e. dz * scale; rows t >= num_frames
 *   are zero), dshift [K] (nullable) = dL/dshift.  da_ws: scratch fp32 [B*T, K].  dL/dCw^T [K, D] then is
 *   yt8m_wgrad(dzs, x).  K in {32, 64, 128}, D % 32 == 0.
 * yt8m_netvlad_bwd_assign_fused: the same result in ONE tcgen05 kernel that recomputes the logits itself (no z, no
 *   scratch): cw_packed bf16 [K, ldcw >= D] and scale / shift [K] (nullable) are the forward's operands, dv_hi / dv_lo
 *   come from yt8m_netvlad_bwd_norm.  K in {64, 128}, D % 64 == 0 (yt8m_netvlad_bwd_assign_fused_supported).
 * yt8m_act_bwd: d_pre = dy * act'(y) * col_scale  (y = post-activation output of yt8m_linear_fwd) as bf16 hi/lo. */
int yt8m_netvlad_bwd_norm(const float* dy, const float* y, const float* stats, const float* cw2, int B, int D, int K,
                          float* dv, float* dasum, float* dcw2, yt8m_bf16* dv_hi, yt8m_bf16* dv_lo, yt8m_stream_t stream);
int yt8m_netvlad_bwd_assign_fused_supported(int T, int D, int K);
int yt8m_netvlad_bwd_assign_fused(const yt8m_bf16* x, const int* num_frames, const yt8m_bf16* cw_packed, long long ldcw,
                                  const float* scale, const float* shift, const yt8m_bf16* dv_hi, const yt8m_bf16* dv_lo,
                                  const float* dasum, int B, int T, int D, int K, yt8m_bf16* dzs_hi, yt8m_bf16* dzs_lo,
                                  float* dshift, yt8m_stream_t stream);
int yt8m_netvlad_bwd_assign(const yt8m_bf16* x, const int* num_frames, const float* z, const float* dv, const float* dasum,
                            const float* scale, int B, int T, int D, int K, float* da_ws, yt8m_bf16* dzs_hi,
                            yt8m_bf16* dzs_lo, float* dshift, yt8m_stream_t stream);
int yt8m_act_bwd(const float* dy, const float* y, long long rows, int cols, int act, const float* col_scale,
                 yt8m_bf16* out_hi, yt8m_bf16* out_lo, long long ld_out, yt8m_stream_t stream);

/* ---- training-mode batch normalisation (slim.batch_norm(is_training=True); wh/all_frame_models/dbof_model.py:64-108,
 * update ops wh/train.py:449-456).  x: fp32 or bf16 [rows, ld >= cols].
 * yt8m_bn_stats: mean[c], var[c] = mean / BIASED variance of every column over the rows (tf.nn.moments).
 * yt8m_bn_fold:  scale = gamma * rsqrt(var + eps), shift = beta - mean * scale (gamma / beta NULL = 1 / 0); when moving_mean /
 *   moving_var are given they are updated in place: moving <- decay * moving + (1 - decay) * batch (slim decay 0.999).
 * yt8m_col_affine_act: out = act(x * scale + shift) as fp32 and/or bf16 hi (+lo), row stride ld_out.
 * yt8m_bn_bwd: backward of y = act(gamma * (x - mean) * rsqrt(var + eps) + beta) THROUGH the batch statistics, given dy = dL/dy
 *   and y (NULL when act = none):  g = dy * act'(y);  dbeta = sum_r g;  dgamma = sum_r g * xhat;
 *   dx = gamma * rstd * (g - dbeta / rows - xhat * dgamma / rows) as fp32 and/or bf16 hi (+lo) (all nullable). */
int yt8m_bn_stats(const void* x, int src_dtype, long long rows, int cols, long long ld, float* mean, float* var, yt8m_stream_t stream);
int yt8m_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int cols, float* scale,
                 float* shift, float* moving_mean, float* moving_var, float decay, yt8m_stream_t stream);
int yt8m_col_affine_act(const void* x, int src_dtype, long long rows, int cols, long long ld, const float* scale, const float* shift,
                        int act, float* out_f32, yt8m_bf16* out_hi, yt8m_bf16* out_lo, long long ld_out, yt8m_stream_t stream);
int yt8m_bn_bwd(const float* dy, long long ld_dy, const float* y, long long ld_y, const void* x, int src_dtype, long long ld_x,
                const float* mean, const float* var, float eps, const float* gamma, int act, long long rows, int cols,
                float* dgamma, float* dbeta, float* dx_f32, yt8m_bf16* dx_hi, yt8m_bf16* dx_lo, long long ld_dx, yt8m_stream_t stream);

/* top-k per row, descending (wh/inference.py:76-87 format_lines; wh/eval_util.py:164 top_k_triplets).
 * k <= 32.  idx_out: int32 [rows, k]; val_out: fp32 [rows, k]. */
int yt8m_topk_rows(const float* x, long long rows, int cols, int k, int* idx_out, float* val_out,
                   yt8m_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* YT8M_B200_H_ */
