"""The C ABI validates arguments before touching CUDA: bad shapes / null pointers / unsupported variants return
the documented negative status and set yt8m_last_error() -- checked here on the CPU box with dummy non-null
pointers (never dereferenced because validation fails first)."""
import ctypes

import yt8m_native as nat

lib = nat._lib
P = 0x1000          # dummy non-null "device pointer"
OK, BADSHAPE, BADPTR, UNSUPPORTED = 0, -1, -2, -4


def err():
  return lib.yt8m_last_error().decode()


def test_linear_argument_checks():
  assert lib.yt8m_linear_fwd(None, None, 64, P, 64, 8, 8, 64, None, None, 0, 0, 0, P, None, None, 8, None, 0, None) == BADPTR
  assert "null" in err()
  assert lib.yt8m_linear_fwd(P, None, 64, P, 64, 0, 8, 64, None, None, 0, 0, 0, P, None, None, 8, None, 0, None) == BADSHAPE
  assert lib.yt8m_linear_fwd(P, None, 60, P, 64, 8, 8, 60, None, None, 0, 0, 0, P, None, None, 8, None, 0, None) == BADSHAPE   # lda % 8
  assert "multiples of 8" in err()
  assert lib.yt8m_linear_fwd(P, None, 64, P, 64, 8, 8, 64, None, None, 0, 0, 0, None, None, None, 8, None, 0, None) == BADPTR   # no output
  assert lib.yt8m_linear_fwd(P, None, 64, P, 64, 8, 16, 64, None, None, 0, 0, 0, P, None, None, 8, None, 0, None) == BADSHAPE  # ld_out < N


def test_moe_argument_checks():
  assert lib.yt8m_moe_packed_rows(4716, 2) == 128 * 189
  assert lib.yt8m_moe_packed_rows(0, 2) == -1 and lib.yt8m_moe_packed_rows(10, 0) == -1 and lib.yt8m_moe_packed_rows(10, 64) == -1
  assert lib.yt8m_moe_fwd(P, None, 64, P, 64, P, 4, 64, 100, 5, 0, P, 100, None) == UNSUPPORTED
  assert "num_mixtures=5" in err()
  assert lib.yt8m_moe_fwd(P, None, 64, P, 64, P, 0, 64, 100, 2, 0, P, 100, None) == BADSHAPE
  assert lib.yt8m_moe_fwd(P, None, 64, P, 64, P, 4, 64, 100, 2, 0, P, 50, None) == BADSHAPE      # ld_out < vocab
  assert lib.yt8m_moe_fwd(None, None, 64, P, 64, P, 4, 64, 100, 2, 0, P, 100, None) == BADPTR
  assert lib.yt8m_moe_pack_weights(P, P, P, 64, 100, 2, P, 60, P, None) == BADSHAPE             # ldw < D
  assert lib.yt8m_moe_bwd_dlogits(P, None, 64, P, 64, P, P, 100, 4, 64, 100, 8, P, P, 128 * 4, None) in (UNSUPPORTED, BADSHAPE)


def test_netvlad_lstm_attention_argument_checks():
  assert lib.yt8m_netvlad_fwd(P, P, 2, 300, 1152, 48, P, None, None, P, None, None, None, P, None, 1152 * 48, 0, None, None) == UNSUPPORTED
  assert "K=48" in err()
  assert lib.yt8m_netvlad_fwd(P, P, 2, 400, 1152, 64, P, None, None, P, None, None, None, P, None, 1152 * 64, 0, None, None) == BADSHAPE   # T > 384
  assert lib.yt8m_netvlad_fwd(P, P, 2, 300, 1100, 64, P, None, None, P, None, None, None, P, None, 1100 * 64, 0, None, None) == BADSHAPE   # D % 128
  assert lib.yt8m_netvlad_fwd(P, P, 2, 300, 1152, 64, P, None, None, P, None, None, None, None, None, 1152 * 64, 0, None, None) == BADPTR  # stash
  assert lib.yt8m_lstm_workspace_bytes(4, 10, 64, 32, 0) == 0 and lib.yt8m_lstm_workspace_bytes(4, 10, 64, 32, 9) == 0
  assert lib.yt8m_lstm_workspace_bytes(64, 300, 1152, 1024, 2) > 64 * 300 * 4096 * 4
  assert lib.yt8m_lstm_pack_weights(P, P, 60, 32, P, P, None) == BADSHAPE                       # in_dim % 8
  assert lib.yt8m_lstm_pack_weights(P, P, 64, 40, P, P, None) == BADSHAPE                       # H % 32
  assert lib.yt8m_attn_pool_fwd(P, 8, P, None, 2, 300, 17, 1152, 0, P, None, None, None) == BADSHAPE     # A > 16
  assert lib.yt8m_attn_pool_fwd(P, 8, P, None, 2, 300, 8, 1152, 2, P, None, None, None) == BADSHAPE      # mode
  assert lib.yt8m_attn_pool_fwd(P, 4, P, None, 2, 300, 8, 1152, 0, P, None, None, None) == BADSHAPE      # ld_logits < A


def test_row_kernel_argument_checks():
  assert lib.yt8m_l2norm_rows_fwd(P, 0, 4, 60, 1, None, 0, P, None, None) == BADSHAPE                     # dim % 8
  assert lib.yt8m_l2norm_rows_fwd(P, 7, 4, 64, 1, None, 0, P, None, None) == UNSUPPORTED or "src_dtype" in err()
  assert lib.yt8m_l2norm_rows_fwd(P, 0, 0, 64, 1, None, 0, P, None, None) == OK                           # empty input is fine
  assert lib.yt8m_topk_rows(P, 4, 10, 0, P, P, None) == BADSHAPE and lib.yt8m_topk_rows(P, 4, 10, 33, P, P, None) == BADSHAPE
  assert lib.yt8m_topk_rows(P, 4, 10, 11, P, P, None) == BADSHAPE                                          # k > cols
  assert lib.yt8m_xent_fwd_bwd(P, None, 4, 10, P, None, 1.0, None) == BADPTR
  assert lib.yt8m_wgrad(P, None, 60, P, 64, 8, 8, 8, P, 8, None) == BADSHAPE
  assert lib.yt8m_group_max_rows(P, 0, 8, 10, P, None) == BADSHAPE


def test_backward_and_ingest_argument_checks():
  """The entry points added for the training step and the ragged ingest validate before touching CUDA as well."""
  # ragged ingest
  assert lib.yt8m_frames_unpack_u8(None, P, P, 4, 300, 1152, 1, P, None, None) == BADPTR
  assert lib.yt8m_frames_unpack_u8(P, P, P, 4, 300, 1150, 1, P, None, None) == BADSHAPE                   # dim % 8
  assert lib.yt8m_frames_unpack_u8(P + 4, P, P, 4, 300, 1152, 1, P, None, None) == BADPTR                 # 8-byte alignment
  assert "aligned" in err()
  assert lib.yt8m_frames_unpack_u8(P, P, P, 0, 300, 1152, 1, P, None, None) == BADSHAPE
  # LSTM backward
  assert lib.yt8m_lstm_bwd_workspace_bytes(4, 10, 64, 32, 0) == 0
  assert lib.yt8m_lstm_bwd_workspace_bytes(64, 300, 1152, 1024, 2) > 64 * 300 * 4096 * (4 + 2 + 2)
  assert lib.yt8m_lstm_bwd(P, P, 4, 10, 64, 256, 1, P, P, P, 1.0, P, P, None, None, P, P, P, 1 << 40, None) == BADPTR   # no gradient given
  assert "neither" in err()
  assert lib.yt8m_lstm_bwd(P, P, 4, 10, 60, 256, 1, P, P, P, 1.0, P, P, P, None, P, P, P, 1 << 40, None) == BADSHAPE    # D % 8
  assert lib.yt8m_lstm_bwd(P, P, 4, 10, 64, 256, 1, P, P, P, 1.0, P, P, P, None, P, P, P, 16, None) == BADSHAPE         # workspace
  assert "workspace" in err()
  assert lib.yt8m_lstm_fwd_train(P, P, 4, 10, 64, 256, 1, P, P, 1.0, P, None, None, None, P, 1 << 40, None) == BADPTR   # no sequence buffers
  # attention pooling / context gate / group max backward
  assert lib.yt8m_attn_pool_bwd(P, 8, P, None, 2, 300, 17, 1152, 0, P, P, 17, None, None) == BADSHAPE                  # A > 16
  assert lib.yt8m_attn_pool_bwd(P, 8, P, None, 2, 300, 8, 1150, 0, P, P, 8, None, None) == BADSHAPE                    # F % 8
  assert lib.yt8m_attn_pool_bwd(P, 8, P, None, 2, 300, 8, 1152, 0, P, P, 4, None, None) == BADSHAPE                    # ld_dl < A
  assert lib.yt8m_attn_pool_bwd(P, 8, P, None, 2, 300, 8, 1152, 0, None, P, 8, None, None) == BADPTR
  assert lib.yt8m_context_gate_bwd(P, P, P, None, None, 4, 16, None, None, None, None, 16, None) == BADPTR             # no output
  assert lib.yt8m_context_gate_bwd(P, P, P, None, None, 4, 16, None, None, P, P, 8, None) == BADSHAPE                  # ld_dg < cols
  assert lib.yt8m_group_max_rows_bwd(P, P, 0, 8, 10, P, None) == BADSHAPE
  assert lib.yt8m_group_max_rows_bwd(P, None, 4, 8, 10, P, None) == BADPTR
  assert lib.yt8m_add_inplace(P, None, 10, None) == BADPTR and lib.yt8m_add_inplace(P, P, 0, None) == OK
