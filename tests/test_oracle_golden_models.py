"""The torch-CPU oracle (oracle/yt8m_oracle.py, oracle/model_oracle.py) against golden vectors produced by EXECUTING THE
REFERENCE'S OWN create_model() / calculate_loss() / Dequantize code (oracle/make_model_golden.py: the reference files are
exec'd against a numpy stand-in for the TF / slim ops they call -- see that script for what is the reference's and what
is the shim's).  This pins the restatement's variable names, the class-major / mixture-minor MoE layout, the concat
orders, einsum subscripts, masking / renormalisation and max-over-heads to the reference source; inputs and weights are
regenerated from seeds, the fixture tests/golden/model_golden.json holds only the reference's outputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import make_model_golden as G
from oracle import model_oracle as MO
from oracle import yt8m_oracle as O

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_golden.json")))
TOL = 3e-6


def _t(d):
  return {k: torch.from_numpy(np.asarray(v)) for k, v in d.items()}


def _close(got, want):
  want = torch.tensor(want, dtype=torch.float64)
  assert tuple(got.shape) == tuple(want.shape), (got.shape, want.shape)
  err = float((got.double() - want).abs().max())
  assert err < TOL, err


def test_moe_and_logistic_against_reference_code():
  inp, w, fl = G.case_inputs("moe")
  w = _t(w)
  _close(O.moe_model(torch.from_numpy(inp["x"]), w["gates/weights"], w["experts/weights"], w["experts/biases"], fl["vocab"],
                     fl["moe_num_mixtures"]), GOLDEN["moe"]["predictions"])
  inp, w, fl = G.case_inputs("logistic")
  w = _t(w)
  _close(O.logistic_model(torch.from_numpy(inp["x"]), w["fully_connected/weights"], w["fully_connected/biases"]),
         GOLDEN["logistic"]["predictions"])


def test_chain_models_against_reference_code():
  inp, w, fl = G.case_inputs("chain")
  p, sp = MO.chain_moe(_t(w), torch.from_numpy(inp["x"]), fl["vocab"], fl["moe_num_mixtures"], fl["num_supports"])
  _close(p, GOLDEN["chain"]["predictions"])
  _close(sp, GOLDEN["chain"]["support_predictions"])
  inp, w, fl = G.case_inputs("deep_chain")
  p, sup = MO.deep_combine_chain(_t(w), torch.from_numpy(inp["x"]), fl["vocab"], fl["moe_num_mixtures"], fl["deep_chain_layers"])
  _close(p, GOLDEN["deep_chain"]["predictions"])
  _close(sup, GOLDEN["deep_chain"]["support_predictions"])


def test_cross_entropy_and_dequantize_against_reference_code():
  inp, _, _ = G.case_inputs("xent")
  loss = O.cross_entropy_loss(torch.from_numpy(inp["p"]), torch.from_numpy(inp["labels"]).float())
  assert abs(float(loss) - GOLDEN["xent"]["loss"]) < 2e-5 * GOLDEN["xent"]["loss"]
  inp, _, _ = G.case_inputs("dequantize")
  _close(O.dequantize(torch.from_numpy(inp["u8"])), GOLDEN["dequantize"]["predictions"])


@pytest.mark.parametrize("case", ["lstm_att_max", "lstm_multi_att"])
def test_lstm_attention_models_against_reference_code(case):
  inp, w, fl = G.case_inputs(case)
  fwd = MO.lstm_attention_max_pooling if case == "lstm_att_max" else MO.lstm_multi_attention
  heads = fl["lstm_attentions"] if case == "lstm_att_max" else fl["attention_size"]
  p = fwd(_t(w), torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"]), fl["vocab"], fl["moe_num_mixtures"], heads,
          layers=fl["lstm_layers"])
  _close(p, GOLDEN[case]["predictions"])


@pytest.mark.parametrize("case", ["lstm", "lstm_memory"])
def test_lstm_models_against_reference_code(case):
  """LstmModel feeds the NON-tuple MultiRNNCell state [c0, h0, c1, h1] to the classifier (lstm_model.py:34-52), LstmMemoryModel the
  concatenated c states (lstm_memory_model.py:61)."""
  inp, w, fl = G.case_inputs(case)
  fwd = MO.lstm_model if case == "lstm" else MO.lstm_memory_model
  p = fwd(_t(w), torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"]), fl["vocab"], fl["moe_num_mixtures"], layers=fl["lstm_layers"])
  _close(p, GOLDEN[case]["predictions"])


def test_zt_attention_model_against_reference_code():
  """zt AttentionModel + MoeExtendModel (zt/frame_level_models.py:4355-4405, zt/video_level_models.py:2272-2330): the two classes
  are cut out of the reference files and executed; mask from all-zero rows, mean pooling by num_frames, softmax over T of
  [x_t, mean] . W + b, renormalisation, MoE on B*A rows, max over the A heads."""
  inp, w, fl = G.case_inputs("zt_attention")
  p = MO.attention_model(_t(w), torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"]), fl["vocab"], fl["moe_num_mixtures"],
                         fl["moe_num_extend"])
  _close(p, GOLDEN["zt_attention"]["predictions"])


@pytest.mark.parametrize("case", ["dbof_bn", "dbof_bias"])
def test_dbof_model_against_reference_code(case):
  """DbofModel + model_utils.SampleRandomFrames / FramePooling (wh/all_frame_models/dbof_model.py:62-123, wh/model_utils.py:56-94)
  executed from the reference files; the tf.random_uniform draw is pinned by the harness, so the frame index arithmetic
  (int(uniform * num_frames)) is checked too.  Batch-norm in inference form and the bias form."""
  inp, w, fl = G.case_inputs(case)
  x, nf = torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"])
  fidx = (torch.from_numpy(inp["uniform"]) * nf.float().unsqueeze(1)).to(torch.int64)
  sd = _t(w)
  if case == "dbof_bn":
    p = MO.dbof(sd, x, fidx, fl["vocab"], fl["moe_num_mixtures"], pooling="max")
  else:
    hid = O.dbof_pool(x, fidx, {"cluster_w": sd["cluster_weights"], "cluster_b": sd["cluster_biases"], "hidden_w": sd["hidden1_weights"],
                                "hidden_b": sd["hidden1_biases"]}, add_batch_norm=False, pooling="max")
    p = O.moe_model(hid, sd["gates/weights"], sd["experts/weights"], sd["experts/biases"], fl["vocab"], fl["moe_num_mixtures"])
  _close(p, GOLDEN[case]["predictions"])


def test_frame_reader_video_matrix_against_reference_code():
  """wh/readers.py:159-186 get_video_matrix + :21-56 resize_axis executed from the reference file: decode_raw, Dequantize FIRST,
  then zero padding / truncation to max_frames (padded rows are 0.0, not Dequantize(0)).  Our reader keeps uint8 and
  de-quantises + pads on the GPU; the host half (get_video_matrix -> uint8 [max_frames, D] + n) and the oracle's dequantize
  reproduce the reference matrix exactly."""
  import sys
  pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "youtube-8m_b200")
  if pkg not in sys.path:
    sys.path.insert(0, pkg)
  import readers
  inp, _, fl = G.case_inputs("video_matrix")
  r = readers.YT8MFrameFeatureReader(feature_names=["rgb"], feature_sizes=[fl["feature_size"]], max_frames=fl["max_frames"])
  for key in ("frames_short", "frames_long"):
    u8, n = r.get_video_matrix([row.tobytes() for row in inp[key]], fl["feature_size"])
    want = GOLDEN["video_matrix"][key]
    assert n == want["num_frames"]
    got = O.dequantize(torch.from_numpy(u8.astype(np.float32)))
    got[n:] = 0.0                                                   # what yt8m_l2norm_rows_fwd / yt8m_frames_unpack_u8 write past num_frames
    _close(got, want["matrix"])


def test_inference_csv_lines_against_reference_code():
  """wh/inference.py:76-87 format_lines executed from the reference file: 'id,cls conf cls conf ...' with the top_k classes by
  descending confidence and '%i %f' pairs.  Our inference.format_lines takes the per-video top-k that yt8m_topk_rows
  extracts on the GPU (here: numpy's descending sort) and must emit the same lines."""
  import importlib.util
  import sys
  pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "youtube-8m_b200")
  if pkg not in sys.path:
    sys.path.insert(0, pkg)
  try:
    import inference
  except ImportError as e:                      # imports the CUDA binding (built library required)
    pytest.skip(str(e))
  inp, _, fl = G.case_inputs("format_lines")
  order = np.argsort(-inp["predictions"], axis=1, kind="stable")[:, :fl["top_k"]]
  vals = np.take_along_axis(inp["predictions"], order, axis=1)
  assert list(inference.format_lines(inp["video_ids"], order, vals)) == GOLDEN["format_lines"]["lines"]


def test_epoch_log_line_against_reference_code():
  """wh/utils.py:100-138 AddEpochSummary executed from the reference file (TensorBoard writer stubbed): the
  'epoch/eval number ...' line that eval.py prints, incl. its '{5:3f}' (sic) loss format."""
  import sys
  pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "youtube-8m_b200")
  if pkg not in sys.path:
    sys.path.insert(0, pkg)
  import utils
  inp, _, _ = G.case_inputs("log_lines")
  assert utils.FormatEpochInfo(inp["epoch"]) == GOLDEN["log_lines"]["epoch"]


def test_multitask_loss_and_support_labels_match_the_reference():
  """wh/losses.py:222-279 executed from the reference source (oracle/make_model_golden.py case multitask_xent): the oracle's
  MultiTaskCrossEntropyLoss and the product's host-side support-label construction."""
  import torch
  import losses
  from yt8m_flags import FLAGS
  inp, _, fl = G.case_inputs("multitask_xent")
  gold = GOLDEN["multitask_xent"]
  labels = torch.from_numpy(inp["labels"])
  for st, key in (("label,label", "support_ll"), ("frequent", "support_freq")):
    want_sup = np.asarray(gold["support_labels:" + st], dtype=np.float32)
    sup = O.support_labels(labels, st, num_frequents=fl["num_frequents"])
    assert np.array_equal(sup.numpy(), want_sup)
    got = O.multitask_cross_entropy_loss(torch.from_numpy(inp["p"]), torch.from_numpy(inp[key]), labels, sup, fl["support_loss_percent"])
    assert abs(float(got) - gold[st]) < 2e-5 * gold[st]
    with FLAGS.override(support_type=st, num_frequents=fl["num_frequents"]):
      assert np.array_equal(losses.MultiTaskCrossEntropyLoss().get_support(labels), want_sup)


def test_vertical_support_labels(tmp_path):
  """support_type="vertical" (wh/losses.py:231-247): (labels . mapping) > 0.2 with the 0/1 table read from --vertical_file."""
  import torch
  import losses
  from yt8m_flags import FLAGS
  f = tmp_path / "vertical.tsv"
  f.write_text("0 1\n1 1\n2 0\n3 2\n7\n")
  labels = torch.tensor([[1, 0, 0, 0, 0], [0, 0, 1, 1, 0], [0, 0, 0, 0, 1]], dtype=torch.bool)
  vm = torch.zeros(5, 3)
  for a, b in ((0, 1), (1, 1), (2, 0), (3, 2)):
    vm[a, b] = 1
  want = O.support_labels(labels, "vertical", vertical_mapping=vm).numpy()
  losses.MultiTaskLoss._vertical = None
  with FLAGS.override(support_type="vertical", num_classes=5, num_verticals=3, vertical_file=str(f)):
    got = losses.MultiTaskCrossEntropyLoss().get_support(labels)
  losses.MultiTaskLoss._vertical = None
  assert np.array_equal(got, want) and got.tolist() == [[0, 1, 0], [1, 0, 1], [0, 0, 0]]


def test_cnn_deep_combine_chain_matches_the_reference():
  """wh/all_frame_models/cnn_deep_combine_chain_model.py:10-124 executed from the reference source (shifted-concat temporal CNN,
  unmasked max over time, chain of MoE stages): predictions and support predictions."""
  inp, w, fl = G.case_inputs("cnn_deep_chain")
  sd = {k: torch.from_numpy(v) for k, v in w.items()}
  got, sup = MO.cnn_deep_combine_chain(sd, torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"]), fl["vocab"],
                                       fl["moe_num_mixtures"], fl["deep_chain_layers"])
  want, want_sup = np.asarray(GOLDEN["cnn_deep_chain"]["predictions"]), np.asarray(GOLDEN["cnn_deep_chain"]["support_predictions"])
  assert np.abs(got.numpy() - want).max() < 3e-6 and np.abs(sup.numpy() - want_sup).max() < 3e-6


def test_lstm_parallel_finaloutput_matches_the_reference():
  """wh/all_frame_models/lstm_parallel_finaloutput_model.py:13-73 executed from the reference source: per-modality L2
  normalisation, one LSTM stack per modality under RNN<i>, concatenation of every layer's h state."""
  inp, w, fl = G.case_inputs("lstm_parallel")
  sd = {k: torch.from_numpy(v) for k, v in w.items()}
  sizes = [int(s) for s in fl["feature_sizes"].split(",")]
  got = MO.lstm_parallel_finaloutput(sd, torch.from_numpy(inp["x"]), torch.from_numpy(inp["num_frames"]), fl["vocab"],
                                     fl["moe_num_mixtures"], sizes, fl["lstm_layers"])
  assert np.abs(got.numpy() - np.asarray(GOLDEN["lstm_parallel"]["predictions"])).max() < 3e-6
