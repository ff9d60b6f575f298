"""Generate tests/golden/model_golden.json by EXECUTING THE REFERENCE'S OWN model / loss code on small seeded inputs.

Run in the build container only (needs /root/reference):

    python oracle/make_model_golden.py

The reference is Python 2 + TensorFlow 1.0 and neither is available, so its files are read from /root/reference, the py2
`print x` statements are rewritten (they do not touch arithmetic), `xrange` is bound to `range`, and the module is exec'd
against oracle/tf_numpy_shim.py -- a numpy stand-in for the dozen TF / slim ops those files call.  What runs is therefore
the reference's graph-building code itself: its variable names, its reshapes (the class-major / mixture-minor MoE layout),
its concat orders, einsum subscripts, masks, renormalisations and max-over-heads.  Inputs and weights are regenerated from
seeds by `case_inputs` (shared with tests/test_oracle_golden_models.py), so the fixture holds only seeds and outputs.
"""
import json
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402

REF = "/root/reference/youtube-8m-wangheda"
REF_ZT = "/root/reference/youtube-8m-zhangteng"
OUT = os.path.join(HERE, "..", "tests", "golden", "model_golden.json")


def py3ify(src):
  out = []
  for line in src.splitlines():
    m = re.match(r"^(\s*)print (.*)$", line)
    if m and not m.group(2).lstrip().startswith("("):
      line = "%sprint(%s)" % (m.group(1), m.group(2))
    out.append(line)
  return "\n".join(out)


def load(rel, name, extra=None):
  """exec one reference file as module `name` (stubs for its project-local imports come from sys.modules)."""
  mod = types.ModuleType(name)
  mod.__dict__["xrange"] = range
  mod.__dict__["map"] = lambda f, *a: list(map(f, *a))       # py2: map returns a list (lstm_parallel_finaloutput_model.py:34,56)
  mod.__dict__["print"] = lambda *a, **k: None
  if extra:
    mod.__dict__.update(extra)
  src = py3ify(open(os.path.join(REF, rel), errors="ignore").read())
  exec(compile(src, os.path.join(REF, rel), "exec"), mod.__dict__)
  sys.modules[name] = mod
  return mod


def load_function(path, func_name, namespace):
  """exec ONE top-level function of a reference file."""
  import ast
  src = py3ify(open(path, errors="ignore").read())
  node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == func_name)
  ns = dict(namespace)
  exec(compile(ast.get_source_segment(src, node), path, "exec"), ns)
  return ns[func_name]


def load_class(path, class_name, namespace):
  """exec ONE class of a reference file (zt's model files are thousands of lines of other models): the class source is cut
  out with the ast module and run in `namespace`."""
  import ast
  src = py3ify(open(path, errors="ignore").read())
  node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == class_name)
  ns = dict(namespace)
  ns["xrange"], ns["print"] = range, (lambda *a, **k: None)
  exec(compile(ast.get_source_segment(src, node), path, "exec"), ns)
  return ns[class_name]


# ---- seeded inputs, shared with the test ----------------------------------------------------------------------------
def rnd(rs, shape, scale=1.0):
  return (rs.standard_normal(shape) * scale).astype(np.float32)


def case_inputs(case):
  """Returns (inputs dict, weights dict by TF variable name, flags dict) for a named case."""
  rs = np.random.RandomState({"moe": 1, "logistic": 2, "chain": 3, "deep_chain": 4, "xent": 5, "lstm_att_max": 6, "lstm_multi_att": 7,
                              "dequantize": 8, "lstm": 9, "lstm_memory": 10, "zt_attention": 11, "dbof_bn": 12, "dbof_bias": 13, "video_matrix": 14, "format_lines": 15, "log_lines": 16, "multitask_xent": 17, "cnn_deep_chain": 18, "lstm_parallel": 19}[case])
  b, d, v, m = 4, 8, 6, 2
  if case == "moe":
    return ({"x": rnd(rs, (b, d))}, {"gates/weights": rnd(rs, (d, v * (m + 1))), "experts/weights": rnd(rs, (d, v * m)),
                                      "experts/biases": rnd(rs, (v * m,), 0.3)}, {"moe_num_mixtures": m, "vocab": v})
  if case == "logistic":
    return ({"x": rnd(rs, (b, d))}, {"fully_connected/weights": rnd(rs, (d, v)), "fully_connected/biases": rnd(rs, (v,), 0.3)},
            {"vocab": v})
  if case == "chain":
    s = 5
    w = {}
    for scope, din, vv in (("-support", d, s), ("-main", d + s, v)):
      w["gates%s/weights" % scope] = rnd(rs, (din, vv * (m + 1)))
      w["experts%s/weights" % scope] = rnd(rs, (din, vv * m))
      w["experts%s/biases" % scope] = rnd(rs, (vv * m,), 0.3)
    return ({"x": rnd(rs, (b, d))}, w, {"moe_num_mixtures": m, "num_supports": s, "vocab": v})
  if case == "deep_chain":
    layers, r = 2, 4
    w = {}
    for i in range(layers + 1):
      sc = "-prediction-%d" % i if i < layers else "--main"
      din = d + i * r
      w["gates%s/weights" % sc] = rnd(rs, (din, v * (m + 1)))
      w["experts%s/weights" % sc] = rnd(rs, (din, v * m))
      w["experts%s/biases" % sc] = rnd(rs, (v * m,), 0.3)
    for i in range(layers):
      w["relu-%d/weights" % i] = rnd(rs, (v, r))
      w["relu-%d/biases" % i] = rnd(rs, (r,), 0.3)
    return ({"x": rnd(rs, (b, d))}, w, {"moe_num_mixtures": m, "num_supports": 25, "deep_chain_layers": layers, "deep_chain_relu_cells": r,
                                        "deep_chain_relu_type": "relu", "deep_chain_use_length": False, "vocab": v})
  if case == "xent":
    p = 1.0 / (1.0 + np.exp(-rnd(rs, (b, v), 2.0)))
    p[0, 0], p[1, 1] = 0.0, 1.0                     # the epsilon inside the logs matters exactly here
    return ({"p": p.astype(np.float32), "labels": rs.random_sample((b, v)) < 0.4}, {}, {"label_smoothing": False})
  if case == "cnn_deep_chain":
    # wh/all_frame_models/cnn_deep_combine_chain_model.py: 3 videos x 6 frames x 4 features, relu_cells = 3, 2 chain layers
    bb, t, dd, rc, layers = 3, 6, 4, 3, 2
    x = rnd(rs, (bb, t, dd))
    nf = np.array([6, 3, 1], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]
    w = {"mean-relu/weights": rnd(rs, (dd, rc)), "mean-relu/biases": rnd(rs, (rc,), 0.3)}
    for l in range(layers + 1):
      for fs, nfil in ((1, rc), (2, rc), (3, 2 * rc)):
        w["cnn%dcnn-filter-len%d" % (l, fs)] = rnd(rs, (dd * fs, nfil), 0.7)
    din = 4 * rc
    for l in range(layers + 1):
      sc = "prediction-%d" % l if l < layers else "-main"
      w["gates-%s/weights" % sc] = rnd(rs, (din, v * (m + 1)))
      w["experts-%s/weights" % sc] = rnd(rs, (din, v * m))
      w["experts-%s/biases" % sc] = rnd(rs, (v * m,), 0.3)
      if l < layers:
        w["relu-%d/weights" % l] = rnd(rs, (v, rc))
        w["relu-%d/biases" % l] = rnd(rs, (rc,), 0.3)
      din = dd + 4 * rc + (l + 2) * rc              # mean_input + cnn descriptor + the relu layers so far
    return ({"x": x.astype(np.float32), "num_frames": nf}, w,
            {"moe_num_mixtures": m, "num_supports": 25, "deep_chain_layers": layers, "deep_chain_relu_cells": rc, "vocab": v})
  if case == "lstm_parallel":
    # wh/all_frame_models/lstm_parallel_finaloutput_model.py: two modalities (5 + 3 features), LSTM sizes 4 and 2, 2 layers each
    bb, t, sizes, hs, layers = 3, 5, (5, 3), (4, 2), 2
    x = rnd(rs, (bb, t, sum(sizes)))
    nf = np.array([5, 2, 4], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]
    w = {}
    for i, (dsz, h) in enumerate(zip(sizes, hs)):
      for l in range(layers):
        din = dsz if l == 0 else h
        w["RNN%d/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % (i, l)] = rnd(rs, (din + h, 4 * h), 0.5)
        w["RNN%d/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % (i, l)] = rnd(rs, (4 * h,), 0.2)
    feat = layers * sum(hs)
    w["gates/weights"], w["experts/weights"] = rnd(rs, (feat, v * (m + 1)), 0.5), rnd(rs, (feat, v * m), 0.5)
    w["experts/biases"] = rnd(rs, (v * m,), 0.3)
    return ({"x": x.astype(np.float32), "num_frames": nf}, w,
            {"lstm_cells": ",".join(str(h) for h in hs), "lstm_layers": layers, "feature_names": "rgb,audio",
             "feature_sizes": ",".join(str(d_) for d_ in sizes), "moe_num_mixtures": m, "rnn_swap_memory": False,
             "video_level_classifier_model": "MoeModel", "vocab": v})
  if case == "multitask_xent":
    # wh/losses.py:271-279 with --support_type="label,label" (the chain scripts, training_scripts/run-cascade-75-chaining-video.sh:17-20)
    # and with --support_type="frequent"
    sig = lambda a: (1.0 / (1.0 + np.exp(-a))).astype(np.float32)
    return ({"p": sig(rnd(rs, (b, v), 2.0)), "support_ll": sig(rnd(rs, (b, 2 * v), 2.0)), "support_freq": sig(rnd(rs, (b, 3), 2.0)),
             "labels": rs.random_sample((b, v)) < 0.4}, {},
            {"label_smoothing": False, "support_loss_percent": 0.3, "num_frequents": 3, "num_classes": v})
  if case in ("lstm", "lstm_memory"):
    bb, t, dd, h, layers = 3, 6, 4, 5, 2
    x = rnd(rs, (bb, t, dd))
    nf = np.array([6, 3, 1], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]
    w = {}
    for l in range(layers):
      din = dd if l == 0 else h
      w["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % l] = rnd(rs, (din + h, 4 * h), 0.5)
      w["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % l] = rnd(rs, (4 * h,), 0.2)
    feat = layers * 2 * h if case == "lstm" else layers * h          # [c0, h0, c1, h1] vs concat of the c states
    w["gates/weights"], w["experts/weights"] = rnd(rs, (feat, v * (m + 1)), 0.5), rnd(rs, (feat, v * m), 0.5)
    w["experts/biases"] = rnd(rs, (v * m,), 0.3)
    return ({"x": x.astype(np.float32), "num_frames": nf}, w,
            {"lstm_cells": str(h), "lstm_layers": layers, "rnn_swap_memory": False, "video_level_classifier_model": "MoeModel",
             "moe_num_mixtures": m, "vocab": v})
  if case in ("lstm_att_max", "lstm_multi_att"):
    bb, t, dd, h, layers, a = 3, 5, 4, 6, 2, 3
    x = rnd(rs, (bb, t, dd))
    nf = np.array([5, 2, 1], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]
    w = {}
    for l in range(layers):
      din = dd if l == 0 else h
      w["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % l] = rnd(rs, (din + h, 4 * h), 0.5)
      w["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % l] = rnd(rs, (4 * h,), 0.2)
    if case == "lstm_att_max":
      w["attention-/weights"], w["attention-/biases"] = rnd(rs, (dd + h, a)), rnd(rs, (a,), 0.3)
      gn, en, pool = "gates-sub-moe", "experts-sub-moe", h
    else:
      w["fully_connected/weights"], w["fully_connected/biases"] = rnd(rs, (h, a)), rnd(rs, (a,), 0.3)
      gn, en, pool = "gates", "experts", dd
    w[gn + "/weights"], w[en + "/weights"] = rnd(rs, (pool, v * (m + 1))), rnd(rs, (pool, v * m))
    w[en + "/biases"] = rnd(rs, (v * m,), 0.3)
    return ({"x": x.astype(np.float32), "num_frames": nf}, w,
            {"lstm_cells": str(h), "lstm_layers": layers, "lstm_attentions": a, "attention_size": a, "rnn_swap_memory": False,
             "video_level_classifier_model": "MoeModel", "moe_num_mixtures": m, "vocab": v})
  if case in ("dbof_bn", "dbof_bias"):
    bb, t, dd, c, hd, n = 3, 7, 5, 9, 4, 4
    x = rnd(rs, (bb, t, dd))
    nf = np.array([7, 3, 1], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]
    w = {"cluster_weights": rnd(rs, (dd, c)), "hidden1_weights": rnd(rs, (c, hd)), "gates/weights": rnd(rs, (hd, v * (m + 1))),
         "experts/weights": rnd(rs, (hd, v * m)), "experts/biases": rnd(rs, (v * m,), 0.3)}
    if case == "dbof_bn":
      for scope, width in (("input_bn", dd), ("cluster_bn", c), ("hidden1_bn", hd)):
        w[scope + "/gamma"], w[scope + "/beta"] = 1 + rnd(rs, (width,), 0.2), rnd(rs, (width,), 0.2)
        w[scope + "/moving_mean"], w[scope + "/moving_variance"] = rnd(rs, (width,), 0.3), 0.5 + rs.random_sample(width).astype(np.float32)
    else:
      w["cluster_biases"], w["hidden1_biases"] = rnd(rs, (c,), 0.3), rnd(rs, (hd,), 0.3)
    return ({"x": x.astype(np.float32), "num_frames": nf, "uniform": rs.random_sample((bb, n)).astype(np.float32)}, w,
            {"iterations": n, "dbof_add_batch_norm": case == "dbof_bn", "sample_random_frames": True, "dbof_cluster_size": c,
             "dbof_hidden_size": hd, "dbof_pooling_method": "max", "video_level_classifier_model": "MoeModel", "moe_num_mixtures": m,
             "vocab": v})
  if case == "zt_attention":
    bb, t, dd, a = 3, 6, 5, 3
    x = rnd(rs, (bb, t, dd))
    nf = np.array([6, 4, 1], dtype=np.int32)
    x = x * (np.arange(t)[None, :] < nf[:, None])[:, :, None]           # padded frames are all-zero rows: the model's own mask
    w = {"Attention/W": rnd(rs, (2 * dd, a)), "Attention/b": rnd(rs, (a,), 0.3), "gates/weights": rnd(rs, (dd, v * (m + 1))),
         "experts/weights": rnd(rs, (dd, v * m)), "experts/biases": rnd(rs, (v * m,), 0.3)}
    return ({"x": x.astype(np.float32), "num_frames": nf}, w,
            {"moe_num_extend": a, "moe_num_mixtures": m, "video_level_classifier_model": "MoeExtendModel", "vocab": v})
  if case == "log_lines":
    return ({"epoch": {"epoch_id": 1234, "avg_hit_at_one": 0.87654, "avg_perr": 0.71234, "avg_loss": 4.56789123,
                       "aps": [0.5, 0.25, 0.8], "gap": 0.81234},
             "step": {"hit_at_one": 0.9, "perr": 0.75, "loss": 3.25, "examples_per_second": 1234.5678}}, {}, {})
  if case == "format_lines":
    return ({"video_ids": [b"abc", b"vid-2", b"x"], "predictions": rs.random_sample((3, 30)).astype(np.float32)}, {}, {"top_k": 5})
  if case == "video_matrix":
    # two videos: 5 frames (padded to max_frames = 8) and 11 frames (truncated to 8); one byte string per frame
    fs = 6
    return ({"frames_short": rs.randint(0, 256, (5, fs)).astype(np.uint8), "frames_long": rs.randint(0, 256, (11, fs)).astype(np.uint8)},
            {}, {"max_frames": 8, "feature_size": fs})
  if case == "dequantize":
    return ({"u8": np.arange(256, dtype=np.float32)}, {}, {})
  raise KeyError(case)


def run_reference(case):
  inputs, weights, flag_dict = case_inputs(case)
  fv = types.SimpleNamespace(**flag_dict)
  shim.install(fv)
  shim.STORE.clear()
  shim.STORE.update(weights)
  del shim.SCOPES[:]
  for stub in ("utils", "model_utils"):
    sys.modules[stub] = types.ModuleType(stub)
  models = load("models.py", "models")
  v = flag_dict.get("vocab")
  if case == "moe":
    out = load("all_video_models/moe_model.py", "ref_moe").MoeModel().create_model(shim.t(inputs["x"]), v)["predictions"]
  elif case == "logistic":
    out = load("all_video_models/logistic_model.py", "ref_logistic").LogisticModel().create_model(shim.t(inputs["x"]), v)["predictions"]
  elif case == "chain":
    res = load("all_video_models/chain_moe_model.py", "ref_chain").ChainMoeModel().create_model(shim.t(inputs["x"]), v)
    return {"predictions": np.asarray(res["predictions"]).tolist(), "support_predictions": np.asarray(res["support_predictions"]).tolist()}
  elif case == "deep_chain":
    res = load("all_video_models/deep_combine_chain_model.py", "ref_deep").DeepCombineChainModel().create_model(shim.t(inputs["x"]), v)
    return {"predictions": np.asarray(res["predictions"]).tolist(), "support_predictions": np.asarray(res["support_predictions"]).tolist()}
  elif case == "xent":
    losses = load("losses.py", "ref_losses")
    out = losses.CrossEntropyLoss().calculate_loss(shim.t(inputs["p"]), inputs["labels"])
    return {"loss": float(out)}
  elif case == "cnn_deep_chain":
    mod = load("all_frame_models/cnn_deep_combine_chain_model.py", "ref_cnn_deep_chain")
    res = mod.CnnDeepCombineChainModel().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])
    return {"predictions": np.asarray(res["predictions"]).tolist(), "support_predictions": np.asarray(res["support_predictions"]).tolist()}
  elif case == "lstm_parallel":
    vlm = types.ModuleType("video_level_models")
    vlm.MoeModel = load("all_video_models/moe_model.py", "ref_moe").MoeModel
    sys.modules["video_level_models"] = vlm
    sys.modules["utils"].GetListOfFeatureNamesAndSizes = load_function(os.path.join(REF, "utils.py"), "GetListOfFeatureNamesAndSizes",
                                                                        {"logging": types.SimpleNamespace(error=lambda *a, **k: None)})
    mod = load("all_frame_models/lstm_parallel_finaloutput_model.py", "ref_lstm_parallel")
    out = mod.LstmParallelFinaloutputModel().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])["predictions"]
  elif case == "multitask_xent":
    losses = load("losses.py", "ref_losses")
    res = {}
    for st, key in (("label,label", "support_ll"), ("frequent", "support_freq")):
      fv.support_type = st
      res[st] = float(losses.MultiTaskCrossEntropyLoss().calculate_loss(shim.t(inputs["p"]), shim.t(inputs[key]), inputs["labels"]))
      res["support_labels:" + st] = np.asarray(losses.MultiTaskCrossEntropyLoss().get_support(inputs["labels"]), dtype=np.float32).tolist()
    return res
  elif case == "lstm_att_max":
    mod = load("all_frame_models/lstm_attention_max_pooling_model.py", "ref_lstm_att_max")
    out = mod.LstmAttentionMaxPoolingModel().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])["predictions"]
  elif case in ("lstm", "lstm_memory"):
    vlm = types.ModuleType("video_level_models")
    vlm.MoeModel = load("all_video_models/moe_model.py", "ref_moe").MoeModel
    sys.modules["video_level_models"] = vlm
    if case == "lstm":
      cls = load("all_frame_models/lstm_model.py", "ref_lstm").LstmModel
    else:
      cls = load("all_frame_models/lstm_memory_model.py", "ref_lstm_memory").LstmMemoryModel
    out = cls().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])["predictions"]
  elif case == "lstm_multi_att":
    vlm = types.ModuleType("video_level_models")
    vlm.MoeModel = load("all_video_models/moe_model.py", "ref_moe").MoeModel
    sys.modules["video_level_models"] = vlm
    mod = load("all_frame_models/lstm_multi_attention_model.py", "ref_lstm_multi_att")
    out = mod.LstmMultiAttentionModel().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])["predictions"]
  elif case in ("dbof_bn", "dbof_bias"):
    vlm = types.ModuleType("video_level_models")
    vlm.MoeModel = load("all_video_models/moe_model.py", "ref_moe").MoeModel
    sys.modules["video_level_models"] = vlm
    load("model_utils.py", "model_utils")
    shim.STORE["__random_uniform__"] = inputs["uniform"]
    unnamed = [weights["cluster_weights"]] + ([weights["cluster_biases"]] if case == "dbof_bias" else []) + [weights["hidden1_weights"]] + \
              ([weights["hidden1_biases"]] if case == "dbof_bias" else [])
    shim.STORE["__unnamed__"] = list(unnamed)                      # tf.Variable(...) without a name: creation order
    cls = load("all_frame_models/dbof_model.py", "ref_dbof").DbofModel
    out = cls().create_model(shim.t(inputs["x"]), v, inputs["num_frames"], is_training=False)["predictions"]
  elif case == "zt_attention":
    tf, slim = sys.modules["tensorflow"], sys.modules["tensorflow.contrib.slim"]
    base = {"tf": tf, "slim": slim, "models": models, "FLAGS": fv, "np": np}
    vlm = types.ModuleType("video_level_models")
    vlm.MoeExtendModel = load_class(os.path.join(REF_ZT, "video_level_models.py"), "MoeExtendModel", base)
    cls = load_class(os.path.join(REF_ZT, "frame_level_models.py"), "AttentionModel", dict(base, video_level_models=vlm))
    out = cls().create_model(shim.t(inputs["x"]), v, inputs["num_frames"])["predictions"]
  elif case == "log_lines":
    class _Writer(object):
      def add_summary(self, *a, **k):
        pass

      def flush(self):
        pass
    ns = {"numpy": np, "MakeSummary": lambda name, value: None}
    epoch = load_function(os.path.join(REF, "utils.py"), "AddEpochSummary", ns)(_Writer(), 1234, inputs["epoch"])
    step = load_function(os.path.join(REF, "utils.py"), "AddGlobalStepSummary", ns)(_Writer(), 77, inputs["step"])
    return {"epoch": epoch, "step": step}
  elif case == "format_lines":
    fn = load_function(os.path.join(REF, "inference.py"), "format_lines", {"numpy": np})
    return {"lines": list(fn(inputs["video_ids"], inputs["predictions"], flag_dict["top_k"]))}
  elif case == "video_matrix":
    sys.modules["utils"] = load("utils.py", "ref_utils")
    readers = load("readers.py", "ref_readers")
    res = {}
    for key in ("frames_short", "frames_long"):
      reader = readers.YT8MFrameFeatureReader.__new__(readers.YT8MFrameFeatureReader)
      mat, n = reader.get_video_matrix([row.tobytes() for row in inputs[key]], flag_dict["feature_size"], flag_dict["max_frames"], 2, -2)
      res[key] = {"matrix": np.asarray(mat, dtype=np.float64).tolist(), "num_frames": int(n)}
    return res
  elif case == "dequantize":
    utils = load("utils.py", "ref_utils", extra={})
    out = utils.Dequantize(inputs["u8"], 2, -2)
  return {"predictions": np.asarray(out, dtype=np.float64).tolist()}


CASES = ["moe", "logistic", "chain", "deep_chain", "xent", "multitask_xent", "cnn_deep_chain", "lstm_parallel", "lstm", "lstm_memory", "lstm_att_max", "lstm_multi_att", "zt_attention",
         "dbof_bn", "dbof_bias", "video_matrix", "format_lines", "log_lines", "dequantize"]


def main():
  golden = {c: run_reference(c) for c in CASES}
  json.dump(golden, open(OUT, "w"), indent=0, sort_keys=True)
  for c in CASES:
    print(c, {k: (np.asarray(val).shape if isinstance(val, list) else val if isinstance(val, float) else "...") for k, val in golden[c].items()})


if __name__ == "__main__":
  main()
