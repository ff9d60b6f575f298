"""Generate tests/golden/gap_golden.json by running the REFERENCE's own metric code.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

wh/average_precision_calculator.py and wh/mean_average_precision_calculator.py import
unchanged under Python 3; wh/eval_util.py needs a stub for its one unused TensorFlow import
(``from tensorflow.python.platform import gfile``, wh/eval_util.py:19).  Inputs are regenerated
from seeds by ``golden_case`` (legacy numpy RandomState: stable across numpy versions), so the
fixture stores only seeds, shapes and the reference's outputs.
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/youtube-8m-wangheda"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "gap_golden.json")

CASES = [
    # name, seed, batch, classes, labels/video, top_k, quantise (ties)
    ("small_dense", 1, 16, 50, 3.0, 20, 0),
    ("ties", 2, 32, 40, 4.0, 20, 8),          # predictions quantised to 1/8 => many exact ties
    ("yt8m_shape", 3, 64, 4716, 3.4, 20, 0),
    ("topk_gt_classes", 4, 8, 12, 2.0, 20, 0),
    ("k5", 5, 48, 300, 3.0, 5, 0),
    ("empty_label_rows", 6, 12, 60, 2.0, 20, 0),   # rows 3 and 7 carry no label at all (wh/eval_util.py:87-96: PERR counts them as 0)
]
ZERO_ROWS = {"empty_label_rows": [3, 7]}


def golden_case(seed, batch, classes, labels_per_video, quant, zero_rows=()):
  """Deterministic synthetic (predictions, labels); shared with tests/test_gap_oracle.py.
  zero_rows: videos whose labels are all cleared afterwards (label-free videos)."""
  rs = np.random.RandomState(seed)
  labels = (rs.random_sample((batch, classes)) < labels_per_video / classes)
  for b in range(batch):                                  # force >= 1 positive per video
    if not labels[b].any():
      labels[b, rs.randint(classes)] = True
  scores = rs.random_sample((batch, classes)) * 0.6 + labels * rs.random_sample((batch, classes)) * 0.6
  preds = np.clip(scores, 0.0, 1.0).astype(np.float32)
  if quant:
    preds = (np.round(preds * quant) / quant).astype(np.float32)
  labels = labels.astype(np.float32)
  for r in zero_rows:
    labels[r] = 0.0
  return preds, labels


def _load_reference():
  tf = types.ModuleType("tensorflow")
  tfp = types.ModuleType("tensorflow.python")
  tfpp = types.ModuleType("tensorflow.python.platform")
  tfpp.gfile = None
  sys.modules.update({"tensorflow": tf, "tensorflow.python": tfp, "tensorflow.python.platform": tfpp})
  sys.path.insert(0, REF)
  import eval_util  # noqa: the reference's
  import average_precision_calculator as apc
  return eval_util, apc


def main():
  eval_util, apc = _load_reference()
  out = {"generator": "oracle/make_golden.py", "reference": "wh/eval_util.py + wh/average_precision_calculator.py",
         "cases": []}
  for name, seed, b, v, lpv, k, quant in CASES:
    preds, labels = golden_case(seed, b, v, lpv, quant, ZERO_ROWS.get(name, ()))
    rec = {"name": name, "seed": seed, "batch": b, "classes": v, "labels_per_video": lpv,
           "top_k": k, "quant": quant, "zero_rows": ZERO_ROWS.get(name, [])}
    rec["hit_at_one"] = float(eval_util.calculate_hit_at_one(preds, labels))
    rec["perr"] = float(eval_util.calculate_precision_at_equal_recall_rate(preds, labels))
    rec["gap"] = float(eval_util.calculate_gap(preds, labels, top_k=k))
    # streaming: two half-batches through EvaluationMetrics
    em = eval_util.EvaluationMetrics(v, k)
    h = b // 2
    em.accumulate(preds[:h], labels[:h], np.full(h, 1.5))
    em.accumulate(preds[h:], labels[h:], np.full(b - h, 2.5))
    g = em.get()
    rec["stream"] = {"avg_hit_at_one": float(g["avg_hit_at_one"]), "avg_perr": float(g["avg_perr"]),
                     "avg_loss": float(g["avg_loss"]), "gap": float(g["gap"]),
                     "map": float(np.mean(g["aps"]))}
    # plain AP on the first row
    rec["ap_row0"] = float(apc.AveragePrecisionCalculator.ap(preds[0], labels[0]))
    rec["ap_at_7_row0"] = float(apc.AveragePrecisionCalculator.ap_at_n(preds[0], labels[0], n=7))
    out["cases"].append(rec)
    print(name, rec["gap"], rec["hit_at_one"], rec["perr"])
  with open(OUT, "w") as f:
    json.dump(out, f, indent=1)
  print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
  main()
