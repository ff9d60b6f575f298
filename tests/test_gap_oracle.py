"""The metric path is PINNED: oracle/gap_oracle.py and the product's vectorised eval_util.py are checked
against golden vectors produced by the reference's own eval_util.py / average_precision_calculator.py
(oracle/make_golden.py, run in the build container)."""
import json
import os

import numpy as np
import pytest

from oracle import gap_oracle
from oracle.make_golden import golden_case
import eval_util

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gap_golden.json")))
CASES = {c["name"]: c for c in GOLD["cases"]}


def _inputs(c):
  return golden_case(c["seed"], c["batch"], c["classes"], c["labels_per_video"], c["quant"], c.get("zero_rows", ()))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
  c = CASES[name]
  p, y = _inputs(c)
  assert gap_oracle.hit_at_one(p, y) == pytest.approx(c["hit_at_one"], abs=1e-12)
  assert gap_oracle.perr(p, y) == pytest.approx(c["perr"], abs=1e-7)
  # 1e-6: the fixture was generated under numpy 2 (NEP 50), where the reference's `1.0 / numpy.float32`
  # and `ap += ...` stay in float32; the oracle accumulates in float64 like the 2017 numpy did.
  assert gap_oracle.gap(p, y, c["top_k"]) == pytest.approx(c["gap"], abs=1e-6)        # incl. the tie order
  assert gap_oracle.ap_at_n(p[0], y[0]) == pytest.approx(c["ap_row0"], abs=1e-6)
  assert gap_oracle.ap_at_n(p[0], y[0], n=7) == pytest.approx(c["ap_at_7_row0"], abs=1e-6)
  sg = gap_oracle.StreamingGap(c["top_k"])
  h = c["batch"] // 2
  sg.accumulate(p[:h], y[:h], np.full(h, 1.5))
  sg.accumulate(p[h:], y[h:], np.full(c["batch"] - h, 2.5))
  g = sg.get()
  for k in ("avg_hit_at_one", "avg_perr", "avg_loss", "gap"):
    assert g[k] == pytest.approx(c["stream"][k], abs=1e-6), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_eval_util_matches_reference_golden(name):
  c = CASES[name]
  p, y = _inputs(c)
  assert eval_util.calculate_hit_at_one(p, y) == pytest.approx(c["hit_at_one"], abs=1e-12)
  assert eval_util.calculate_precision_at_equal_recall_rate(p, y) == pytest.approx(c["perr"], abs=1e-7)
  # With exact ties the reference's AP depends on its heap-array order; the vectorised implementation
  # pools in video order, so tied cases agree only to the tie ambiguity.  Tie-free cases are exact.
  tol = 2e-2 if c["quant"] else 1e-6
  assert eval_util.calculate_gap(p, y, c["top_k"]) == pytest.approx(c["gap"], abs=tol)
  em = eval_util.EvaluationMetrics(c["classes"], c["top_k"])
  h = c["batch"] // 2
  em.accumulate(p[:h], y[:h], np.full(h, 1.5))
  em.accumulate(p[h:], y[h:], np.full(c["batch"] - h, 2.5))
  g = em.get()
  assert g["avg_hit_at_one"] == pytest.approx(c["stream"]["avg_hit_at_one"], abs=1e-7)
  assert g["avg_perr"] == pytest.approx(c["stream"]["avg_perr"], abs=1e-7)
  assert g["avg_loss"] == pytest.approx(c["stream"]["avg_loss"], abs=1e-7)
  assert g["gap"] == pytest.approx(c["stream"]["gap"], abs=tol)
  assert float(np.mean(g["aps"])) == pytest.approx(c["stream"]["map"], abs=tol)


def test_gap_edge_cases():
  y = np.zeros((3, 10), dtype=np.float32)
  p = np.random.RandomState(0).random_sample((3, 10)).astype(np.float32)
  assert eval_util.calculate_gap(p, y) == 0.0                      # no positives -> 0 (numpos == 0)
  assert gap_oracle.gap(p, y) == 0.0
  y[0, 3] = 1
  p[0, 3] = 2.0                                                   # the single positive ranked first
  assert eval_util.calculate_gap(p, y, top_k=5) == pytest.approx(1.0)
  with pytest.raises(ValueError):
    eval_util.top_k_by_class(p, y, 0)
  with pytest.raises(ValueError):
    eval_util.EvaluationMetrics(10, 20).get()


def test_step_metrics_from_topk_equal_the_full_metrics():
  """eval_util.step_metrics_from_topk (the GPU top-32 path of train.py's log line) == the three calculate_* functions on
  the full predictions, including a video without labels (PERR nan convention) and the fallback signal for a video with
  more positives than extracted entries."""
  import numpy as np
  import eval_util
  rs = np.random.RandomState(7)
  n, v, kp = 40, 300, 32
  pred = rs.rand(n, v).astype(np.float32)
  act = (rs.rand(n, v) < 0.02).astype(np.float32)
  act[3] = 0
  act[5, :20] = 1                                   # 20 positives: still within the 32 extracted entries
  order = np.argsort(-pred, axis=1, kind="stable")[:, :kp]
  tv = np.take_along_axis(pred, order, axis=1)
  h1, perr, gap = eval_util.step_metrics_from_topk(tv, order.astype(np.int32), act, top_k=20)
  assert abs(h1 - eval_util.calculate_hit_at_one(pred, act)) < 1e-6       # float32 vs float64 averaging
  want_perr = eval_util.calculate_precision_at_equal_recall_rate(pred, act)
  assert (np.isnan(perr) and np.isnan(want_perr)) or abs(perr - want_perr) < 1e-6
  assert abs(gap - eval_util.calculate_gap(pred, act, 20)) < 1e-9
  act2 = act.copy()
  act2[3, 0] = 1                                    # no empty video: PERR is a number
  _, perr2, _ = eval_util.step_metrics_from_topk(tv, order, act2, top_k=20)
  assert abs(perr2 - eval_util.calculate_precision_at_equal_recall_rate(pred, act2)) < 1e-6
  act2[7, :40] = 1                                  # 40 positives > 32 extracted: the caller must fall back
  assert eval_util.step_metrics_from_topk(tv, order, act2, top_k=20)[1] is None


def test_label_free_video_counts_zero_in_perr():
  """a video without labels contributes 0 to PERR (the reference's behaviour, golden case empty_label_rows) in
  both the full-prediction function and the GPU top-k path -- never NaN."""
  import eval_util
  c = CASES["empty_label_rows"]
  preds, act = _inputs(c)
  assert abs(eval_util.calculate_precision_at_equal_recall_rate(preds, act) - c["perr"]) < 1e-6
  order = np.argsort(-preds, axis=1, kind="stable")[:, :32]
  tv = np.take_along_axis(preds, order, axis=1)
  h1, perr, gap = eval_util.step_metrics_from_topk(tv, order, act, top_k=20)
  assert abs(perr - c["perr"]) < 1e-6 and abs(h1 - c["hit_at_one"]) < 1e-6 and np.isfinite(gap)
