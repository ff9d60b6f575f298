"""End-to-end CLI test on the GPU box: train.py -> eval.py -> inference.py with the reference's flags on
synthetic TFRecords (video-level MoeModel), and a frame-level eval / inference pass through the
uint8 -> de-quantise + L2-normalise -> NetVLAD plugin path."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "youtube-8m_b200")
V = 4716


def _run(script, *args):
  r = subprocess.run([sys.executable, os.path.join(PKG, script)] + list(args), capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  return r.stderr + r.stdout


def test_train_eval_inference_video_level(tmp_path):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  sys.path.insert(0, PKG)
  import readers
  rs = np.random.RandomState(0)
  recs = []
  for i in range(96):
    lab = sorted(set(rs.randint(0, 50, size=3).tolist()))
    f = rs.randn(1152).astype(np.float32)
    f[lab] += 6.0                                                     # learnable signal: label index lights up a feature
    recs.append(readers.encode_example({"video_id": ("bytes", [b"vid%03d" % i]), "labels": ("int64", lab),
                                        "mean_rgb": ("float", f[:1024]), "mean_audio": ("float", f[1024:])}))
  data = str(tmp_path / "train.tfrecord")
  readers.write_tfrecord(data, recs)
  train_dir = str(tmp_path / "model")
  common = ["--feature_names=mean_rgb,mean_audio", "--feature_sizes=1024,128", "--model=MoeModel", "--moe_num_mixtures=2"]
  log = _run("train.py", "--train_data_pattern=" + data, "--train_dir=" + train_dir, "--start_new_model", "--batch_size=32",
             "--num_epochs=8", "--base_learning_rate=0.01", *common)
  assert "training step 24|" in log and "Exited training loop." in log
  first = float(log.split("training step 1|")[1].split("Loss: ")[1].split()[0])
  last = float(log.split("training step 24|")[1].split("Loss: ")[1].split()[0])
  assert last < first                                                  # the step actually learns
  assert os.path.exists(os.path.join(train_dir, "model.ckpt-24"))
  # resume: two more epochs continue from step 24
  log2 = _run("train.py", "--train_data_pattern=" + data, "--train_dir=" + train_dir, "--batch_size=32", "--num_epochs=2", *common)
  assert "Restoring from checkpoint" in log2 and "training step 30|" in log2
  ev = _run("eval.py", "--eval_data_pattern=" + data, "--train_dir=" + train_dir, "--run_once", "--batch_size=64", *common)
  assert "epoch/eval number 30 | Avg_Hit@1:" in ev
  out = str(tmp_path / "pred.csv")
  _run("inference.py", "--input_data_pattern=" + data, "--train_dir=" + train_dir, "--output_file=" + out, "--batch_size=50", "--top_k=5",
       *common)
  lines = open(out).read().strip().split("\n")
  assert lines[0] == "VideoId,LabelConfidencePairs" and len(lines) == 97
  vid, pairs = lines[1].split(",")
  toks = pairs.split(" ")
  assert vid == "vid000" and len(toks) == 10
  confs = [float(c) for c in toks[1::2]]
  assert confs == sorted(confs, reverse=True) and all(0 <= c <= 1 for c in confs)
  # pre-ensemble prediction records (wh/inference-pre-ensemble.py:291-308): full vectors, file_size examples per file,
  # batches that do not divide file_size; the top-5 of the CSV are the 5 largest entries of the stored vector
  pdir = str(tmp_path / "pre")
  _run("inference-pre-ensemble.py", "--input_data_pattern=" + data, "--train_dir=" + train_dir, "--output_dir=" + pdir,
       "--batch_size=50", "--file_size=40", *common)
  files = sorted(os.listdir(pdir))
  assert files == ["predictions-0000.tfrecord", "predictions-0001.tfrecord", "predictions-0002.tfrecord"]
  counts = [len(list(readers.tfrecord_iterator(os.path.join(pdir, f), True))) for f in files]
  assert counts == [40, 40, 16]
  ex = readers.parse_example(next(readers.tfrecord_iterator(os.path.join(pdir, files[0]), True)))
  assert ex["video_id"][1][0] == b"vid000" and len(ex["predictions"][1]) == V
  vec = np.asarray(ex["predictions"][1], dtype=np.float32)
  top5 = np.argsort(-vec)[:5]
  assert [int(c) for c in toks[0::2]] == top5.tolist()
  assert np.allclose(vec[top5], confs, atol=1e-6)


def test_eval_frame_level_netvlad(tmp_path):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  sys.path.insert(0, PKG)
  import readers, utils
  rs = np.random.RandomState(1)
  recs = []
  for i in range(6):
    n = int(rs.randint(40, 300))
    recs.append(readers.encode_sequence_example(
        {"video_id": ("bytes", [b"f%d" % i]), "labels": ("int64", [int(rs.randint(0, V))])},
        {"rgb": [("bytes", [rs.randint(0, 256, 1024).astype(np.uint8).tobytes()]) for _ in range(n)],
         "audio": [("bytes", [rs.randint(0, 256, 128).astype(np.uint8).tobytes()]) for _ in range(n)]}))
  data = str(tmp_path / "frames.tfrecord")
  readers.write_tfrecord(data, recs)
  # a checkpoint with the model's freshly initialised variables
  import yt8m_ops as ops, frame_level_models
  from yt8m_flags import FLAGS
  FLAGS.parse([], known_only=True)
  with FLAGS.override(netvlad_cluster_size=64, moe_num_mixtures=2):
    ops.get_store().reset(seed=9)
    x = torch.zeros((2, 300, 1152), dtype=torch.bfloat16, device="cuda")
    frame_level_models.NetVLADModel().create_model(x, vocab_size=V, num_frames=torch.tensor([300, 300], device="cuda"))
    train_dir = str(tmp_path / "m")
    utils.save_checkpoint(train_dir, 5, ops.get_store().state_dict())
  common = ["--frame_features", "--feature_names=rgb,audio", "--feature_sizes=1024,128", "--model=NetVLADModel",
            "--netvlad_cluster_size=64", "--moe_num_mixtures=2", "--train_dir=" + train_dir]
  ev = _run("eval.py", "--eval_data_pattern=" + data, "--run_once", "--batch_size=4", *common)
  assert "epoch/eval number 5 | Avg_Hit@1:" in ev
  out = str(tmp_path / "p.csv")
  _run("inference.py", "--input_data_pattern=" + data, "--output_file=" + out, "--batch_size=4", *common)
  assert len(open(out).read().strip().split("\n")) == 7


def test_train_frame_level_netvlad(tmp_path):
  """train.py --frame_features --model=NetVLADModel: the full CUDA training step (NetVLAD + FC + MoE backward)
  learns a synthetic task from uint8 frame TFRecords, checkpoints, and eval.py reads the checkpoint back."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  sys.path.insert(0, PKG)
  import readers
  rs = np.random.RandomState(2)
  recs = []
  for i in range(32):
    n = int(rs.randint(20, 120))
    lab = int(rs.randint(0, 8))
    base = rs.randint(96, 160, (n, 1152))
    base[:, lab * 16:(lab + 1) * 16] += 90                             # the label lights up a block of features
    fr = np.clip(base, 0, 255).astype(np.uint8)
    recs.append(readers.encode_sequence_example(
        {"video_id": ("bytes", [b"t%d" % i]), "labels": ("int64", [lab])},
        {"rgb": [("bytes", [fr[j, :1024].tobytes()]) for j in range(n)], "audio": [("bytes", [fr[j, 1024:].tobytes()]) for j in range(n)]}))
  data = str(tmp_path / "frames.tfrecord")
  readers.write_tfrecord(data, recs)
  train_dir = str(tmp_path / "m")
  common = ["--frame_features", "--feature_names=rgb,audio", "--feature_sizes=1024,128", "--model=NetVLADModel",
            "--netvlad_cluster_size=64", "--netvlad_hidden_size=256", "--netvlad_add_batch_norm=False", "--moe_num_mixtures=2",
            "--train_dir=" + train_dir]
  log = _run("train.py", "--train_data_pattern=" + data, "--start_new_model", "--batch_size=16", "--num_epochs=10",
             "--base_learning_rate=0.002", *common)
  assert "training step 20|" in log and "Exited training loop." in log
  first = float(log.split("training step 1|")[1].split("Loss: ")[1].split()[0])
  last = float(log.split("training step 20|")[1].split("Loss: ")[1].split()[0])
  assert last < 0.7 * first, (first, last)
  assert os.path.exists(os.path.join(train_dir, "model.ckpt-20"))
  ev = _run("eval.py", "--eval_data_pattern=" + data, "--run_once", "--batch_size=16", *common)
  assert "epoch/eval number 20 | Avg_Hit@1:" in ev


def test_train_frame_level_lstm(tmp_path):
  """train.py --frame_features --model=LstmModel: persistent-recurrence forward + back-propagation through time + MoE
  backward learn a synthetic task from uint8 frame TFRecords; eval.py reads the checkpoint back (BASELINE config 3 at
  reduced width)."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  sys.path.insert(0, PKG)
  import readers
  rs = np.random.RandomState(3)
  recs = []
  for i in range(32):
    n = int(rs.randint(20, 120))
    lab = int(rs.randint(0, 8))
    base = rs.randint(96, 160, (n, 1152))
    base[:, lab * 16:(lab + 1) * 16] += 90
    fr = np.clip(base, 0, 255).astype(np.uint8)
    recs.append(readers.encode_sequence_example(
        {"video_id": ("bytes", [b"t%d" % i]), "labels": ("int64", [lab])},
        {"rgb": [("bytes", [fr[j, :1024].tobytes()]) for j in range(n)], "audio": [("bytes", [fr[j, 1024:].tobytes()]) for j in range(n)]}))
  data = str(tmp_path / "frames.tfrecord")
  readers.write_tfrecord(data, recs)
  train_dir = str(tmp_path / "m")
  common = ["--frame_features", "--feature_names=rgb,audio", "--feature_sizes=1024,128", "--model=LstmModel", "--lstm_cells=256",
            "--lstm_layers=2", "--moe_num_mixtures=2", "--train_dir=" + train_dir]
  log = _run("train.py", "--train_data_pattern=" + data, "--start_new_model", "--batch_size=16", "--num_epochs=10",
             "--base_learning_rate=0.002", *common)
  assert "training step 20|" in log and "Exited training loop." in log
  first = float(log.split("training step 1|")[1].split("Loss: ")[1].split()[0])
  last = float(log.split("training step 20|")[1].split("Loss: ")[1].split()[0])
  assert last < 0.7 * first, (first, last)
  assert os.path.exists(os.path.join(train_dir, "model.ckpt-20"))
  ev = _run("eval.py", "--eval_data_pattern=" + data, "--run_once", "--batch_size=16", *common)
  assert "epoch/eval number 20 | Avg_Hit@1:" in ev
