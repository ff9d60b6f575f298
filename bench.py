#!/usr/bin/env python
"""Benchmark of the yt8m_b200 hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # BASELINE.json configs[1] (default, --config 2)
    python bench.py --config 3|4|5 ...                             # the other BASELINE.json frame-level configs
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU restatement of the reference path, host cores

Workloads (numbering = BASELINE.json `configs`, 1-based):
  2  NetVLAD K=64 over 300x1152 frame features -> FC 73,728->1024 (+BN, ReLU6) -> MoE-2 head, batch 256 per GPU (default)
  3  2-layer LSTM-1024 over 300x1152 -> MoE-4 on the 4096-d state, batch 64 per GPU
  4  Gated NetVLAD K=128 -> FC 147,456->1024 -> context gating -> MoE-4, GLOBAL batch 512
  5  8-head attention pooling over 300x1152 -> chained MoE (DeepCombineChainModel, 3 layers, MoE-4), batch 256 per GPU
     (--batch-sweep adds B in {64,128,256,512,1024})
One "step" = one forward pass of the plugin (create_model) over one batch of synthetic frame features.  The path shards by
video: each rank processes its own batch, no data-path collective.  `value` = videos/s with inputs resident in HBM;
`e2e` = the same through the reference-facing plugin call chain (DefaultTransformer.transform + create_model) from pinned
HOST uint8 features with the predictions copied back to the host, every step.  `train_step` = the full optimiser step of the
same workload (forward + backward + clip + Adam; under torchrun ONE NCCL all-reduce of the flat gradient) -- the part of
the path `north_star` partitions across GPUs.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

T, D, V = 300, 1152, 4716

# name -> what the plugin, the oracle and the trainer need
CONFIGS = {
    2: {"workload": "NetVLAD K=64 over 300x1152 frame feats + FC 73728->1024 + MoE-2 head (4716 labels), forward pass",
        "model": "NetVLADModel", "batch": 256, "scaling": "weak",
        "flags": {"netvlad_cluster_size": 64, "netvlad_hidden_size": 1024, "moe_num_mixtures": 2,
                  "video_level_classifier_model": "MoeModel"},
        "tag": "netvlad", "bound": "hbm", "kernel": "netvlad one-pass kernel (K=64)", "cpu_sample": 256},
    3: {"workload": "2-layer LSTM-1024 over 300x1152 frame feats + MoE-4 head on the 4096-d state (4716 labels), forward pass",
        "model": "LstmModel", "batch": 64, "scaling": "weak",
        "flags": {"lstm_cells": "1024", "lstm_layers": 2, "moe_num_mixtures": 4, "video_level_classifier_model": "MoeModel"},
        "tag": "lstm_fwd", "bound": "tensor", "kernel": "yt8m_lstm_fwd (input-projection GEMMs + persistent recurrence)",
        "cpu_sample": 16},
    4: {"workload": "Gated NetVLAD K=128 over 300x1152 frame feats + FC 147456->1024 + context gating + MoE-4 head, forward pass",
        "model": "GatedNetVLADModel", "batch": 512, "scaling": "strong",
        "flags": {"netvlad_cluster_size": 128, "netvlad_hidden_size": 1024, "moe_num_mixtures": 4,
                  "video_level_classifier_model": "MoeModel"},
        "tag": "netvlad", "bound": "hbm", "kernel": "netvlad kernel (K=128)", "cpu_sample": 64},
    5: {"workload": "8-head attention pooling over 300x1152 frame feats + chained MoE (DeepCombineChainModel, 3 layers, MoE-4) "
                    "per head, max over heads, forward pass",
        "model": "AttentionModel", "batch": 256, "scaling": "weak",
        "flags": {"moe_num_mixtures": 4, "moe_num_extend": 8, "video_level_classifier_model": "DeepCombineChainModel",
                  "deep_chain_layers": 3, "deep_chain_relu_cells": 256},
        "tag": "attention_pool", "bound": "hbm", "kernel": "attention pooling (logits + masked softmax over T + weighted sum)",
        "cpu_sample": 32},
}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based)")
  ap.add_argument("--batch", type=int, default=0, help="videos per GPU per step (0 = the config's own)")
  ap.add_argument("--batch-sweep", action="store_true", help="config 5: also time B in {64,128,256,512,1024} (resident inputs)")
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--cpu-sample", type=int, default=0, help="videos per CPU-baseline forward (0 = the config's own)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--dp-gradient-dtype", default="float32", choices=["float32", "bfloat16"],
                  help="wire format of the train step's gradient exchange (train.py --dp_gradient_dtype); float32 = the reference's arithmetic")
  ap.add_argument("--train-steps", type=int, default=8, help="timed steps of the training-step measurement (0 = skip)")
  ap.add_argument("--graphs", type=int, default=4, help="captured copies of the step, each with ITS OWN input batch, rotated through the timed region")
  ap.add_argument("--operand-format", default="f16", choices=["f16", "bf16x2"],
                  help="how the descriptor / hidden layer travel between kernels (see --netvlad_operand_format)")
  return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------

class ClockSampler(object):
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
    mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU restatement (oracle) timing -- cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------

def oracle_state(cfg_id, g):
  """Seed-fixed weights of the workload under the reference's variable names (TF layouts), bf16-representable."""
  import synth
  f = CONFIGS[cfg_id]["flags"]
  m = f["moe_num_mixtures"]
  sd = {}

  def moe(prefix_g, prefix_e, d_in):
    sd[prefix_g + "/weights"] = synth.xavier((d_in, V * (m + 1)), g)
    sd[prefix_e + "/weights"] = synth.xavier((d_in, V * m), g)
    sd[prefix_e + "/biases"] = torch.zeros(V * m)

  if cfg_id in (2, 4):
    k, h = f["netvlad_cluster_size"], f["netvlad_hidden_size"]
    sd["cluster_weights"] = synth.normal((D, k), g, 1 / math.sqrt(D))
    sd["cluster_weights2"] = torch.randn(D, k, generator=g) / math.sqrt(D)
    sd["hidden1_weights"] = synth.normal((k * D, h), g, 1 / math.sqrt(k))
    scopes = [("cluster_bn", k), ("hidden1_bn", h)]
    if cfg_id == 4:
      sd["gating_weights"] = synth.normal((h, h), g, 1 / math.sqrt(h))
      scopes.append(("gating_bn", h))
    for scope, c in scopes:
      sd[scope + "/gamma"], sd[scope + "/beta"] = torch.ones(c), torch.zeros(c)
      sd[scope + "/moving_mean"], sd[scope + "/moving_variance"] = torch.zeros(c), torch.ones(c)
    moe("gates", "experts", h)
  elif cfg_id == 3:
    h, layers = int(f["lstm_cells"]), f["lstm_layers"]
    for l in range(layers):
      in_dim = D if l == 0 else h
      sd["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/weights" % l] = synth.xavier((in_dim + h, 4 * h), g)
      sd["RNN/multi_rnn_cell/cell_%d/basic_lstm_cell/biases" % l] = torch.zeros(4 * h)
    moe("gates", "experts", layers * 2 * h)
  else:
    a, layers, cells = f["moe_num_extend"], f["deep_chain_layers"], f["deep_chain_relu_cells"]
    sd["Attention/W"] = synth.normal((2 * D, a), g, 0.1)
    sd["Attention/b"] = torch.full((a,), 0.1)
    d_in = D
    for i in range(layers):
      moe("gates-prediction-%d" % i, "experts-prediction-%d" % i, d_in)
      sd["relu-%d/weights" % i] = synth.xavier((V, cells), g)
      sd["relu-%d/biases" % i] = torch.zeros(cells)
      d_in += cells
    moe("gates--main", "experts--main", d_in)
  return sd


def cpu_forward_fn(cfg_id, sample, seed=8):
  """Returns (fn, n_videos): fn() runs the oracle forward of the workload on `sample` videos."""
  import synth
  from oracle import model_oracle
  f = CONFIGS[cfg_id]["flags"]
  m = f["moe_num_mixtures"]
  g = torch.Generator().manual_seed(9)
  x, nf, _ = synth.model_input(sample, T, D, seed=seed)
  sd = oracle_state(cfg_id, g)
  if cfg_id in (2, 4):
    fn = lambda: model_oracle.netvlad(sd, x, nf, V, m, gating=(cfg_id == 4))
  elif cfg_id == 3:
    fn = lambda: model_oracle.lstm_model(sd, x, nf, V, m, layers=f["lstm_layers"])
  else:
    fn = lambda: model_oracle.attention_chain(sd, x, nf, V, m, f["moe_num_extend"], f["deep_chain_layers"])
  return fn, sample


def pick_threads(fn):
  """All host threads is not the fastest setting for these small fp32 GEMMs: try a few counts, keep the best."""
  best, best_t = None, None
  for t in sorted({os.cpu_count(), min(os.cpu_count(), 32), min(os.cpu_count(), 16), min(os.cpu_count(), 8)}, reverse=True):
    torch.set_num_threads(t)
    fn()
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    if best_t is None or dt < best_t:
      best, best_t = t, dt
  torch.set_num_threads(best)
  return best


def time_cpu(cfg_id, sample, batch, min_seconds=10.0, max_reps=100000):
  fn, n = cpu_forward_fn(cfg_id, sample)
  threads = pick_threads(fn)             # also the warm-up (pages in the fp32 weights)
  reps, t0 = 0, time.perf_counter()
  while reps < max_reps and (reps < 2 or time.perf_counter() - t0 < min_seconds):
    fn()
    reps += 1
  dt = time.perf_counter() - t0
  return {"value": n * reps / dt, "unit": "videos/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
          "same_batch_as_gpu": n == batch,
          "sample": "%d forwards of %d videos (fp32 torch-CPU restatement of the reference ops, %.1f s; %d threads chosen "
                    "out of %d host cores)" % (reps, n, dt, threads, os.cpu_count())}


def config1_pair(dev, seconds=3.0):
  """BASELINE.json configs[0]: LogisticModel on mean-pooled 1152-d features, batch 128, one FULL train.py step including
  the per-step Hit@1 / PERR / GAP of the log line (wh/train.py:578-591) -- the reference's own CPU-runnable case.  Timed
  on the host cores with the oracle port (torch autograd + oracle clip / TF-Adam + eval_util on the full predictions) and
  on the GPU through HeadTrainer (+ the top-32 metrics path of our train.py).  A side measurement: never fatal."""
  try:
    import synth
    import eval_util
    import yt8m_native as nat
    import yt8m_trainer
    from oracle import yt8m_oracle as O
    b1 = 128
    g = torch.Generator().manual_seed(8)
    x = torch.randn(b1, D, generator=g)
    x = synth.bf16r(x * torch.rsqrt((x * x).sum(1, keepdim=True)))
    y = synth.labels(b1, V, seed=8)
    w0, b0 = synth.xavier((D, V), g), torch.zeros(V)
    lv = y.numpy()
    params = [w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)]
    m, v = [torch.zeros_like(t) for t in params], [torch.zeros_like(t) for t in params]

    def cpu_step(step):
      p = O.logistic_model(x, params[0], params[1])
      loss = O.cross_entropy_loss(p, y) + O.l2_regularizer(params[0], 1e-8)
      grads = torch.autograd.grad(loss, params)
      lr = O.exponential_decay(0.01, step, b1, 4000000, 0.95)
      with torch.no_grad():
        for i, (t, gr) in enumerate(zip(params, grads)):
          new, m[i], v[i] = O.adam_step(t, O.clip_by_norm(gr, 1.0), m[i], v[i], step + 1, lr)
          t.copy_(new)
      pv = p.detach().numpy()
      return eval_util.calculate_hit_at_one(pv, lv), eval_util.calculate_precision_at_equal_recall_rate(pv, lv), eval_util.calculate_gap(pv, lv)

    cpu_step(0)
    n, t0 = 0, time.perf_counter()
    while n < 3 or time.perf_counter() - t0 < seconds:
      n += 1
      cpu_step(n)
    cpu = b1 * n / (time.perf_counter() - t0)

    tr = yt8m_trainer.HeadTrainer("logistic", D, V, device=dev)
    tr.import_state({"fully_connected/weights": w0.to(dev), "fully_connected/biases": b0.to(dev)})
    xd, yd = x.to(dev).to(torch.bfloat16), y.to(dev)

    def gpu_step():
      p = tr.step(xd, yd)
      ti, tv = nat.topk_rows(p, 32)
      return eval_util.step_metrics_from_topk(tv.cpu().numpy(), ti.cpu().numpy(), lv, 20)

    for _ in range(3):
      gpu_step()
    torch.cuda.synchronize()
    k, t0 = 50, time.perf_counter()
    for _ in range(k):
      gpu_step()
    torch.cuda.synchronize()
    gpu = b1 * k / (time.perf_counter() - t0)
    return {"workload": "LogisticModel video-level, batch 128, full train step + per-step Hit@1/PERR/GAP (BASELINE configs[0])",
            "gpu": {"value": gpu, "unit": "videos/s", "steps": k, "timing": "wall clock incl. the host metrics"},
            "cpu_port": {"value": cpu, "unit": "videos/s", "steps": n, "cores": torch.get_num_threads(), "kind": "port"}}
  except Exception as e:                                     # side measurement only
    return {"error": "%s: %s" % (type(e).__name__, e)}


def per_gpu_batch(args, world):
  cfg = CONFIGS[args.config]
  if args.batch:
    return args.batch
  return cfg["batch"] // world if cfg["scaling"] == "strong" else cfg["batch"]


def run_reference(args, rank):
  """--impl reference: the reference's own path is TF-1.0 / python2 and cannot run; the arm
  times the oracle port (kind "port") on the host cores, rank 0 only.  Each step is one forward of the config's CPU sample
  (config 2: the full 256-video batch of the GPU arm; the LSTM / K=128 / attention configs: a bounded sample, stated)."""
  if rank != 0:
    return
  cfg = CONFIGS[args.config]
  sample = args.cpu_sample or cfg["cpu_sample"]
  fn, n = cpu_forward_fn(args.config, sample)
  threads = pick_threads(fn)
  for _ in range(max(1, min(args.warmup, 2))):
    fn()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    fn()
  dt = time.perf_counter() - t0
  val = n * args.steps / dt
  line = {"impl": "reference", "metric": "videos/sec", "value": val, "unit": "videos/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": cfg["workload"], "baseline_config": args.config, "batch_per_step": n, "frames": T, "feature_dim": D,
                     "batch_per_gpu": per_gpu_batch(args, 1)},
          "cpu_baseline": {"value": val, "unit": "videos/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                           "sample": "each step = one forward of %d videos (%d threads chosen out of %d host cores)" %
                                     (n, threads, os.cpu_count())},
          "e2e": {"value": val, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def make_trainer(cfg_id, dev, store_vars):
  """The optimiser step of the workload (yt8m_trainer), initialised from the plugin's own variables."""
  import yt8m_trainer
  f = CONFIGS[cfg_id]["flags"]
  m = f["moe_num_mixtures"]
  if cfg_id in (2, 4):
    tr = yt8m_trainer.NetVLADTrainer(D, clusters=f["netvlad_cluster_size"], hidden=f["netvlad_hidden_size"], vocab=V, mixtures=m,
                                     device=dev, gating=(cfg_id == 4))
    what = ("NetVLAD K=%d + FC + %sMoE-%d forward, full backward, per-tensor clip + Adam" %
            (f["netvlad_cluster_size"], "context gating + " if cfg_id == 4 else "", m))
  elif cfg_id == 3:
    tr = yt8m_trainer.LstmTrainer(D, hidden=int(f["lstm_cells"]), layers=f["lstm_layers"], vocab=V, mixtures=m, device=dev)
    what = "2 x LSTM-1024 + MoE-%d forward, BPTT, per-tensor clip + Adam" % m
  else:
    tr = yt8m_trainer.AttentionTrainer(D, heads=f["moe_num_extend"], vocab=V, mixtures=m, device=dev)
    what = ("8-head attention pooling + MoE-%d per head + max over heads (AttentionModel + MoeExtendModel: the chained head has no "
            "fused trainer over B*A rows yet) forward, full backward, per-tensor clip + Adam" % m)
  tr.import_state(store_vars)
  return tr, what


def main():
  args = parse()
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if args.impl == "reference":
    run_reference(args, rank)
    return
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device; the yt8m_b200 path has no CPU fallback")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  import torch.distributed as dist
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  import yt8m_native as nat
  import yt8m_ops as ops
  import frame_level_models
  import feature_transform
  import readers
  import synth
  from yt8m_flags import FLAGS

  if os.environ.get("YT8M_DEBUG_FLAGS"):          # A/B switches of the kernels (tools/); unset in normal runs
    nat.debug_set_flags(int(os.environ["YT8M_DEBUG_FLAGS"]))
  cfg = CONFIGS[args.config]
  B = per_gpu_batch(args, world)
  FLAGS.parse([], known_only=True)
  for k, v in cfg["flags"].items():
    setattr(FLAGS, k, v)
  FLAGS.netvlad_operand_format = args.operand_format
  ops.get_store().reset(seed=9)
  model = getattr(frame_level_models, cfg["model"])()
  transformer = feature_transform.DefaultTransformer()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t)
    return ms

  # ---- resident inputs: one DISTINCT batch per captured graph (seeds differ), so that consecutive timed steps never read
  # the same frames (what one step reads cannot still sit in L2 from the step before)
  n_graphs = max(args.graphs, 3)
  u8_0, nf_0 = synth.frames_u8(B, T, D, seed=8 + rank)
  resident = []
  for i in range(n_graphs):
    u8_i, nf_i = (u8_0, nf_0) if i == 0 else synth.frames_u8(B, T, D, seed=108 + 16 * i + rank)
    nf_d = nf_i.to(dev)
    x_d, _ = transformer.transform(u8_i.to(dev), nf_d)        # resident bf16, L2-normalised rows
    resident.append((x_d, nf_d, nf_i))
  torch.cuda.synchronize()

  def step_resident(i=0):
    x_d, nf_d, _ = resident[i]
    return model.create_model(x_d, vocab_size=V, num_frames=nf_d)["predictions"]

  # ---- end-to-end: every step copies ITS batch host->device and its predictions device->host inside the timed region.
  # The host batch is what readers.YT8MFrameFeatureReader(packed=True) yields: a readers.PackedFrames holding only the
  # REAL frames of every video (uint8, as stored in the TFRecords); the zero padding up to 300 frames never crosses PCIe.
  # Like the reference's queue-runner input pipeline (wh/train.py:199-209) the next batch's upload is prefetched: a
  # copy stream fills the other of two device buffers while the compute stream works on the current one.
  packed_host = readers.PackedFrames.from_padded(u8_0, nf_0).pin_memory()
  pred_host = torch.empty((B, V), dtype=torch.float32).pin_memory()
  copy_stream = torch.cuda.Stream()
  bufs = [readers.PackedFrames(torch.empty_like(packed_host.data, device=dev), torch.empty_like(packed_host.num_frames, device=dev), T,
                               torch.empty_like(packed_host.offsets, device=dev)) for _ in range(2)]
  ready = [torch.cuda.Event(), torch.cuda.Event()]
  consumed = [torch.cuda.Event(), torch.cuda.Event()]
  state = {"i": 0, "primed": False}

  def upload(slot):
    with torch.cuda.stream(copy_stream):
      copy_stream.wait_event(consumed[slot])                 # the compute stream is done reading this buffer
      bufs[slot].data.copy_(packed_host.data, non_blocking=True)
      bufs[slot].num_frames.copy_(packed_host.num_frames, non_blocking=True)
      bufs[slot].offsets.copy_(packed_host.offsets, non_blocking=True)
      ready[slot].record(copy_stream)

  def step_e2e():
    cur = state["i"] & 1
    if not state["primed"]:
      for sl in (0, 1):
        consumed[sl].record(torch.cuda.current_stream())
      upload(cur)
      state["primed"] = True
    upload(cur ^ 1)                                          # prefetch the next step's batch
    torch.cuda.current_stream().wait_event(ready[cur])
    xi, _ = transformer.transform(bufs[cur], bufs[cur].num_frames)
    p = model.create_model(xi, vocab_size=V, num_frames=bufs[cur].num_frames)["predictions"]
    consumed[cur].record(torch.cuda.current_stream())
    pred_host.copy_(p, non_blocking=True)
    torch.cuda.current_stream().synchronize()                # the caller holds this step's predictions on the host
    state["i"] += 1
    return pred_host

  W = max(args.warmup, 3)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  # The resident step is captured ONCE per input batch into CUDA graphs and replayed (ops.CapturedStep): the host side of a
  # step (Python, ctypes, tensor-map encoding, output allocation) would otherwise outlast the GPU work.  Each capture
  # holds its own CUDA-event pair around the dominant kernel; the pairs are read after the timed region (durations of the
  # last timed replays -- measured live, inside the region).
  step_resident(0)                      # lazy weight packing happens here, outside the launch count
  l0 = nat.launch_count()
  step_resident(0)
  launches = nat.launch_count() - l0
  launch_mode = "step captured once per input batch as a CUDA graph (%d kernels) and replayed; %d captures with distinct inputs rotated"
  try:
    graphs = [ops.CapturedStep(lambda i=i: step_resident(i), warmup=1, time_tag=cfg["tag"]) for i in range(n_graphs)]
    launch_mode = launch_mode % (launches, n_graphs)
  except Exception as e:                 # a launch the driver cannot capture: time eager steps instead, and say so
    torch.cuda.synchronize()
    graphs = None
    launch_mode = "eager launches (%d kernels per step; graph capture refused: %s)" % (launches, type(e).__name__)
  it = {"i": 0}

  def step_graph():
    i = it["i"] % n_graphs
    it["i"] += 1
    return graphs[i]() if graphs is not None else step_resident(i)

  ms = timed(step_graph, args.steps, W)
  if graphs is not None:
    kt = [t for g in graphs[:min(n_graphs, args.steps)] for t in g.kernel_ms()]
  else:
    nat.kernel_timer_begin(cfg["tag"])
    for i in range(n_graphs):
      step_resident(i)
    kt = nat.kernel_timer_end()
  ms_e2e = timed(step_e2e, args.steps, W)

  sweep = None
  if args.batch_sweep:
    sweep = []
    for bs in (64, 128, 256, 512, 1024):
      u8_s, nf_s = synth.frames_u8(bs, T, D, seed=208 + rank)
      nfd = nf_s.to(dev)
      xs, _ = transformer.transform(u8_s.to(dev), nfd)
      fn = lambda: model.create_model(xs, vocab_size=V, num_frames=nfd)["predictions"]
      try:
        g = ops.CapturedStep(fn, warmup=2, time_tag=cfg["tag"])
        ms_s = timed(g, max(args.steps // 2, 5), 3)
        k_s = g.kernel_ms()
      except Exception:
        torch.cuda.synchronize()
        ms_s, k_s = timed(fn, max(args.steps // 2, 5), 3), []
      real = int(nf_s.clamp(0, T).sum())
      sweep.append({"batch_per_gpu": bs, "value": world * bs / (ms_s * 1e-3), "ms_per_step": ms_s,
                    "kernel_ms": (sum(k_s) / len(k_s)) if k_s else None, "real_frame_rows": real})
      del xs
  clocks = sampler.stop() if rank == 0 else None

  # ---- the training step of the same workload: forward + full backward + per-tensor clip + Adam; under torchrun the flat
  # gradient is summed over the ranks once per step (NCCL); inputs resident.  This is the part of the path that communicates.
  ms_train, train_what, grad_floats = None, None, None
  if args.train_steps > 0:
    bn_flag = FLAGS.netvlad_add_batch_norm
    FLAGS.netvlad_add_batch_norm = False                      # the trainer's variables (bias form) come from the plugin itself
    if args.config == 5:
      FLAGS.video_level_classifier_model = "MoeExtendModel"
    ops.get_store().reset(seed=9)
    getattr(frame_level_models, cfg["model"])().create_model(resident[0][0][:2], vocab_size=V, num_frames=resident[0][1][:2])
    tr, train_what = make_trainer(args.config, dev, {k: v.value for k, v in ops.get_store().vars.items()})
    ops.get_store().reset(seed=9)
    FLAGS.netvlad_add_batch_norm = bn_flag
    grad_floats = int(tr.grad.numel())
    if args.dp_gradient_dtype == "bfloat16":
      tr.wire_dtype = torch.bfloat16
    y_dev = synth.labels(B, V, seed=8 + rank).to(dev)
    x_d, nf_d, _ = resident[0]
    ms_train = timed(lambda: tr.step(x_d, nf_d, y_dev, global_batch=world * B), args.train_steps, 3)
    del tr

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
  tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1409.0)))
  peak_src = "MEASURED_PEAKS.json (measured)" if peaks else "fallback"
  fmt = FLAGS.netvlad_operand_format
  k_ms = sum(kt) / len(kt) if kt else None
  used = resident[:min(n_graphs, args.steps)]
  real_rows = sum(int(r[2].clamp(0, T).sum()) for r in used) / float(len(used))      # per launch, averaged like k_ms
  roof = {"kernel": cfg["kernel"], "kernel_ms": k_ms, "peak_source": peak_src}
  if args.config in (2, 4):
    # HBM bound.  Algorithmic bytes per launch: the REAL frames once (padded tiles are not read) + the descriptor once
    # (one fp16 tensor; x2 for a bf16 hi + lo pair) -- SURVEY.md §8(d); achieved_nominal counts every
    # video as 300 frames (691,200 B per video).
    kc = cfg["flags"]["netvlad_cluster_size"]
    out_bytes = B * D * kc * 2 * (1 if fmt == "f16" else 2)
    alg = real_rows * D * 2 + out_bytes
    roof.update({"bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "algorithmic_bytes": alg,
                 "achieved": alg / (k_ms * 1e-3) / 1e9 if k_ms else None,
                 "achieved_nominal": (B * T * D * 2 + out_bytes) / (k_ms * 1e-3) / 1e9 if k_ms else None,
                 "bytes_note": "real frames only (%d of %d frame rows on average; padded tiles are not read) + %s descriptors" %
                               (real_rows, B * T, "fp16" if fmt == "f16" else "bf16 hi+lo")})
  elif args.config == 3:
    h, layers = int(cfg["flags"]["lstm_cells"]), cfg["flags"]["lstm_layers"]
    per_frame = sum(2 * ((D if l == 0 else h) + h) * 4 * h for l in range(layers))
    alg = real_rows * per_frame
    roof.update({"bound": "tensor", "peak": tensor_peak, "unit": "TFLOP/s", "algorithmic_flops": alg,
                 "achieved": alg / (k_ms * 1e-3) / 1e12 if k_ms else None,
                 "flops_note": "%d FLOP per real frame row (both layers: input projection + recurrence), %d real rows; the "
                               "recurrence is latency bound (600 dependent steps), the fraction is reported against the sustained "
                               "bf16 GEMM peak" % (per_frame, real_rows)})
  else:
    a = cfg["flags"]["moe_num_extend"]
    alg = real_rows * D * 2 + B * a * D * 4
    roof.update({"bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "algorithmic_bytes": alg,
                 "achieved": alg / (k_ms * 1e-3) / 1e9 if k_ms else None,
                 "achieved_nominal": (B * T * D * 2 + B * a * D * 4) / (k_ms * 1e-3) / 1e9 if k_ms else None,
                 "bytes_note": "real frames once (%d of %d rows) + the pooled [B, 8, 1152] fp32 output" % (real_rows, B * T)})
  roof["frac"] = (roof["achieved"] / roof["peak"]) if roof.get("achieved") else None
  traffic = None
  try:
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("config%d_dram_bytes_per_launch" % args.config)
  except Exception:
    pass
  roof["traffic"] = traffic
  line = {
      "metric": "videos/sec", "value": world * B / (ms * 1e-3), "unit": "videos/s", "n_gpus": world, "steps": args.steps,
      "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16",
      "data": "synthetic",
      "config": {"workload": cfg["workload"], "baseline_config": args.config, "batch_per_gpu": B, "global_batch": world * B, "frames": T,
                 "feature_dim": D, "vocab": V, "flags": cfg["flags"], "parallelism": "dp%d" % world, "launch": launch_mode,
                 "operands": "frames bf16, weights bf16, inter-kernel activations %s, fp32 accumulate" %
                             ("fp16 (11 significant bits)" if fmt == "f16" else "bf16 hi+lo pairs"),
                 "l2": "every timed step reads a different input batch (%d distinct resident batches rotated; %d MB of real frames "
                       "each) plus the layer weights, more than the 126 MB L2 between two uses of the same bytes; no explicit flush"
                       % (n_graphs, int(real_rows * D * 2 / 1e6))},
      "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "videos/s", "ms_per_step": ms_e2e,
              "h2d_bytes_per_step": packed_host.nbytes(), "d2h_bytes_per_step": pred_host.numel() * 4,
              # per rank; a PCIe 5.0 x16 link moves ~55 GB/s of pinned host memory: when this figure sits there, e2e is bound
              # by the host link, not by the kernels (compare ms_per_step of the resident `value`)
              "h2d_gb_per_s_per_gpu": packed_host.nbytes() / (ms_e2e * 1e-3) / 1e9,
              "host_batch": "readers.PackedFrames: uint8 real frames only (%d of %d frame rows; num_frames ~ U{30..300}), padded on "
                            "the GPU" % (packed_host.data.shape[0], B * T)},
      "gpu_launches": int(launches),
      "train_step": None if ms_train is None else {
          "value": world * B / (ms_train * 1e-3), "unit": "videos/s", "ms_per_step": ms_train, "steps": args.train_steps,
          "gradient_wire_dtype": args.dp_gradient_dtype,
          "n_gpus": world, "global_batch": world * B,
          "what": "%s; resident inputs; when n_gpus > 1 the flat fp32 gradient (%.1f M floats) is summed over the ranks once per step (NetVLAD models: in three contiguous pieces started as the backward finishes them)" %
                  (train_what, grad_floats / 1e6)},
      "roofline": roof,
      "clocks": clocks,
  }
  if sweep is not None:
    line["batch_sweep"] = sweep
  if world == 1 and not args.no_cpu_baseline:
    line["cpu_baseline"] = time_cpu(args.config, args.cpu_sample or cfg["cpu_sample"], B)
    if args.config == 2:
      line["config1_logistic_train"] = config1_pair(dev)
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
