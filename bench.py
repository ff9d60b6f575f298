#!/usr/bin/env python
"""Benchmark of the yt8m_b200 hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU restatement of the reference path, host cores

Workload = BASELINE.json configs[1]: NetVLAD K=64 over 300x1152 frame features -> hidden FC 73,728->1024
(+BN, ReLU6) -> MoE-2 head over 4716 labels, bf16 operands / fp32 accumulate, batch 256 per GPU.
One "step" = one forward pass of the plugin (NetVLADModel.create_model) over one batch of synthetic
frame features.  The path shards by video: each rank processes its own batch, no data-path collective
("weak" scaling).  `value` = videos/s with inputs resident in HBM; `e2e` = the same through the
reference-facing plugin call chain (DefaultTransformer.transform + create_model) from pinned HOST uint8
features with the predictions copied back to the host, every step.
"""
import argparse
import json
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
K_CLUSTERS, HIDDEN, MIXTURES = 64, 1024, 2
WORKLOAD = "NetVLAD K=64 over 300x1152 frame feats + FC 73728->1024 + MoE-2 head (4716 labels), forward pass"


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--batch", type=int, default=256, help="videos per GPU per step")
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--cpu-sample", type=int, default=16, help="videos per CPU-baseline forward")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--train-steps", type=int, default=8, help="timed steps of the training-step side measurement (0 = skip)")
  ap.add_argument("--graphs", type=int, default=4, help="captured copies of the step rotated through the timed region")
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

def cpu_forward_fn(sample, seed=8):
  """Returns (fn, n_videos): fn() runs the oracle forward of the same workload on `sample` videos."""
  import synth
  from oracle import model_oracle
  g = torch.Generator().manual_seed(9)
  x, nf, _ = synth.model_input(sample, T, D, seed=seed)
  import math
  sd = {
      "cluster_weights": synth.normal((D, K_CLUSTERS), g, 1 / math.sqrt(D)),
      "cluster_weights2": torch.randn(D, K_CLUSTERS, generator=g) / math.sqrt(D),
      "hidden1_weights": synth.normal((K_CLUSTERS * D, HIDDEN), g, 1 / math.sqrt(K_CLUSTERS)),
      "gates/weights": synth.xavier((HIDDEN, V * (MIXTURES + 1)), g),
      "experts/weights": synth.xavier((HIDDEN, V * MIXTURES), g),
      "experts/biases": torch.zeros(V * MIXTURES),
  }
  for scope, c in (("cluster_bn", K_CLUSTERS), ("hidden1_bn", HIDDEN)):
    sd[scope + "/gamma"], sd[scope + "/beta"] = torch.ones(c), torch.zeros(c)
    sd[scope + "/moving_mean"], sd[scope + "/moving_variance"] = torch.zeros(c), torch.ones(c)
  return (lambda: model_oracle.netvlad(sd, x, nf, V, MIXTURES)), sample


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


def time_cpu(sample, min_seconds=10.0, max_reps=100000):
  fn, n = cpu_forward_fn(sample)
  threads = pick_threads(fn)             # also the warm-up (pages in the 300 MB of fp32 weights)
  reps, t0 = 0, time.perf_counter()
  while reps < max_reps and (reps < 2 or time.perf_counter() - t0 < min_seconds):
    fn()
    reps += 1
  dt = time.perf_counter() - t0
  return {"value": n * reps / dt, "unit": "videos/s", "cores": threads, "kind": "port",
          "sample": "%d forwards of %d videos (fp32 torch-CPU restatement of the reference ops, %.1f s)" % (reps, n, dt)}


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


def run_reference(args, rank):
  """--impl reference: the reference's own path is TF-1.0 / python2 and cannot run; the arm
  times the oracle port (kind "port") on the host cores, rank 0 only."""
  if rank != 0:
    return
  fn, n = cpu_forward_fn(args.cpu_sample)
  threads = pick_threads(fn)
  for _ in range(max(1, min(args.warmup, 2))):
    fn()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    fn()
  dt = time.perf_counter() - t0
  val = n * args.steps / dt
  line = {"impl": "reference", "metric": "videos/sec", "value": val, "unit": "videos/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": WORKLOAD, "batch_per_step": n, "frames": T, "feature_dim": D},
          "cpu_baseline": {"value": val, "unit": "videos/s", "cores": threads, "kind": "port",
                           "sample": "each step = one forward of %d videos" % n},
          "e2e": {"value": val, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

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
  import synth
  from yt8m_flags import FLAGS

  if os.environ.get("YT8M_DEBUG_FLAGS"):          # A/B switches of the kernels (tools/); unset in normal runs
    nat.debug_set_flags(int(os.environ["YT8M_DEBUG_FLAGS"]))
  B = args.batch
  FLAGS.parse([], known_only=True)
  FLAGS.netvlad_cluster_size, FLAGS.netvlad_hidden_size, FLAGS.moe_num_mixtures = K_CLUSTERS, HIDDEN, MIXTURES
  FLAGS.video_level_classifier_model = "MoeModel"
  FLAGS.netvlad_operand_format = args.operand_format
  ops.get_store().reset(seed=9)
  model = frame_level_models.NetVLADModel()
  transformer = feature_transform.DefaultTransformer()

  u8, nf = synth.frames_u8(B, T, D, seed=8 + rank)
  u8_pinned, nf_pinned = u8.pin_memory(), nf.pin_memory()
  u8_dev = torch.empty_like(u8, device=dev)
  nf_dev = torch.empty_like(nf, device=dev)
  pred_host = torch.empty((B, V), dtype=torch.float32).pin_memory()
  u8_dev.copy_(u8_pinned)
  nf_dev.copy_(nf_pinned)
  x_dev, _ = transformer.transform(u8_dev, nf_dev)          # resident bf16, L2-normalised rows

  def step_resident():
    return model.create_model(x_dev, vocab_size=V, num_frames=nf_dev)["predictions"]

  # end-to-end: every step copies ITS batch host->device and its predictions device->host inside the timed region.
  # The host batch is what readers.YT8MFrameFeatureReader(packed=True) yields: a readers.PackedFrames holding only the
  # REAL frames of every video (uint8, as stored in the TFRecords) -- the zero padding up to 300 frames is produced
  # on the GPU by the ingest kernel (yt8m_frames_unpack_u8: de-quantise + L2-normalise + pad), so it never crosses PCIe.
  # Like the reference's queue-runner input pipeline (wh/train.py:199-209) the next batch's upload is prefetched: a
  # copy stream fills the other of two device buffers while the compute stream works on the current one.
  import readers
  packed_host = readers.PackedFrames.from_padded(u8, nf).pin_memory()
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

  W = max(args.warmup, 3)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  # The resident step is captured ONCE into CUDA graphs and replayed (ops.CapturedStep): the host side of a
  # step (Python, ctypes, tensor-map encoding, output allocation) would otherwise outlast the ~0.15 ms of GPU work.
  # A few captures are rotated so that each holds its own CUDA-event pair around the NetVLAD kernel; the pairs
  # are read after the timed region (durations of the last timed replays -- measured live, inside the region).
  step_resident()                       # lazy weight packing happens here, outside the launch count
  l0 = nat.launch_count()
  graphs = [ops.CapturedStep(step_resident, warmup=2 if i == 0 else 1, time_tag="netvlad") for i in range(args.graphs)]
  launches = (nat.launch_count() - l0) // (len(graphs) + 1 + len(graphs))     # warm-up calls + one capture each
  it = {"i": 0}

  def step_graph():
    g = graphs[it["i"] % len(graphs)]
    it["i"] += 1
    return g()

  ms = timed(step_graph, args.steps, W)
  kt = [t for g in graphs[:min(len(graphs), args.steps)] for t in g.kernel_ms()]
  ms_e2e = timed(step_e2e, args.steps, W)
  clocks = sampler.stop() if rank == 0 else None

  # side measurement: the training step of the same workload (NetVLADTrainer: forward + full backward + per-tensor
  # clip + Adam; under torchrun ONE NCCL all-reduce of the flat fp32 gradient per step), inputs resident
  ms_train = None
  if args.train_steps > 0:
    import math
    import yt8m_trainer
    gw = torch.Generator(device=dev).manual_seed(9)

    def rnd(shape, std):
      return (torch.randn(shape, generator=gw, device=dev) * std).to(torch.bfloat16).float()

    tr = yt8m_trainer.NetVLADTrainer(D, clusters=K_CLUSTERS, hidden=HIDDEN, vocab=V, mixtures=MIXTURES, device=dev)
    tr.import_state({"cluster_weights": rnd((D, K_CLUSTERS), 1 / math.sqrt(D)), "cluster_biases": torch.zeros(K_CLUSTERS, device=dev),
                     "cluster_weights2": rnd((D, K_CLUSTERS), 1 / math.sqrt(D)),
                     "hidden1_weights": rnd((K_CLUSTERS * D, HIDDEN), 1 / math.sqrt(K_CLUSTERS)),
                     "hidden1_biases": torch.zeros(HIDDEN, device=dev),
                     "gates/weights": rnd((HIDDEN, V * (MIXTURES + 1)), 0.03), "experts/weights": rnd((HIDDEN, V * MIXTURES), 0.03),
                     "experts/biases": torch.zeros(V * MIXTURES, device=dev)})
    y_dev = synth.labels(B, V, seed=8 + rank).to(dev)
    ms_train = timed(lambda: tr.step(x_dev, nf_dev, y_dev, global_batch=world * B), args.train_steps, 3)
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
  # dominant kernel: netvlad_v4_kernel, HBM bound.  Algorithmic bytes / video: the frames once
  # (300*1152*2) + the descriptor once (1152*64*2: one fp16 tensor; x2 for a bf16 hi + lo pair) -- SURVEY.md §8(d),
  # DESIGN.md §Kernels.
  # The kernel streams only the tiles that hold real frames (ceil(num_frames / 32) tiles of 32 frames per video), so the
  # bytes it has to move are those of the REAL frames, not of the zero padding up to 300: `achieved` counts
  # sum_b num_frames[b] * 1152 * 2 + the descriptors; the nominal figure with every video counted as 300 frames
  # (691,200 B per video) is reported beside it as `achieved_nominal`.
  fmt = FLAGS.netvlad_operand_format
  out_bytes = B * D * K_CLUSTERS * 2 * (1 if fmt == "f16" else 2)
  real_rows = int(nf.clamp(0, T).sum())
  alg_bytes = real_rows * D * 2 + out_bytes
  nominal_bytes = B * T * D * 2 + out_bytes
  k_ms = sum(kt) / len(kt) if kt else None
  achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None
  achieved_nominal = nominal_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None
  traffic = None
  try:
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("netvlad_fused_kernel_dram_bytes_per_launch")
  except Exception:
    pass
  line = {
      "metric": "videos/sec", "value": world * B / (ms * 1e-3), "unit": "videos/s", "n_gpus": world, "steps": args.steps,
      "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
      "data": "synthetic",
      "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "frames": T, "feature_dim": D,
                 "clusters": K_CLUSTERS, "hidden": HIDDEN, "mixtures": MIXTURES, "vocab": V, "parallelism": "dp%d" % world,
                 "launch": "step captured once as a CUDA graph (%d kernels) and replayed; %d captures rotated" % (launches, len(graphs)),
                 "operands": "frames bf16, weights bf16, descriptor + hidden layer %s, fp32 accumulate" %
                             ("fp16 (11 significant bits)" if fmt == "f16" else "bf16 hi+lo pairs"),
                 "l2": "inputs larger than L2 (frames 177 MB + FC weights 151 MB per step vs 126 MB L2), no explicit flush"},
      "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "videos/s", "ms_per_step": ms_e2e,
              "h2d_bytes_per_step": packed_host.nbytes(), "d2h_bytes_per_step": pred_host.numel() * 4,
              "host_batch": "readers.PackedFrames: uint8 real frames only (%d of %d frame rows; num_frames ~ U{30..300}), padded on "
                            "the GPU" % (packed_host.data.shape[0], B * T)},
      "gpu_launches": int(launches),
      "train_step": None if ms_train is None else {
          "value": world * B / (ms_train * 1e-3), "unit": "videos/s", "ms_per_step": ms_train, "steps": args.train_steps,
          "what": "NetVLAD + FC + MoE-2 forward, full backward, per-tensor clip + Adam; resident inputs; "
                  "one all-reduce of the flat fp32 gradient (%d M floats) per step when n_gpus > 1" % 100},
      "roofline": {"bound": "hbm", "kernel": "netvlad_v4_kernel (K=64)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                   "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic,
                   "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes, "achieved_nominal": achieved_nominal,
                   "bytes_note": "real frames only (%d of %d frame rows; padded tiles are not read) + fp16 descriptors; "
                                 "achieved_nominal counts every video as 300 frames" % (real_rows, B * T), "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback"},
      "clocks": clocks,
  }
  if world == 1 and not args.no_cpu_baseline:
    line["cpu_baseline"] = time_cpu(args.cpu_sample)
    line["config1_logistic_train"] = config1_pair(dev)
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
