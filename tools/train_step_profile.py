"""Per-kernel breakdown of one optimiser step (NetVLAD + FC + MoE, BASELINE config 2 shapes) from the CUPTI activity records:

    python tools/train_step_profile.py [B] [K]

Prints the event-timed step and the kernels sorted by device time (3 profiled steps; per-step figures)."""
import math, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import synth, yt8m_dp, yt8m_trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64          # K = 0: the LSTM trainer of BASELINE config 3 (2 x 1024 cells, MoE-4)
T, D, H, V, M = 300, 1152, 1024, 4716, 2
rank, world, local = yt8m_dp.init_from_env()          # under torchrun: every rank steps its own shard, rank 0 prints
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
g = torch.Generator().manual_seed(0)
x, nf, _ = synth.model_input(B, T, D, seed=1)
y = synth.labels(B, V, seed=1, per_video=3.4)
if K > 0:
  sd = {"cluster_weights": synth.normal((D, K), g, 4.0), "cluster_biases": 0.1 * torch.randn(K, generator=g),
        "cluster_weights2": synth.normal((D, K), g, 1 / math.sqrt(D)), "hidden1_weights": synth.normal((K * D, H), g, 12.0 / math.sqrt(K)),
        "hidden1_biases": 0.1 * torch.randn(H, generator=g), "gates/weights": synth.xavier((H, V * (M + 1)), g, 2.0),
        "experts/weights": synth.xavier((H, V * M), g, 2.0), "experts/biases": 0.1 * torch.randn(V * M, generator=g)}
if K > 0:
  tr = yt8m_trainer.NetVLADTrainer(D, clusters=K, hidden=H, vocab=V, mixtures=M, device=dev)
else:
  M, L = 4, 2
  sd = {"gates/weights": synth.xavier((L * 2 * H, V * (M + 1)), g, 2.0), "experts/weights": synth.xavier((L * 2 * H, V * M), g, 2.0),
        "experts/biases": 0.1 * torch.randn(V * M, generator=g)}
  for l in range(L):
    sd[yt8m_trainer.LstmTrainer.SCOPE % l + "/weights"] = synth.xavier(((D if l == 0 else H) + H, 4 * H), g, 2.0)
    sd[yt8m_trainer.LstmTrainer.SCOPE % l + "/biases"] = synth.bf16r(0.1 * torch.randn(4 * H, generator=g))
  tr = yt8m_trainer.LstmTrainer(D, hidden=H, layers=L, vocab=V, mixtures=M, device=dev)
tr.import_state(sd)
xd, nfd, yd = x.to(dev).to(torch.bfloat16), nf.to(dev), y.to(dev)
for _ in range(3):
  tr.step(xd, nfd, yd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
  tr.step(xd, nfd, yd)
e1.record()
torch.cuda.synchronize()
if rank == 0:
  print("world %d step: %.3f ms (10 steps, events)" % (world, e0.elapsed_time(e1) / 10))
N = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  for _ in range(N):
    tr.step(xd, nfd, yd)
  torch.cuda.synchronize()
if rank != 0:
  if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
  sys.exit(0)
rows = [(e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("device time per step: %.1f us in %d kernels" % (tot, sum(r[2] for r in rows)))
for k, us, n in rows[:28]:
  print("%9.1f us %5.1f%% x%-4g %s" % (us, 100 * us / tot, n, k[:150]))
# the launches of the LAST profiled step, in order
ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time_total > 0], key=lambda e: e.time_range.start)
per = len(ev) // N
print("---- launches of one step in start order (start offset us, duration us)")
t0 = ev[-per].time_range.start
for e in (ev[-per:] if per < 200 else []):
  print("%9.1f %8.1f  %s" % (e.time_range.start - t0, e.device_time_total, e.name[:110]))
if world > 1:
  torch.distributed.barrier()
  torch.distributed.destroy_process_group()
