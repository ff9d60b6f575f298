"""Run a few NetVLADTrainer steps at the bench configuration (for `ncu --metrics gpu__time_duration.sum` launch lists):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step_profile.py
and, without ncu, print the CUDA-event time of a step."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import synth, yt8m_trainer, yt8m_native as nat

dev = torch.device("cuda", 0)
B, T, D, K, H, V, M = int(os.environ.get("B", 256)), 300, 1152, 64, 1024, 4716, 2
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device=dev).manual_seed(9)
rnd = lambda shape, std: (torch.randn(shape, generator=g, device=dev) * std).to(torch.bfloat16).float()
tr = yt8m_trainer.NetVLADTrainer(D, clusters=K, hidden=H, vocab=V, mixtures=M, device=dev)
tr.import_state({"cluster_weights": rnd((D, K), 1 / math.sqrt(D)), "cluster_biases": torch.zeros(K, device=dev),
                 "cluster_weights2": rnd((D, K), 1 / math.sqrt(D)), "hidden1_weights": rnd((K * D, H), 1 / math.sqrt(K)),
                 "hidden1_biases": torch.zeros(H, device=dev), "gates/weights": rnd((H, V * (M + 1)), 0.03),
                 "experts/weights": rnd((H, V * M), 0.03), "experts/biases": torch.zeros(V * M, device=dev)})
u8, nf = synth.frames_u8(B, T, D, seed=8)
x = nat.l2norm_rows(u8.to(dev), num_frames=nf.to(dev))
x = x[0] if isinstance(x, tuple) else x
nfd, y = nf.to(dev), synth.labels(B, V, seed=8).to(dev)
for _ in range(2):
  tr.step(x, nfd, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
  tr.step(x, nfd, y)
e1.record()
torch.cuda.synchronize()
print("train step: %.3f ms" % (e0.elapsed_time(e1) / steps))
