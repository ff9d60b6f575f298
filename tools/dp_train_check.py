"""torchrun --nproc-per-node N tools/dp_train_check.py : N-rank data-parallel training of the MoE head equals the
single-process run on the concatenated batch; rank 0 prints the verdict."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("youtube-8m_b200", "tests", ""):
  sys.path.insert(0, os.path.join(ROOT, p))
import synth, yt8m_dp, yt8m_trainer

rank, world, local = yt8m_dp.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B, D, V, M = 256, 1024, 4716, 2
g = torch.Generator().manual_seed(0)
x = torch.randn(B, D, generator=g)
x = synth.bf16r(x * torch.rsqrt((x * x).sum(1, keepdim=True)))
y = synth.labels(B, V)
gain = math.sqrt(D) / 4
sd = {"gates/weights": synth.xavier((D, V * (M + 1)), g, gain), "experts/weights": synth.xavier((D, V * M), g, gain),
      "experts/biases": torch.zeros(V * M)}

def run(group_world):
  t = yt8m_trainer.HeadTrainer("moe", D, V, mixtures=M, device=dev)
  t.keep_grads = True
  if group_world == 1:
    t.world, t.group = 1, None
    lo, hi = 0, B
  else:
    lo, hi = yt8m_dp.shard_rows(B)
  t.import_state({k: v.to(dev) for k, v in sd.items()})
  for _ in range(3):
    if group_world == 1:
      # single-process reference: bypass the collective
      import yt8m_dp as dp
      saved = dp.all_reduce_sum_
      dp.all_reduce_sum_ = lambda flat, group=None: flat
      t.step(x[lo:hi].to(dev).to(torch.bfloat16), y[lo:hi].to(dev), global_batch=B)
      dp.all_reduce_sum_ = saved
    else:
      t.step(x[lo:hi].to(dev).to(torch.bfloat16), y[lo:hi].to(dev), global_batch=B)
  torch.cuda.synchronize()
  return t.param.clone(), t.last_grad.clone()

dp_param, dp_grad = run(world)
single, single_grad = run(1)
# gradients of the third step (after the all-reduce): relative to the largest entry
gerr = float((dp_grad - single_grad).abs().max() / single_grad.abs().max())
# parameters after 3 Adam steps: each step moves a weight by ~lr = 1e-2; near-zero gradients are ill-conditioned
frac_bad = float(((dp_param - single).abs() > 0.05 * 3e-2).float().mean())
stat = torch.tensor([gerr, frac_bad], device=dev)
if world > 1:
  torch.distributed.all_reduce(stat, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
  ok = float(stat[0]) < 2e-3 and float(stat[1]) < 2e-3
  print("DP world=%d vs single process: grad rel err %.2e, params off by >5%% of their displacement: %.4f%%  -> %s"
        % (world, float(stat[0]), 100 * float(stat[1]), "OK" if ok else "MISMATCH"))

# ---- the frame-level step: NetVLAD + FC + MoE, gradient of the FIRST step (all-reduced) vs the single-process batch
Bf, T, Df, K, H, Vf = 16, 300, 1152, 64, 256, 500
xf, nff, _ = synth.model_input(Bf, T, Df, seed=5)
yf = synth.labels(Bf, Vf, seed=5, per_video=3.4)
sdf = {"cluster_weights": synth.normal((Df, K), g, 4.0), "cluster_biases": 0.1 * torch.randn(K, generator=g),
       "cluster_weights2": synth.normal((Df, K), g, 1 / math.sqrt(Df)), "hidden1_weights": synth.normal((K * Df, H), g, 12.0 / math.sqrt(K)),
       "hidden1_biases": 0.1 * torch.randn(H, generator=g), "gates/weights": synth.xavier((H, Vf * (M + 1)), g, 2.0),
       "experts/weights": synth.xavier((H, Vf * M), g, 2.0), "experts/biases": 0.1 * torch.randn(Vf * M, generator=g)}

def run_frames(group_world):
  t = yt8m_trainer.NetVLADTrainer(Df, clusters=K, hidden=H, vocab=Vf, mixtures=M, device=dev)
  t.keep_grads = True
  t.import_state(sdf)
  import yt8m_dp as dp
  saved = dp.all_reduce_sum_
  if group_world == 1:
    t.world = t.head.world = 1                      # the gradient exchange of this trainer is skipped
    dp.all_reduce_sum_ = lambda flat, group=None: flat
    lo, hi = 0, Bf
  else:
    lo, hi = yt8m_dp.shard_rows(Bf)
  t.step(xf[lo:hi].to(dev).to(torch.bfloat16), nff[lo:hi].to(dev), yf[lo:hi].to(dev), global_batch=Bf)
  dp.all_reduce_sum_ = saved
  torch.cuda.synchronize()
  return t.last_grad.clone()

gd, gs = run_frames(world), run_frames(1)
ferr = torch.tensor([float((gd - gs).norm() / gs.norm())], device=dev)
if world > 1:
  torch.distributed.all_reduce(ferr, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
  print("NetVLAD step, DP world=%d vs single process: flat gradient rel L2 err %.2e -> %s" % (world, float(ferr), "OK" if float(ferr) < 2e-3 else "MISMATCH"))

# ---- LstmModel step (persistent recurrence forward + BPTT) and AttentionModel + MoeExtend step: same check
def run_generic(make, xs, nfs, ys, sd_, group_world, bsz):
  t = make()
  t.keep_grads = True
  t.import_state(sd_)
  import yt8m_dp as dp
  saved = dp.all_reduce_sum_
  if group_world == 1:
    t.world = 1
    if hasattr(t, "head"):
      t.head.world = 1
    dp.all_reduce_sum_ = lambda flat, group=None: flat
    lo, hi = 0, bsz
  else:
    lo, hi = yt8m_dp.shard_rows(bsz)
  t.step(xs[lo:hi].to(dev).to(torch.bfloat16), nfs[lo:hi].to(dev), ys[lo:hi].to(dev), global_batch=bsz)
  dp.all_reduce_sum_ = saved
  torch.cuda.synchronize()
  return t.last_grad.clone()

Bl, Tl, Dl, Hl, Ll, Vl = 16, 40, 128, 256, 2, 300
xl, nfl, _ = synth.model_input(Bl, Tl, Dl, seed=6, min_frames=4)
yl = synth.labels(Bl, Vl, seed=6, per_video=3.4)
sdl = {"gates/weights": synth.xavier((Ll * 2 * Hl, Vl * (M + 1)), g, 2.0), "experts/weights": synth.xavier((Ll * 2 * Hl, Vl * M), g, 2.0),
       "experts/biases": 0.1 * torch.randn(Vl * M, generator=g)}
for l in range(Ll):
  sdl[yt8m_trainer.LstmTrainer.SCOPE % l + "/weights"] = synth.xavier(((Dl if l == 0 else Hl) + Hl, 4 * Hl), g, 2.0)
  sdl[yt8m_trainer.LstmTrainer.SCOPE % l + "/biases"] = synth.bf16r(0.1 * torch.randn(4 * Hl, generator=g))
mk = lambda: yt8m_trainer.LstmTrainer(Dl, hidden=Hl, layers=Ll, vocab=Vl, mixtures=M, device=dev)
gd, gs = run_generic(mk, xl, nfl, yl, sdl, world, Bl), run_generic(mk, xl, nfl, yl, sdl, 1, Bl)
lerr = torch.tensor([float((gd - gs).norm() / gs.norm())], device=dev)
A = 8
sda = {"Attention/W": synth.bf16r(torch.randn(2 * Dl, A, generator=g) * 3.0), "Attention/b": torch.full((A,), 0.1),
       "gates/weights": synth.xavier((Dl, Vl * (M + 1)), g, 6.0), "experts/weights": synth.xavier((Dl, Vl * M), g, 6.0),
       "experts/biases": 0.1 * torch.randn(Vl * M, generator=g)}
mka = lambda: yt8m_trainer.AttentionTrainer(Dl, heads=A, vocab=Vl, mixtures=M, device=dev)
gd, gs = run_generic(mka, xl, nfl, yl, sda, world, Bl), run_generic(mka, xl, nfl, yl, sda, 1, Bl)
aerr = torch.tensor([float((gd - gs).norm() / gs.norm())], device=dev)
if world > 1:
  torch.distributed.all_reduce(lerr, op=torch.distributed.ReduceOp.MAX)
  torch.distributed.all_reduce(aerr, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
  print("LSTM step, DP world=%d vs single process: flat gradient rel L2 err %.2e -> %s" % (world, float(lerr), "OK" if float(lerr) < 2e-3 else "MISMATCH"))
  print("Attention step, DP world=%d vs single process: flat gradient rel L2 err %.2e -> %s" % (world, float(aerr), "OK" if float(aerr) < 2e-3 else "MISMATCH"))
if world > 1:
  torch.distributed.destroy_process_group()
