"""World-size-2 gloo test (CPU) of the data-parallel host logic: the flat-buffer all-reduce of gradients
pre-scaled by B_local / B_global reproduces the single-process gradient of the mean loss on the concatenated
batch (the semantics yt8m_trainer.HeadTrainer relies on), and the row sharding covers the batch exactly once."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
  sys.path.insert(0, ROOT)
  os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                     "MASTER_PORT": str(port)})
  import yt8m_dp
  from oracle import yt8m_oracle as O
  r, w, _ = yt8m_dp.init_from_env(backend="gloo")
  assert (r, w) == (rank, world) and yt8m_dp.world_size() == world
  g = torch.Generator().manual_seed(0)
  B, D, V = 12, 16, 10
  x = torch.randn(B, D, generator=g)
  y = (torch.rand(B, V, generator=g) < 0.3).float()
  w0 = (torch.randn(D, V, generator=g) * 0.3).requires_grad_(True)
  b0 = torch.zeros(V, requires_grad=True)
  lo, hi = yt8m_dp.shard_rows(B)
  # local gradient of the LOCAL mean loss, rescaled by B_local / B_global, summed over ranks
  loss = O.cross_entropy_loss(O.logistic_model(x[lo:hi], w0, b0), y[lo:hi]) * ((hi - lo) / float(B))
  loss.backward()
  flat = torch.cat([w0.grad.reshape(-1), b0.grad.reshape(-1)])
  # the trainers' exchange: contiguous pieces of the flat buffer, started separately, finished together
  pieces = flat.clone()
  xch = yt8m_dp.GradExchange()
  xch.start(pieces[100:], "a")
  xch.start(pieces[40:100], "b")
  xch.start(pieces[:40], "c")
  xch.start(pieces[:0], "empty")           # an empty piece is skipped
  assert len(xch.pending) == 3
  xch.wait("b")                            # pieces complete in issue order: waiting for b covers a
  assert [n for n, _ in xch.pending] == ["c"]
  xch.wait("empty")                        # unknown / skipped name: nothing to wait for
  xch.finish()
  assert not xch.pending
  solo = flat.clone()
  one = yt8m_dp.GradExchange(world=1)      # a single-process run inside a larger job: no exchange
  one.start(solo)
  one.finish()
  assert torch.equal(solo, flat)
  yt8m_dp.all_reduce_sum_(flat)
  assert torch.equal(pieces, flat)
  q.put((rank, lo, hi, flat.clone()))
  torch.distributed.barrier()
  torch.distributed.destroy_process_group()


def test_dp_allreduce_matches_single_process():
  sys.path.insert(0, ROOT)
  from oracle import yt8m_oracle as O
  world, port = 2, _free_port()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in range(world)])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  # shards tile the batch
  assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 12
  # both ranks hold the same reduced gradient == single-process gradient on the whole batch
  g = torch.Generator().manual_seed(0)
  B, D, V = 12, 16, 10
  x = torch.randn(B, D, generator=g)
  y = (torch.rand(B, V, generator=g) < 0.3).float()
  w0 = (torch.randn(D, V, generator=g) * 0.3).requires_grad_(True)
  b0 = torch.zeros(V, requires_grad=True)
  O.cross_entropy_loss(O.logistic_model(x, w0, b0), y).backward()
  want = torch.cat([w0.grad.reshape(-1), b0.grad.reshape(-1)])
  assert torch.allclose(res[0][3], res[1][3])
  assert torch.allclose(res[0][3], want, atol=1e-6)


def test_lr_schedule_and_adam_scalars():
  sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
  import importlib
  import math
  # yt8m_trainer imports the native library (loads on CPU; no kernels are launched here)
  tr = importlib.import_module("yt8m_trainer")
  from oracle import yt8m_oracle as O
  for step in (0, 1, 3999, 4000, 8001):
    assert tr.exponential_decay(0.01, step, 1000, 4000000, 0.95) == O.exponential_decay(0.01, step, 1000, 4000000, 0.95)
  assert abs(tr.adam_lr_t(0.01, 1) - 0.01 * math.sqrt(0.001) / 0.1) < 1e-12
  gates, experts = tr._moe_row_index(4716, 2)
  assert gates.numel() == 4716 * 3 and experts.numel() == 4716 * 2
  assert len(set(gates.tolist()) | set(experts.tolist())) == 4716 * 5            # a bijection onto distinct packed rows
  assert int(max(gates.max(), experts.max())) < 128 * 189
