"""Data-parallel plumbing: one process per GPU, torch.distributed for the rendezvous and the collective.

The path shards by video: forward / inference needs no collective at all; the training
step has exactly one exchange -- the sum over ranks of the flat fp32 gradient buffer (NCCL over NVLink/NVSwitch on
the GPU box; gloo for the CPU tests of this host logic), issued as one all-reduce or as a few contiguous pieces
that overlap the backward (GradExchange)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
  """Initialises torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if world > 1 and not dist.is_initialized():
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
      torch.cuda.set_device(local_rank)
      dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    else:
      dist.init_process_group(backend)
  return rank, world, local_rank


def world_size(group=None):
  return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
  return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def all_reduce_sum_(flat, group=None):
  """In-place sum over ranks of one flat buffer; a no-op for a single process."""
  if world_size(group) > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  return flat


class _Converted:
  """An all-reduce that ran on a narrower copy: wait() also widens the sum back into the gradient buffer."""

  def __init__(self, work, piece, wire):
    self.work, self.piece, self.wire = work, piece, wire

  def wait(self):
    self.work.wait()
    self.piece.copy_(self.wire)


class GradExchange:
  """The gradient exchange of one training step.  The flat fp32 gradient buffer is summed over the ranks in a few
  CONTIGUOUS pieces, each handed to the collective library as soon as the backward has written its last element:
  NCCL runs on its own stream, so the 300 MB weight gradient of the hidden layer travels over NVLink while the
  pooling layer's backward still computes.  (Round 1 issued one all-reduce after the whole backward: 1.07 ms of a
  4.14 ms step were a fully exposed collective at 8 ranks -- VERDICT r01 item 6.)  start() orders the piece after
  everything already enqueued on the current stream; wait(name) / finish() make the current stream wait for one piece /
  every piece (the host is not blocked with NCCL; gloo, used by the CPU tests, blocks).  A single process: all no-ops."""

  def __init__(self, group=None, world=None, wire_dtype=None):
    """wire_dtype=torch.bfloat16: the pieces travel as bf16 (half the bytes; every rank's contribution is rounded to 8
    significant bits before the sum -- NOT the reference's arithmetic, off unless asked for: --dp_gradient_dtype)."""
    self.group = group
    self.pending = []                                 # (name, work, finalize) in issue order
    self.world = world_size(group) if world is None else world      # world=1: a single-process run inside a larger job
    self.wire_dtype = wire_dtype

  def start(self, piece, name=None):
    if self.world > 1 and piece.numel() > 0:
      if self.wire_dtype is not None and self.wire_dtype != piece.dtype and piece.numel() >= 1 << 20:
        wire = piece.to(self.wire_dtype)
        work = dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((name, _Converted(work, piece, wire)))
      else:
        self.pending.append((name, dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group, async_op=True)))

  def wait(self, name):
    """Orders the current stream after the piece called `name` (and, the collectives being issued in order, after every
    piece started before it): the optimiser update of a piece that has arrived runs under the transfer of the next."""
    for i, (n, work) in enumerate(self.pending):
      if n == name:
        for _, w in self.pending[:i + 1]:
          w.wait()
        self.pending = self.pending[i + 1:]
        return

  def finish(self):
    for _, work in self.pending:
      work.wait()
    self.pending = []


def shard_rows(n_rows, group=None):
  """Contiguous row range [lo, hi) of the global batch owned by this rank (rank r owns rows
  [r*B/W, (r+1)*B/W) -- SURVEY.md §8e)."""
  w, r = world_size(group), rank(group)
  base, extra = divmod(n_rows, w)                # balanced: shard sizes differ by at most one row; a rank is empty only
  lo = r * base + min(r, extra)                  # when n_rows < world (callers drop such a batch on every rank together)
  return lo, lo + base + (1 if r < extra else 0)
