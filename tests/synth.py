"""Seed-fixed synthetic inputs shared by the tests, smoke() and bench.py.

Frames are uint8 U{0..255} de-quantised with the reader's formula (wh/utils.py:34-38), rows at or beyond
num_frames zeroed (reader padding, wh/readers.py:47-50), L2-normalised per row
(default_transformer.py) and rounded to bf16 -- both the oracle and the CUDA path see exactly these values.
"""
import math

import torch


def bf16r(x):
  return x.to(torch.bfloat16).to(torch.float32)


def frames_u8(batch, frames=300, dim=1152, seed=8, min_frames=30):
  g = torch.Generator().manual_seed(seed)
  u8 = torch.randint(0, 256, (batch, frames, dim), generator=g, dtype=torch.uint8)
  nf = torch.randint(min(min_frames, frames), frames + 1, (batch,), generator=g, dtype=torch.int32)
  return u8, nf


def model_input(batch, frames=300, dim=1152, seed=8, min_frames=30):
  """Returns (x fp32 [B,T,D] bf16-representable & L2-normalised, num_frames int32 [B], raw uint8)."""
  u8, nf = frames_u8(batch, frames, dim, seed, min_frames)
  x = u8.to(torch.float32) * (4.0 / 255.0) + (4.0 / 512.0 - 2.0)
  mask = (torch.arange(frames).unsqueeze(0) < nf.unsqueeze(1)).to(torch.float32).unsqueeze(2)
  x = x * mask
  ss = (x * x).sum(dim=2, keepdim=True)
  x = x * torch.rsqrt(torch.clamp(ss, min=1e-12))
  return bf16r(x), nf, u8


def labels(batch, vocab=4716, seed=8, per_video=3.4):
  g = torch.Generator().manual_seed(seed + 1000)
  y = (torch.rand((batch, vocab), generator=g) < per_video / vocab)
  for b in range(batch):
    if not y[b].any():
      y[b, int(torch.randint(0, vocab, (1,), generator=g))] = True
  return y.to(torch.float32)


def xavier(shape, gen, gain=1.0):
  fan_in, fan_out = shape[0], shape[1]
  lim = gain * math.sqrt(6.0 / (fan_in + fan_out))
  return bf16r((torch.rand(shape, generator=gen) * 2 - 1) * lim)


def normal(shape, gen, std):
  return bf16r(torch.randn(shape, generator=gen) * std)
