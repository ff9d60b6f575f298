"""torchrun --nproc-per-node N tools/allreduce_probe.py : device-timed NCCL all-reduce of the training step's payloads
(the 400 MB fp32 flat gradient of BASELINE config 2, its three pieces, and the same element count in bf16)."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
import yt8m_dp

rank, world, local = yt8m_dp.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
N = 100_400_000
for name, n, dt in (("flat fp32", N, torch.float32), ("head piece fp32", 24_300_000, torch.float32), ("fc piece fp32", 75_500_000, torch.float32),
                    ("pool piece fp32", 147_520, torch.float32), ("flat bf16", N, torch.bfloat16)):
  buf = torch.zeros(n, dtype=dt, device=dev)
  for _ in range(3):
    dist.all_reduce(buf)
  torch.cuda.synchronize()
  dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    dist.all_reduce(buf)
  e1.record()
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  if rank == 0:
    gb = n * buf.element_size() / 1e9
    print("%-16s %7.1f MB  %.3f ms  algbw %.0f GB/s  busbw %.0f GB/s  (NCCL_ALGO=%s)" % (
        name, gb * 1e3, float(ms), gb / float(ms) * 1e3, gb / float(ms) * 1e3 * 2 * (world - 1) / world, os.environ.get("NCCL_ALGO", "default")))
  del buf
dist.barrier()
dist.destroy_process_group()
