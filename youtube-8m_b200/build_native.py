"""Builds libyt8m_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python youtube-8m_b200/build_native.py [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so lands next to this file so that it travels
to the GPU box with the source snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libyt8m_b200.so")
SOURCES = ["yt8m_host.cu", "yt8m_gemm.cu", "yt8m_rowops.cu", "yt8m_netvlad.cu", "yt8m_train.cu", "yt8m_netvlad_bwd.cu", "yt8m_netvlad_bwd_tc.cu", "yt8m_netvlad_v4.cu", "yt8m_netvlad_v5.cu", "yt8m_netvlad_v6.cu", "yt8m_lstm_rec.cu", "yt8m_seq_bwd.cu", "yt8m_bn.cu", "yt8m_attn_fused.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("YT8M_NVCC_EXTRA", "").split()


def _deps():
  deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
  deps.append(os.path.join(HERE, "..", "include", "yt8m_b200.h"))
  deps.append(os.path.abspath(__file__))
  return deps


def needs_build():
  if not os.path.exists(OUT):
    return True
  t = os.path.getmtime(OUT)
  return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def build(force=False, verbose=False):
  """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
  if not force and not needs_build():
    return OUT
  srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
  objdir = os.path.join(HERE, "build")
  os.makedirs(objdir, exist_ok=True)

  def cc(src):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
      sys.stderr.write(r.stderr)
    return obj

  with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
    objs = list(ex.map(cc, srcs))
  cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
  return OUT


if __name__ == "__main__":
  p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
  print(p)
