"""pytest configuration: registers the `gpu` marker and puts the product directory
(youtube-8m_b200/, a flat script directory like the reference's youtube-8m-wangheda/) and the repo
root (for `oracle`) on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "youtube-8m_b200")
for p in (PKG, ROOT):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
