import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-8m_b200"))
import yt8m_native as nat
dev = "cuda:0"
torch.manual_seed(0)
for (M, N, Kb, split) in [(256, 256, 128, False), (256, 256, 128, True), (256, 256, 256, False), (128, 128, 64, False), (4716, 1152, 128, True)]:
  a = torch.randn(Kb, M, device=dev).to(torch.bfloat16)
  b = torch.randn(Kb, N, device=dev).to(torch.bfloat16)
  lda, ldb = nat.pad8(M), nat.pad8(N)
  ah = torch.zeros(Kb, lda, dtype=torch.bfloat16, device=dev); ah[:, :M] = a
  bh = torch.zeros(Kb, ldb, dtype=torch.bfloat16, device=dev); bh[:, :N] = b
  al = torch.zeros_like(ah) if split else None
  out = nat.wgrad(ah[:, :M], al[:, :M] if split else None, bh[:, :N], M, N)
  want = a.float().t() @ b.float()
  err = (out - want).abs()
  print("M=%d N=%d Kb=%d split=%s  max err %.4f (max |want| %.2f)" % (M, N, Kb, split, float(err.max()), float(want.abs().max())))
  mb, nb = min(M, 256) // 64, min(N, 256) // 64
  blk = err[:mb * 64, :nb * 64].reshape(mb, 64, nb, 64).amax(dim=(1, 3))
  print(blk.cpu().numpy().round(2))
  if err.max() > 0.05:
    # which contraction rows are missing?  use one-hot batch rows
    for r in range(0, Kb, 16):
      a1 = torch.zeros_like(ah); a1[r, :M] = 1
      o = nat.wgrad(a1[:, :M], None, bh[:, :N], M, N)
      ok = torch.allclose(o[0], b[r].float(), atol=1e-2)
      ok2 = torch.allclose(o[min(M - 1, 100)], b[r].float(), atol=1e-2)
      print("  batch row %3d -> out row0 %s, out row100 %s" % (r, ok, ok2), end=";")
    print()
