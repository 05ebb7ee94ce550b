#!/usr/bin/env python
"""Check the tcgen05/TMA building blocks (supcon_debug_tc_tile) against torch."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import debug_lib
from wav2vec_contr_loss_b200.functional import _p, _stream

lib = debug_lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
n = 512
z = torch.nn.functional.normalize(torch.randn(n, 256), dim=1).to(dev).to(torch.bfloat16)
TS = 1 << 24
for (ri, rj) in [(0, 0), (128, 256), (384, 128), (448, 64), (TS + 0, 0), (TS + 128, 256), (TS + 448, 64)]:
    s = torch.full((128, 128), float("nan"), device=dev)
    o = torch.full((128, 256), float("nan"), device=dev)
    debug_lib.check(lib.supcon_debug_tc_tile(_p(z), n, 256, ri, rj, _p(s), _p(o), _stream(dev)), "debug")
    torch.cuda.synchronize()
    mode = "TS" if ri >= TS else "SS"
    ri = ri % TS
    zi, zj = torch.zeros(128, 256, device=dev), torch.zeros(128, 256, device=dev)
    a = z[ri:ri + 128].float(); b = z[rj:rj + 128].float()
    zi[: a.size(0)] = a; zj[: b.size(0)] = b
    s_ref = zi.double() @ zj.double().t()
    o_ref = s.to(torch.bfloat16).double() @ zj.double()
    es = (s.double() - s_ref).abs().max().item()
    eo = (o.double() - o_ref).abs().max().item()
    print(f"{mode} tile ({ri},{rj}): max|S-ref|={es:.3e} (|S|max {s_ref.abs().max():.3f})  max|O-ref|={eo:.3e} (|O|max {o_ref.abs().max():.3f})")
    if not (es < 1e-5):
        bad = (s.double() - s_ref).abs() > 1e-5
        print("  S mismatch count", int(bad.sum()), "first rows", bad.any(1).nonzero()[:8].flatten().tolist(),
              "first cols", bad.any(0).nonzero()[:8].flatten().tolist())
        print("  S[0,:8]", s[0, :8].tolist(), "\n  ref   ", s_ref[0, :8].tolist())
    if not (eo < 1e-4):
        bad = (o.double() - o_ref).abs() > 1e-4
        print("  O mismatch count", int(bad.sum()), "rows", bad.any(1).nonzero()[:8].flatten().tolist(),
              "cols", bad.any(0).nonzero()[:16].flatten().tolist())
        print("  O[0,:8]", o[0, :8].tolist(), "\n  ref   ", o_ref[0, :8].tolist())
