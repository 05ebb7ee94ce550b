#!/usr/bin/env python
"""Is the tcgen05 Gram bit-symmetric (S[i,j] == S[j,i]) and shape-independent?  (needed for the
threshold-based hard-negative membership test on the tensor path)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import debug_lib
from wav2vec_contr_loss_b200.functional import _p, _stream
lib = debug_lib.load(); dev = torch.device("cuda:0")
torch.manual_seed(1)
n = 512
z = torch.nn.functional.normalize(torch.randn(n, 256), dim=1).to(dev).to(torch.bfloat16)
z[300] = z[7]; z[301] = z[7]      # exact duplicates -> exact ties
TS = 1 << 24
def tile(ri, rj, ts):
    s = torch.empty(128, 128, device=dev); o = torch.empty(128, 256, device=dev)
    debug_lib.check(lib.supcon_debug_tc_tile(_p(z), n, 256, ri + (TS if ts else 0), rj, _p(s), _p(o), _stream(dev)), "dbg")
    torch.cuda.synchronize(); return s
for ts in (False, True):
    a = tile(0, 128, ts); b = tile(128, 0, ts)
    print("A-from-TMEM" if ts else "SS", "S(0,128) == S(128,0)^T bitwise:", bool(torch.equal(a, b.t().contiguous())),
          "max abs diff", float((a - b.t()).abs().max()))
    d = tile(256, 0, ts)
    print("  duplicate columns identical:", bool(torch.equal(d[300 - 256], d[301 - 256])))
a = tile(0, 128, False); b = tile(0, 128, True)
print("SS vs TS identical:", bool(torch.equal(a, b)))
ref = (z[:128].double() @ z[128:256].double().t())
print("max |S - fp64|", float((a.double() - ref).abs().max()))
