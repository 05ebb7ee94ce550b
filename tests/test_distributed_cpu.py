"""CPU, gloo, world_size 2: the row-sharded host logic (all-gather -> row-block
forward -> all-reduce -> stats all-gather -> row-block backward) equals the
single-process result on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import supcon_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cfg, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_kernels import OracleKernels
    from wav2vec_contr_loss_b200.distributed import ShardedSupConLoss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d = cfg["n"], cfg["d"]
        x, y = O.make_inputs(n, d, "clustered", classes=3)
        z = F.normalize(x, dim=1)
        n_local = n // world
        zl = z[rank * n_local:(rank + 1) * n_local].clone().requires_grad_(True)
        yl = y[rank * n_local:(rank + 1) * n_local]
        mod = ShardedSupConLoss(cfg["tau"], cfg["sim"], cfg["lam"], 2.0, kernels=OracleKernels)
        loss = mod(zl, yl, topk_neg=cfg["k"], alpha=cfg["alpha"])
        (1.5 * loss).backward()
        out[rank] = (float(loss), zl.grad.double().clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sim,lam,k,alpha", [("cosine", 0.0, 5, 0.0), ("geodesic", 0.1, 4, 0.4)])
def test_two_ranks_match_single_process(sim, lam, k, alpha):
    cfg = dict(n=48, d=16, tau=0.1, sim=sim, lam=lam, k=k, alpha=alpha)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), cfg, out), nprocs=2, join=True)
    x, y = O.make_inputs(48, 16, "clustered", classes=3)
    z = F.normalize(x, dim=1)
    ref = O.closed_form(z, y, temperature=0.1, similarity=sim, uniformity_weight=lam, topk_neg=k, alpha=alpha)
    for rank in (0, 1):
        loss, grad = out[rank]
        assert loss == pytest.approx(ref["loss"], rel=1e-6)
        want = 1.5 * ref["dz"][rank * 24:(rank + 1) * 24]
        assert torch.allclose(grad, want, rtol=1e-5, atol=1e-7)


def test_requires_process_group():
    from wav2vec_contr_loss_b200.distributed import ShardedSupConLoss
    with pytest.raises(RuntimeError, match="process group"):
        ShardedSupConLoss(0.1, "cosine")(torch.randn(4, 4), torch.tensor([0, 1, 0, 1]))
