"""GPU, NCCL, world_size 2 ("N-GPU result vs 1-GPU result vs oracle"; VERDICT r01 missing #3):
each rank holds a row shard of the batch, ShardedSupConLoss computes the loss of the CONCATENATED batch and
d loss / d z_local on real kernels with real collectives, and both are compared with the CPU oracle.  Also one
epoch of stage1.train_one_epoch with ShardedSupConLoss + GradSync against the single-process
trajectory on the merged batches.

Needs >= 2 visible GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py
-m gpu`; the log of that run is committed under profiles/."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < WORLD:
        pytest.skip(f"needs {WORLD} GPUs (have {torch.cuda.device_count() if torch.cuda.is_available() else 0})")


CASES = [
    # name, n, dtype, sim, tau, lam, k, alpha, kind, classes        (n_local = n / 2)
    ("tc_cosine", 2048, "bf16", "cosine", 0.07, 0.0, 15, 0.0, "iso", 2),          # two-phase fwd AND bwd
    ("tc_cosine_mined", 2048, "bf16", "cosine", 0.07, 0.0, 15, 0.5, "ties", 3),
    ("tc_geodesic_uniformity", 1024, "bf16", "geodesic", 0.1, 0.05, 7, 0.37, "clustered", 3),   # bwd single phase
    ("exact_fp32", 512, "f32", "geodesic", 0.07, 0.05, 15, 0.5, "iso", 2),        # FFMA kernels, 1e-5
    ("exact_fp32_ragged_block", 600, "f32", "cosine", 0.07, 0.0, 15, 1.0, "iso", 4),
]


def _rank_loss_and_grad(rank, world, port, out):
    import torch.distributed as dist
    from oracle import supcon_oracle as O
    from wav2vec_contr_loss_b200.distributed import ShardedSupConLoss
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        res = {}
        for (name, n, dt, sim, tau, lam, k, alpha, kind, classes), exchange in [(c, e) for e in ("nccl", "peer")
                                                                              for c in CASES]:
            name = f"{name}/{exchange}"
            dtype = torch.bfloat16 if dt == "bf16" else torch.float32
            x, y = O.make_inputs(n, 256, kind, classes=classes)
            z = F.normalize(x, dim=1).to(dtype)
            nl = n // world
            zl = z[rank * nl:(rank + 1) * nl].to(dev).requires_grad_(True)
            yl = y[rank * nl:(rank + 1) * nl].to(dev)
            mod = ShardedSupConLoss(tau, sim, lam, 2.0, exchange=exchange)
            mod.assume_unit_rows = True
            loss = mod(zl, yl, topk_neg=k, alpha=alpha)
            (1.5 * loss).backward()
            torch.cuda.synchronize()
            ref = O.closed_form(z.float(), y, temperature=tau, similarity=sim, uniformity_weight=lam, topk_neg=k,
                                alpha=alpha)
            want = 1.5 * ref["dz"][rank * nl:(rank + 1) * nl]
            err = float((zl.grad.double().cpu() - want).norm() / want.norm())
            res[name] = (float(loss), ref["loss"], err)
            # second call on the same module (graph-free path re-entered, comm stream reused) must agree bitwise
            zl2 = zl.detach().clone().requires_grad_(True)
            loss2 = mod(zl2, yl, topk_neg=k, alpha=alpha)
            (1.5 * loss2).backward()
            torch.cuda.synchronize()
            assert float(loss2) == float(loss) and torch.equal(zl2.grad, zl.grad), name
        # no gradient wanted: forward only, loss identical; then the module keeps working (peer: step closed)
        for exchange in ("nccl", "peer"):
            name, n, dt, sim, tau, lam, k, alpha, kind, classes = CASES[0]
            x, y = O.make_inputs(n, 256, kind, classes=classes)
            z = F.normalize(x, dim=1).to(torch.bfloat16)
            nl = n // world
            mod = ShardedSupConLoss(tau, sim, lam, 2.0, exchange=exchange)
            mod.assume_unit_rows = True
            zl, yl = z[rank * nl:(rank + 1) * nl].to(dev), y[rank * nl:(rank + 1) * nl].to(dev)
            with torch.no_grad():
                l0 = mod(zl, yl, topk_neg=k, alpha=alpha)
                l1 = mod(zl, yl, topk_neg=k, alpha=alpha)
            zg = zl.clone().requires_grad_(True)
            l2 = mod(zg, yl, topk_neg=k, alpha=alpha)
            if exchange == "peer":      # a second forward before the pending backward is refused, not silently wrong
                try:
                    mod(zg, yl, topk_neg=k, alpha=alpha)
                    res["peer_guard"] = "no error"
                except RuntimeError as exc:
                    res["peer_guard"] = str(exc)
            l2.backward()
            torch.cuda.synchronize()
            assert float(l0) == float(l1) == float(l2)
            res[f"no_grad/{exchange}"] = (float(l0), res[f"{name}/{exchange}"][0], 0.0)
        # ragged shards raise on every rank instead of hanging the collective
        try:
            m = 64 + 16 * rank
            ShardedSupConLoss(0.07, "cosine")(torch.randn(m, 32, device=dev), torch.zeros(m, device=dev))
            res["ragged"] = "no error"
        except ValueError as exc:
            res["ragged"] = str(exc)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_ranks_nccl_loss_and_gradient_vs_oracle():
    _need_two_gpus()
    out = mp.Manager().dict()
    mp.spawn(_rank_loss_and_grad, args=(WORLD, _free_port(), out), nprocs=WORLD, join=True)
    for rank in range(WORLD):
        res = out[rank]
        for exchange in ("nccl", "peer"):
            for (name, n, dt, *_rest) in CASES:
                loss, want, err = res[f"{name}/{exchange}"]
                tol = 2e-3 if dt == "bf16" else 1e-5
                assert loss == pytest.approx(want, rel=tol), (rank, name, exchange)
                assert err < (3 * tol if dt == "bf16" else tol), (rank, name, exchange, err)   # bf16 dz rounded by autograd
            assert res[f"no_grad/{exchange}"][0] == res[f"no_grad/{exchange}"][1]
        assert "before the backward" in res["peer_guard"]
        assert "same local batch size" in res["ragged"]
    for (name, *_r) in CASES:       # identical scalar on both ranks; the two exchanges differ only in summation order
        assert out[0][f"{name}/nccl"][0] == out[1][f"{name}/nccl"][0], name
        assert out[0][f"{name}/peer"][0] == out[1][f"{name}/peer"][0], name
        assert out[0][f"{name}/peer"][0] == pytest.approx(out[0][f"{name}/nccl"][0], rel=1e-6), name


# ------------------------------------------------------------------------------------------------------------
# stage1.train_one_epoch over two ranks (equal-step sampler + ShardedSupConLoss + GradSync) on the GPUs
# ------------------------------------------------------------------------------------------------------------
N_LOCAL = 8
_CFG = dict(finetune_encoder=False, use_rawboost=False, topk_neg=3, warmup_epochs=0, alpha_ramp_epochs=2, alpha_end=1.0)


class _Identity(torch.nn.Module):
    def forward(self, x, attention_mask=None):
        return x


def _problem():
    from oracle.gen_host_golden import TinyHead
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    head = TinyHead(feat=6, dim=8)
    labels = [int(v) for v in (torch.rand(70, generator=g) < 0.45)]
    feats = torch.randn(70, 2, 6, 5, generator=g) + torch.tensor(labels).view(-1, 1, 1, 1) * 0.8
    dataset = SimpleNamespace(data=[(f"utt{i}", lab) for i, lab in enumerate(labels)])
    return head, feats, torch.tensor(labels), dataset


def _loader(sampler, feats, labels):
    return [(feats[idx], labels[idx]) for idx in map(torch.tensor, sampler)]


def _rank_train(rank, world, port, out, exchange="nccl"):
    import torch.distributed as dist
    from wav2vec_contr_loss_b200 import stage1 as S
    from wav2vec_contr_loss_b200.distributed import ShardedSupConLoss
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    assert S.setup_distributed(backend="nccl") == (True, rank, world, rank)
    try:
        dev = torch.device("cuda", rank)
        head, feats, labels, dataset = _problem()
        head = head.to(dev)
        sampler = S.BalancedBatchSampler(dataset, N_LOCAL, seed=5, rank=rank, world_size=world)
        loss_fn = ShardedSupConLoss(0.2, "cosine", 0.05, 2.0, exchange=exchange)
        opt = torch.optim.SGD(head.parameters(), lr=0.5)
        history = []
        for epoch in (1, 2):
            sampler.set_epoch(epoch)
            loader = _loader(sampler, feats, labels)
            avg, alpha = S.train_one_epoch(_Identity(), head, loss_fn, loader, opt, dev, epoch, SimpleNamespace(**_CFG),
                                           grad_sync=S.GradSync(head.parameters()))
            history.append((len(loader), avg, alpha))
        out[rank] = (history, torch.cat([p.detach().reshape(-1) for p in head.parameters()]).double().cpu())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_two_ranks_nccl_train_like_one_process_on_the_global_batch(cuda_device, exchange):
    _need_two_gpus()
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    out = mp.Manager().dict()
    mp.spawn(_rank_train, args=(WORLD, _free_port(), out, exchange), nprocs=WORLD, join=True)

    head, feats, labels, dataset = _problem()
    head = head.to(cuda_device)
    sampler = S.BalancedBatchSampler(dataset, N_LOCAL, seed=5)           # the same batch sequence, one process
    opt = torch.optim.SGD(head.parameters(), lr=0.5)
    want = []
    for epoch in (1, 2):
        sampler.set_epoch(epoch)
        batches = list(sampler)
        steps = len(batches) // WORLD
        merged = [sum((batches[s * WORLD + r] for r in range(WORLD)), []) for s in range(steps)]
        avg, alpha = S.train_one_epoch(_Identity(), head, SupConBinaryLoss(0.2, "cosine", 0.05, 2.0),
                                       _loader(merged, feats, labels), opt, cuda_device, epoch, SimpleNamespace(**_CFG))
        want.append((steps, avg, alpha))
    want_w = torch.cat([p.detach().reshape(-1) for p in head.parameters()]).double().cpu()
    for rank in range(WORLD):
        history, got_w = out[rank]
        for (steps, avg, alpha), (w_steps, w_avg, w_alpha) in zip(history, want):
            assert steps == w_steps and alpha == w_alpha and avg == pytest.approx(w_avg, rel=1e-5)
        assert torch.allclose(got_w, want_w, rtol=1e-4, atol=1e-6)
    assert torch.equal(out[0][1], out[1][1])                              # replicas stay identical
