"""GPU: the Stage-1 host functions (wav2vec_contr_loss_b200.stage1) with the product defaults - the CUDA
normalise kernel and the CUDA loss - against the fixture made by running the reference's stage1_utils.py +
loss.py (tests/golden/stage1_host.json), and the CUDA-graph step against the same step launched eagerly."""
import copy
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import supcon_oracle as O
from oracle.gen_host_golden import STEP_CASES, TinyHead, tiny_stage1

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage1_host.json")))


def _flat(module):
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()]).double().cpu()


@pytest.mark.parametrize("name,sim,tau,lam,finetune", STEP_CASES, ids=[c[0] for c in STEP_CASES])
def test_epochs_equal_reference_run(cuda_device, name, sim, tau, lam, finetune):
    """train_one_epoch + evaluate, three epochs (alpha 0 -> 0.5 -> 1), fp32: same mean losses and same trained
    weights as the reference's own functions with its own loss on the CPU."""
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    want = GOLD["epochs"][name]
    enc, head, train, dev, cfg = tiny_stage1(finetune)
    enc, head = enc.to(cuda_device), head.to(cuda_device)
    params = list(head.parameters()) + (list(enc.parameters()) if finetune else [])
    opt = torch.optim.AdamW(params, lr=5e-3, weight_decay=3e-3)
    loss_fn = SupConBinaryLoss(temperature=tau, similarity=sim, uniformity_weight=lam, uniformity_t=2.0)
    for i, epoch in enumerate((1, 2, 3)):
        tl, alpha = S.train_one_epoch(enc, head, loss_fn, train, opt, cuda_device, epoch, cfg)
        dl = S.evaluate(enc, head, loss_fn, dev, cuda_device, cfg)
        assert alpha == want["alpha"][i]
        assert tl == pytest.approx(want["train"][i], rel=1e-4)
        assert dl == pytest.approx(want["dev"][i], rel=1e-4)
    assert torch.allclose(_flat(head), torch.tensor(want["head"], dtype=torch.float64), rtol=1e-3, atol=5e-4)
    assert torch.allclose(_flat(enc), torch.tensor(want["encoder"], dtype=torch.float64), rtol=1e-3, atol=5e-4)


def test_normalized_loss_entry(cuda_device):
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    x, y = O.make_inputs(96, 256, "clustered")
    xg = (2.5 * x).to(cuda_device).requires_grad_(True)
    loss = S.normalized_supcon_loss(xg, y.to(cuda_device), SupConBinaryLoss(0.07, "cosine"), topk_neg=15, alpha=0.5)
    loss.backward()
    xr = (2.5 * x).double().requires_grad_(True)
    ref = O.anchor_loop_loss(F.normalize(xr, dim=1), y, temperature=0.07, similarity="cosine", topk_neg=15, alpha=0.5)
    ref.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    err = (xg.grad.double().cpu() - xr.grad).norm() / xr.grad.norm()
    assert err < 1e-5


def test_export_embeddings_on_gpu(cuda_device, tmp_path):
    from wav2vec_contr_loss_b200 import stage1 as S
    enc, head, _, dev, _ = tiny_stage1(False)
    with torch.no_grad():
        want = torch.cat([F.normalize(head.eval()(enc.eval()(w, attention_mask=(w != 0).long())).mean(-1), dim=1)
                          for w, *_ in dev])
    enc, head = enc.to(cuda_device), head.to(cuda_device)
    emb, lab, n = S.export_embeddings(enc, head, dev, cuda_device, str(tmp_path), "dev")
    z, y = np.load(emb), np.load(lab)
    assert n == 20 and z.dtype == np.float32 and z.shape == (20, 8)
    assert np.allclose(z, want.numpy(), rtol=1e-5, atol=1e-6)
    assert y.tolist() == torch.cat([b[1] for b in dev]).tolist()


def _graph_problem(device, batch, seed=5):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    head = TinyHead(feat=32, dim=16).to(device)
    hs = torch.randn(5, batch, 3, 32, 10, generator=g)
    y = torch.stack([torch.randperm(batch, generator=g) % 2 for _ in range(5)])
    hs = hs + 0.5 * (2.0 * y.view(5, batch, 1, 1, 1) - 1.0)
    return head, hs.to(device), y.to(device)


@pytest.mark.parametrize("batch,sim,lam", [(64, "cosine", 0.0), (64, "geodesic", 0.1), (512, "cosine", 0.05)])
def test_graphed_step_equals_eager_steps(cuda_device, batch, sim, lam):
    """one CUDA-graph replay per step (head -> normalise -> loss -> backward -> clip -> AdamW) walks the same
    trajectory as the eager step; building/capturing it does not train."""
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    head_e, hs, y = _graph_problem(cuda_device, batch)
    head_g = copy.deepcopy(head_e)
    loss_fn = SupConBinaryLoss(temperature=0.1, similarity=sim, uniformity_weight=lam)
    alphas = [0.0, 0.0, 0.5, 0.5, 0.0]

    opt_e = torch.optim.AdamW(head_e.parameters(), lr=5e-3, weight_decay=3e-3, capturable=True)
    eager = []
    for s, alpha in enumerate(alphas):
        z = S.embed(head_e, hs[s])
        loss = loss_fn(z, y[s], topk_neg=7, alpha=alpha)
        opt_e.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(head_e.parameters(), 5.0)
        opt_e.step()
        eager.append(float(loss))

    opt_g = torch.optim.AdamW(head_g.parameters(), lr=5e-3, weight_decay=3e-3, capturable=True)
    step = S.GraphedHeadStep(head_g, loss_fn, opt_g, hs[0], y[0], topk_neg=7)
    start = _flat(head_g)
    graphed = []
    for s, alpha in enumerate(alphas):
        if s == 0:
            step._graphs[float(alpha)] = step._capture(float(alpha))
            assert torch.equal(_flat(head_g), start)                 # capture (with its warm-up steps) is undone
        graphed.append(float(step(hs[s], y[s], alpha)))
    assert len(step._graphs) == 2                                     # one graph per distinct alpha
    assert graphed == pytest.approx(eager, rel=1e-5)
    assert torch.allclose(_flat(head_g), _flat(head_e), rtol=1e-4, atol=1e-6)


def test_graphed_step_argument_checks(cuda_device):
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    head, hs, y = _graph_problem(cuda_device, 16)
    loss_fn = SupConBinaryLoss(0.1, "cosine")
    with pytest.raises(ValueError, match="capturable"):
        S.GraphedHeadStep(head, loss_fn, torch.optim.AdamW(head.parameters()), hs[0], y[0])
    with pytest.raises(RuntimeError, match="CUDA"):
        S.GraphedHeadStep(head, loss_fn, torch.optim.AdamW(head.parameters(), capturable=True), hs[0].cpu(), y[0].cpu())
    step = S.GraphedHeadStep(head, loss_fn, torch.optim.AdamW(head.parameters(), capturable=True), hs[0], y[0])
    with pytest.raises(ValueError, match="captured for"):
        step(hs[0][:8], y[0][:8], 0.0)
