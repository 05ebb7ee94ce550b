"""GPU: the fused compression head (wav2vec_contr_loss_b200.head, csrc/supcon_head.cu) against the reference's
op chain written in plain torch fp32 (compression_module.py:48-65 + stage1_utils.py:122-123), and inside the
Stage-1 epoch functions against the reference's own run (tests/golden/stage1_host.json).  Tolerances: the fused
form reorders fp32 sums (time mean before the Linear layer), so 2e-5 relative on values and gradients."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle.gen_host_golden import STEP_CASES, tiny_stage1

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage1_host.json")))
TOL = 2e-5

SHAPES = [          # batch, layers, feat, frames
    (3, 25, 256, 199),   # the reference's K and T; 128-bit loads
    (2, 3, 7, 13),       # nothing aligned: scalar loads, one short block
    (2, 2, 18, 199),     # ragged last block of rows -> scalar loads
    (2, 5, 40, 50),
    (1, 2, 6, 3000),     # long clip: 4 rows per block
]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def reference_chain(hs, weight, bias, mask=None, slope=0.01):
    """the reference's ops in the reference's order; mask = dropout multiplier (B, F, T) or None (eval)"""
    x = hs.mean(dim=1)
    if mask is not None:
        x = x * mask
    x = F.leaky_relu(x, slope)
    seq = F.linear(x.transpose(1, 2), weight, bias).transpose(1, 2)
    pooled = seq.mean(dim=-1)
    return x.mean(dim=-1), pooled, F.normalize(pooled, p=2, dim=1)


def make_head(feat, dim, p, device, seed=0):
    from wav2vec_contr_loss_b200 import FusedCompressionHead
    torch.manual_seed(seed)
    return FusedCompressionHead(feat, dim, p).to(device)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_eval_forward_and_backward_equal_reference_ops(cuda_device, shape):
    b, k, f, t = shape
    head = make_head(f, 16, 0.1, cuda_device).eval()
    hs = torch.randn(b, k, f, t, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    w = torch.randn(b, 16, generator=torch.Generator().manual_seed(2)).to(cuda_device)

    hs_a = hs.clone().requires_grad_(True)
    feats = head.pooled_features(hs_a)
    seq1 = head(hs_a)
    z = head.embed(hs_a)
    assert seq1.shape == (b, 16, 1)
    (z * w).sum().backward()
    grads = [hs_a.grad.clone(), head.mlp3.weight.grad.clone(), head.mlp3.bias.grad.clone()]
    head.zero_grad()

    hs_r = hs.clone().requires_grad_(True)
    feats_r, pooled_r, z_r = reference_chain(hs_r, head.mlp3.weight, head.mlp3.bias)
    (z_r * w).sum().backward()
    assert rel(feats, feats_r) < TOL
    assert rel(seq1.mean(dim=-1), pooled_r) < TOL
    assert rel(z, z_r) < TOL
    assert rel(grads[0], hs_r.grad) < TOL
    assert rel(grads[1], head.mlp3.weight.grad) < TOL
    assert rel(grads[2], head.mlp3.bias.grad) < TOL


def test_frozen_encoder_saves_nothing_and_bf16_out(cuda_device):
    head = make_head(64, 256, 0.0, cuda_device).train()
    hs = torch.randn(4, 3, 64, 21, device=cuda_device)
    z = head.embed(hs, out_dtype=torch.bfloat16)
    assert z.dtype == torch.bfloat16 and z.shape == (4, 256)
    _, _, z_r = reference_chain(hs, head.mlp3.weight, head.mlp3.bias)
    assert rel(z.float(), z_r) < 4e-3
    z.float().sum().backward()                       # reaches mlp3 only
    assert head.mlp3.weight.grad is not None


def test_train_mode_dropout_is_consistent_and_calibrated(cuda_device):
    """the mask is recovered from the backward kernel (positive inputs, unit upstream gradient) and fed to the
    reference op chain: the forward value and every gradient must agree for THAT mask; keep rate ~ 1 - p."""
    from wav2vec_contr_loss_b200.head import layer_time_pool
    b, k, f, t, p = 3, 4, 64, 199, 0.1
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=cuda_device)
    pos = torch.rand(b, k, f, t, device=cuda_device).add_(0.5).requires_grad_(True)
    layer_time_pool(pos, p, 0.01, rng).sum().backward()
    mask = pos.grad[:, 0] * (k * t)                  # = dropout multiplier: 0 or 1/(1-p)
    assert torch.equal(pos.grad[:, 0], pos.grad[:, k - 1])
    kept = mask > 0
    assert torch.allclose(mask[kept], torch.full_like(mask[kept], 1 / (1 - p)), rtol=1e-5)
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    assert abs(float(kept.float().mean(dim=(0, 2)).min()) - (1 - p)) < 0.08      # no dead / always-on rows
    # same state -> same mask; next offset -> a different one
    again = torch.rand(b, k, f, t, device=cuda_device).add_(0.5).requires_grad_(True)
    layer_time_pool(again, p, 0.01, rng).sum().backward()
    assert torch.equal(again.grad[:, 0] * (k * t) > 0, kept)
    other = torch.rand(b, k, f, t, device=cuda_device).add_(0.5).requires_grad_(True)
    layer_time_pool(other, p, 0.01, rng + torch.tensor([0, 1], device=cuda_device)).sum().backward()
    assert float(((other.grad[:, 0] > 0) != kept).float().mean()) > 0.1

    hs = torch.randn(b, k, f, t, device=cuda_device)
    up = torch.randn(b, f, device=cuda_device)
    hs_a = hs.clone().requires_grad_(True)
    feats = layer_time_pool(hs_a, p, 0.01, rng)
    (feats * up).sum().backward()
    hs_r = hs.clone().requires_grad_(True)
    feats_r = F.leaky_relu(hs_r.mean(dim=1) * mask, 0.01).mean(dim=-1)
    (feats_r * up).sum().backward()
    assert rel(feats, feats_r) < TOL
    assert rel(hs_a.grad, hs_r.grad) < TOL


def test_module_advances_its_dropout_stream(cuda_device):
    head = make_head(32, 8, 0.25, cuda_device).train()
    hs = torch.randn(2, 3, 32, 40, device=cuda_device)
    a = head.pooled_features(hs)
    b = head.pooled_features(hs)
    assert head.rng_state.tolist()[1] == 2 and not torch.equal(a, b)
    head.eval()
    c = head.pooled_features(hs)
    d = head.pooled_features(hs)
    assert head.rng_state.tolist()[1] == 2 and torch.equal(c, d)


@pytest.mark.parametrize("name,sim,tau,lam,finetune", STEP_CASES, ids=[c[0] for c in STEP_CASES])
def test_epochs_with_fused_head_equal_reference_run(cuda_device, name, sim, tau, lam, finetune):
    """the reference's three-epoch run (its head, its loss, CPU) reproduced with the fused head kernel + the CUDA
    loss: same mean losses, same trained weights (the fine-tuned case runs the head's backward kernel)."""
    from wav2vec_contr_loss_b200 import FusedCompressionHead, SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    want = GOLD["epochs"][name]
    enc, tiny_head, train, dev, cfg = tiny_stage1(finetune)
    head = FusedCompressionHead(12, 8, 0.0)
    with torch.no_grad():
        head.mlp3.weight.copy_(tiny_head.fc.weight)
        head.mlp3.bias.copy_(tiny_head.fc.bias)
    enc, head = enc.to(cuda_device), head.to(cuda_device)
    params = list(head.parameters()) + (list(enc.parameters()) if finetune else [])
    opt = torch.optim.AdamW(params, lr=5e-3, weight_decay=3e-3)
    loss_fn = SupConBinaryLoss(temperature=tau, similarity=sim, uniformity_weight=lam, uniformity_t=2.0)
    for i, epoch in enumerate((1, 2, 3)):
        tl, alpha = S.train_one_epoch(enc, head, loss_fn, train, opt, cuda_device, epoch, cfg)
        dl = S.evaluate(enc, head, loss_fn, dev, cuda_device, cfg)
        assert tl == pytest.approx(want["train"][i], rel=1e-4)
        assert dl == pytest.approx(want["dev"][i], rel=1e-4)
    flat = torch.cat([p.detach().reshape(-1) for p in head.parameters()]).double().cpu()
    assert torch.allclose(flat, torch.tensor(want["head"], dtype=torch.float64), rtol=1e-3, atol=5e-4)
    flat_e = torch.cat([p.detach().reshape(-1) for p in enc.parameters()]).double().cpu()
    assert torch.allclose(flat_e, torch.tensor(want["encoder"], dtype=torch.float64), rtol=1e-3, atol=5e-4)


def test_graphed_step_with_fused_head_and_dropout(cuda_device):
    """one CUDA-graph replay per step with the fused head in train mode: every replay draws a new dropout mask
    (the offset lives on the device), construction leaves weights and stream untouched, the loss goes down
    towards its floor."""
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200 import stage1 as S
    head = make_head(64, 32, 0.1, cuda_device).train()
    g = torch.Generator().manual_seed(3)
    y = (torch.randperm(64, generator=g) % 2).to(cuda_device)
    hs = (torch.randn(64, 3, 64, 20, generator=g).to(cuda_device) + 0.6 * (2.0 * y.view(-1, 1, 1, 1) - 1.0))
    opt = torch.optim.AdamW(head.parameters(), lr=1e-2, capturable=True)
    # an eager twin: same weights, same dropout stream (seed and offset), same optimizer
    import copy
    head_e = copy.deepcopy(head)
    opt_e = torch.optim.AdamW(head_e.parameters(), lr=1e-2, capturable=True)
    twin = S.GraphedHeadStep(head_e, SupConBinaryLoss(0.1, "cosine"), opt_e, hs, y, topk_neg=7)
    step = S.GraphedHeadStep(head, SupConBinaryLoss(0.1, "cosine"), opt, hs, y, topk_neg=7)
    before = torch.cat([p.detach().reshape(-1) for p in head.parameters()]).clone()
    step._graphs[0.0] = step._capture(0.0)
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in head.parameters()]), before)
    assert head.rng_state.tolist()[1] == 0
    losses = [float(step(hs, y, 0.0)) for _ in range(25)]
    assert head.rng_state.tolist()[1] == 25
    # the classes are well separated, so the loss starts close to its floor log(|pos|) = log 31 = 3.434
    assert all(l == l for l in losses) and sum(losses[-5:]) < sum(losses[:5]) - 0.01
    assert min(losses) > 3.43
    # the 25 replays walked the trajectory of 25 eager steps with the same dropout offsets: a replay that stopped
    # updating the weights, or drew the same mask every time, would not
    eager = [float(twin._eager(0.0)) for _ in range(25)]
    assert head_e.rng_state.tolist() == head.rng_state.tolist()
    assert losses == pytest.approx(eager, rel=1e-4)
    flat_g = torch.cat([p.detach().reshape(-1) for p in head.parameters()])
    flat_e = torch.cat([p.detach().reshape(-1) for p in head_e.parameters()])
    assert not torch.equal(flat_g, before)
    assert torch.allclose(flat_g, flat_e, rtol=1e-3, atol=1e-5)
