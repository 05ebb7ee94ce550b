"""GPU: the CUDA path (through the C-ABI) against the oracle and the committed
reference outputs.  Tolerances (north_star): loss and dz within 1e-5 relative
on the fp32 path, 2e-3 on the bf16-input path; hard-negative index sets exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_names, load_golden
from oracle import supcon_oracle as O
import gpu_util as G

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5      # north_star: fp32 / tf32-off path
TOL_BF16 = 2e-3     # north_star: bf16 input, fp32 accumulate


@pytest.mark.parametrize("flags", [0, 4])     # 0: small batches take the single-launch cluster kernel; 4: tiled FFMA kernels
@pytest.mark.parametrize("name", golden_names())
def test_golden_fixtures_fp32(cuda_device, name, flags):
    meta, g = load_golden(name)
    z, y = torch.from_numpy(g["z"]), torch.from_numpy(g["labels"])
    kw = dict(tau=meta["tau"], similarity=meta["similarity"], lam=meta["lambda_uni"], t=meta["uni_t"],
              topk=meta["topk"], alpha=meta["alpha"])
    loss, dz = G.kernel_loss_and_grad(z, y, flags=flags, **kw)
    assert loss == pytest.approx(float(g["loss64"]), rel=TOL_F32, abs=1e-6)
    assert loss == pytest.approx(float(g["loss32"]), rel=TOL_F32, abs=1e-6)
    if meta["n"] < 2:
        return
    ref64 = torch.from_numpy(g["dz64"])
    if meta["kind"] == "ties":      # reference tie order is unspecified: use the lowest-index oracle
        ref64 = G.oracle_for(z, y, **kw)["dz"]
    if float(ref64.norm()) == 0.0:
        assert float(dz.norm()) == 0.0
    else:
        assert G.rel_err(dz, ref64) < TOL_F32


CASES = [
    # n, d, kind, classes, sim, tau, lam, t, K, alpha
    (64, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0),          # C1
    (64, 256, "iso", 2, "geodesic", 0.07, 0.05, 2.0, 15, 0.0),       # C2
    (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0),        # C3 alpha schedule
    (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.0125),
    (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 0.5),
    (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 15, 1.0),
    (1024, 256, "clustered", 2, "cosine", 0.07, 0.0, 2.0, 32, 0.5),
    (1024, 256, "iso", 2, "cosine", 0.07, 0.0, 2.0, 600, 1.0),       # K >= N/2: all negatives
    (1000, 200, "iso", 5, "geodesic", 0.1, 0.2, 2.0, 15, 0.37),      # ragged N, d; multi-class
    (257, 48, "clustered", 3, "geodesic", 0.2, 0.1, 3.0, 8, 1.0),
    (130, 19, "iso", 7, "cosine", 0.07, 0.1, 2.0, 32, 0.37),         # odd d -> scalar loads
    (2, 4, "iso", 1, "cosine", 0.5, 0.5, 2.0, 3, 0.5),
    (3, 5, "iso", 2, "geodesic", 0.5, 0.5, 2.0, 3, 0.5),
]


@pytest.mark.parametrize("n,d,kind,classes,sim,tau,lam,t,k,alpha", CASES)
def test_seeded_cases_fp32(cuda_device, n, d, kind, classes, sim, tau, lam, t, k, alpha):
    x, y = O.make_inputs(n, d, kind, classes=max(classes, 2))
    if classes == 1:
        y = torch.zeros_like(y)
    z = F.normalize(x, dim=1)
    kw = dict(tau=tau, similarity=sim, lam=lam, t=t, topk=k, alpha=alpha)
    loss, dz = G.kernel_loss_and_grad(z, y, **kw)
    ref = G.oracle_for(z, y, **kw)
    assert loss == pytest.approx(ref["loss"], rel=TOL_F32, abs=1e-6)
    assert G.rel_err(dz, ref["dz"]) < TOL_F32


@pytest.mark.parametrize("sim", ["cosine", "geodesic"])
def test_bf16_inputs(cuda_device, sim):
    x, y = O.make_inputs(512, 256, "iso")
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=0.07, similarity=sim, lam=0.05, topk=15, alpha=0.5)
    loss, dz = G.kernel_loss_and_grad(zb, y, dtype=torch.bfloat16, unit_rows=True, **kw)
    ref = G.oracle_for(zb.float(), y, **kw)      # oracle semantics for bf16: loss.py on z_bf16.float()
    assert loss == pytest.approx(ref["loss"], rel=TOL_BF16)
    assert G.rel_err(dz, ref["dz"]) < 2 * TOL_BF16     # dz itself is rounded to bf16 on return


def _exact_arith_inputs(n, d, seed):
    """Entries in {0, +-1/2, +-1/4}: every dot product is exact in fp32 and fp64,
    so hard-negative sets (with their many exact ties) must match bit-for-bit."""
    g = torch.Generator().manual_seed(seed)
    vals = torch.tensor([0.0, 0.5, -0.5, 0.25, -0.25])
    z = vals[torch.randint(0, 5, (n, d), generator=g)]
    y = torch.randint(0, 3, (n,), generator=g)
    return z, y


@pytest.mark.parametrize("n,d,k", [(96, 8, 7), (300, 12, 15), (1024, 16, 32)])
def test_hard_negative_index_sets_exact(cuda_device, n, d, k):
    z, y = _exact_arith_inputs(n, d, seed=n)
    out = G.kernel_stats(z, y, tau=0.5, similarity="cosine", topk=k, alpha=1.0, flags=4)
    ref = G.oracle_for(z, y, tau=0.5, similarity="cosine", topk=k, alpha=1.0, want_topk_idx=True, want_grad=False)
    got = out["idx"].cpu().tolist()
    for i in range(n):
        mine = [j for j in got[i] if j >= 0]
        assert mine == sorted(ref["stats"]["topk_idx"][i]), f"row {i}"
    # and the gradient that depends on those sets
    full = G.oracle_for(z, y, tau=0.5, similarity="cosine", topk=k, alpha=1.0)
    for flags in (0, 4):     # the cluster kernel (n <= 160) must pick the same sets: same gradient
        loss, dz = G.kernel_loss_and_grad(z, y, tau=0.5, similarity="cosine", topk=k, alpha=1.0, flags=flags)
        assert loss == pytest.approx(full["loss"], rel=TOL_F32)
        assert G.rel_err(dz, full["dz"]) < TOL_F32
    if n <= 160:             # threshold (value, index) written by the cluster kernel == tiled kernel
        small = G.kernel_stats(z, y, tau=0.5, similarity="cosine", topk=k, alpha=1.0, flags=0)
        assert torch.equal(small["stats"].view(torch.int32)[:, 5].cpu(), out["stats"].view(torch.int32)[:, 5].cpu())
        assert torch.equal(small["stats"][:, 4].cpu(), out["stats"][:, 4].cpu())


def test_appendix_b_tie_case(cuda_device):
    z = F.normalize(torch.tensor([[1, 0], [1, 0], [0, 1], [0, 1], [0, 1], [.6, .8]]), dim=1)
    y = torch.tensor([1, 1, 0, 0, 0, 0])
    loss, dz = G.kernel_loss_and_grad(z, y, tau=0.5, similarity="cosine", topk=2, alpha=1.0)
    assert loss == pytest.approx(1.0041676759719849, rel=TOL_F32)
    for flags in (0, 4):
        out = G.kernel_stats(z, y, tau=0.5, similarity="cosine", topk=2, alpha=1.0, flags=flags)
        assert out["idx"][0].tolist() == [2, 5] and out["idx"][1].tolist() == [2, 5]
        loss, dz = G.kernel_loss_and_grad(z, y, tau=0.5, similarity="cosine", topk=2, alpha=1.0, flags=flags)
        assert loss == pytest.approx(1.0041676759719849, rel=TOL_F32)


def test_analytic_known_answers(cuda_device):
    y = torch.tensor([1, 0] * 32)
    z = torch.ones(64, 8) / math.sqrt(8.0)
    for sim in ("cosine", "geodesic"):
        loss, _ = G.kernel_loss_and_grad(z, y, tau=0.07, similarity=sim, topk=15, alpha=0.0)
        assert loss == pytest.approx(math.log(63.0), rel=TOL_F32)
        loss, _ = G.kernel_loss_and_grad(z, y, tau=0.07, similarity=sim, topk=15, alpha=1.0)
        assert loss == pytest.approx(math.log(46.0), rel=TOL_F32)
    z = torch.eye(256)[:64]
    loss, _ = G.kernel_loss_and_grad(z, y, tau=0.07, similarity="cosine", lam=0.05, t=2.0)
    assert loss == pytest.approx(3.9431347536906003, rel=TOL_F32)


def test_unnormalised_and_small_temperature(cuda_device):
    """z is taken as given: large logits must not overflow (online max)."""
    x, y = O.make_inputs(200, 32, "iso", seed=5)
    z = 3.0 * F.normalize(x, dim=1)
    for tau in (0.07, 0.01):
        kw = dict(tau=tau, similarity="cosine", lam=0.0, topk=5, alpha=0.5)
        loss, dz = G.kernel_loss_and_grad(z, y, **kw)
        ref = G.oracle_for(z, y, **kw)
        assert math.isfinite(loss)
        assert loss == pytest.approx(ref["loss"], rel=TOL_F32)
        assert G.rel_err(dz, ref["dz"]) < 5 * TOL_F32   # logits ~ 9/tau: conditioning, SURVEY H7


def test_autograd_contract(cuda_device):
    from wav2vec_contr_loss_b200 import SupConBinaryLoss, SupConMultiClassLoss
    x, y = O.make_inputs(64, 32, "iso", classes=4)
    x = x.to(cuda_device).requires_grad_(True)
    z = F.normalize(x, p=2, dim=1)
    mod = SupConBinaryLoss(0.07, "cosine")
    loss = mod(z, y.to(cuda_device), topk_neg=15, alpha=0.0)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.device == z.device
    (2.5 * loss).backward()
    ref = G.oracle_for(F.normalize(x.detach().cpu(), dim=1), y, tau=0.07, similarity="cosine", topk=15)
    zz, nrm = O.normalize_fwd(x.detach().cpu().double())
    dx = O.normalize_bwd(zz, nrm, ref["dz"]) * 2.5
    assert G.rel_err(x.grad.cpu(), dx) < TOL_F32
    # forward-only under no_grad: no graph, same value
    with torch.no_grad():
        l2 = mod(z, y.to(cuda_device), topk_neg=15, alpha=0.0)
    assert not l2.requires_grad and float(l2) == float(loss)
    # multi-class class == binary class with cosine, alpha = 0
    l3 = SupConMultiClassLoss(0.07)(z.detach(), y.to(cuda_device))
    assert float(l3) == pytest.approx(float(loss), rel=1e-6)
    # float labels and (B,1) labels are accepted (loss.py:123)
    l4 = mod(z.detach(), y.to(cuda_device).float().view(-1, 1) * 0.5)
    assert float(l4) == pytest.approx(float(loss), rel=1e-6)


def test_multiclass_fixture(cuda_device):
    import os
    from conftest import GOLDEN
    from wav2vec_contr_loss_b200 import SupConMultiClassLoss
    f = np.load(os.path.join(GOLDEN, "multiclass_n64_d32.npz"))
    z = torch.from_numpy(f["z"]).to(cuda_device).requires_grad_(True)
    loss = SupConMultiClassLoss(0.1)(z, torch.from_numpy(f["labels"]).to(cuda_device))
    loss.backward()
    assert float(loss) == pytest.approx(float(f["loss64"]), rel=TOL_F32)
    assert G.rel_err(z.grad.cpu(), torch.from_numpy(f["dz64"])) < TOL_F32


def test_row_blocks_compose(cuda_device):
    """Two row blocks through the C-ABI == whole batch (multi-rank contract)."""
    from wav2vec_contr_loss_b200 import functional as Fn
    x, y = O.make_inputs(200, 64, "clustered", classes=3)
    z = F.normalize(x, dim=1)
    kw = dict(tau=0.1, similarity="geodesic", lam=0.1, t=2.0, topk=5, alpha=0.3)
    a = G.kernel_stats(z, y, row_offset=0, n_rows=120, **kw)
    b = G.kernel_stats(z, y, row_offset=120, n_rows=80, **kw)
    partials = a["partials"] + b["partials"]
    stats = torch.cat([a["stats"], b["stats"]])
    whole_prob = Fn.make_problem(200, 64, 0, tau=0.1, similarity=1, lambda_uni=0.1, uni_t=2.0, topk=5, alpha=0.3)
    loss = float(Fn.finalize(whole_prob, partials))
    ref = G.oracle_for(z, y, **kw)
    assert loss == pytest.approx(ref["loss"], rel=TOL_F32)
    dz_b = Fn.backward_rows(b["z"], b["y"], stats, partials, None, b["prob"])
    assert G.rel_err(dz_b.cpu(), ref["dz"][120:]) < TOL_F32


def test_normalize_kernels(cuda_device):
    from wav2vec_contr_loss_b200 import l2_normalize
    x = torch.randn(300, 256, generator=torch.Generator().manual_seed(0))
    x[5] = 0.0
    xg = x.to(cuda_device).requires_grad_(True)
    z = l2_normalize(xg)
    w = torch.randn(300, 256, generator=torch.Generator().manual_seed(1))
    (z * w.to(cuda_device)).sum().backward()
    xr = x.clone().requires_grad_(True)
    zr = F.normalize(xr, p=2, dim=1)
    (zr * w).sum().backward()
    assert torch.allclose(z.cpu(), zr, rtol=2e-6, atol=1e-7)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-6)


def test_large_batch_properties_fp32(cuda_device):
    """N = 4096: against the row-blocked fp64 oracle, plus size-independent
    properties (sum_i dz_i . z_i relation is not available; use: gradient of a
    label-permuted/row-permuted batch is the permuted gradient)."""
    x, y = O.make_inputs(4096, 256, "iso")
    z = F.normalize(x, dim=1)
    kw = dict(tau=0.07, similarity="cosine", lam=0.0, topk=15, alpha=0.5)
    loss, dz = G.kernel_loss_and_grad(z, y, **kw)
    ref = G.oracle_for(z, y, **kw)
    assert loss == pytest.approx(ref["loss"], rel=TOL_F32)
    assert G.rel_err(dz, ref["dz"]) < TOL_F32
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(9))
    loss_p, dz_p = G.kernel_loss_and_grad(z[perm], y[perm], **kw)
    assert loss_p == pytest.approx(loss, rel=1e-6)
    assert G.rel_err(dz_p, dz[perm]) < 1e-5


def test_training_step_reduces_loss(cuda_device):
    """The caller's shape (stage1_utils.py:122-130): head -> mean -> normalize -> loss -> backward -> step."""
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    torch.manual_seed(0)
    head = torch.nn.Linear(64, 32).to(cuda_device)
    opt = torch.optim.AdamW(head.parameters(), lr=1e-2)
    loss_fn = SupConBinaryLoss(temperature=0.07, similarity="geodesic", uniformity_weight=0.05)
    feats = torch.randn(64, 64, 10, device=cuda_device)
    labels = (torch.arange(64, device=cuda_device) % 2).long()
    feats[labels == 1, :8] += 1.0
    first = last = None
    for step in range(30):
        seq = head(feats.transpose(1, 2)).transpose(1, 2)
        z = F.normalize(seq.mean(dim=-1), p=2, dim=1)
        loss = loss_fn(z, labels, topk_neg=15, alpha=0.3)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(head.parameters(), 5.0)
        opt.step()
        first = float(loss) if first is None else first
        last = float(loss)
    assert last < first - 0.05


# ---------------------------------------------------------------------------
# bf16 tensor-core path (tcgen05 / TMEM / TMA).  Oracle semantics for bf16:
# the reference algorithm applied to z_bf16.float(); tolerance 2e-3 (north_star).
# ---------------------------------------------------------------------------
TC_CASES = [
    # n, kind, classes, sim, tau, lam
    (512, "iso", 2, "cosine", 0.07, 0.0),
    (1000, "iso", 3, "cosine", 0.07, 0.0),          # ragged N: masked tail tile
    (640, "clustered", 2, "cosine", 0.1, 0.05),
    (512, "iso", 2, "geodesic", 0.07, 0.0),
    (777, "iso", 5, "geodesic", 0.1, 0.2),
    (2048, "iso", 2, "cosine", 0.03, 0.0),
]


@pytest.mark.parametrize("n,kind,classes,sim,tau,lam", TC_CASES)
@pytest.mark.parametrize("flags", [2, 2 | 8, 2 | 16])   # tensor path both ways / forward only / backward only
def test_tensor_core_path_bf16(cuda_device, n, kind, classes, sim, tau, lam, flags):
    from wav2vec_contr_loss_b200 import functional as Fn
    x, y = O.make_inputs(n, 256, kind, classes=classes)
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=tau, similarity=sim, lam=lam, t=2.0, topk=15, alpha=0.0)
    ref = G.oracle_for(zb.float(), y, **kw)
    out = G.kernel_stats(zb, y, dtype=torch.bfloat16, flags=flags, **kw)
    assert float(out["loss"]) == pytest.approx(ref["loss"], rel=TOL_BF16)
    st = out["stats"].cpu()
    assert float((st[:, 0].double() - ref["stats"]["lse"]).abs().max()) < 1e-4
    assert bool((st.view(torch.int32)[:, 2].long() == ref["stats"]["npos"]).all())
    dz = Fn.backward_rows(out["z"], out["y"], out["stats"], out["partials"], None, out["prob"],
                          out_dtype=torch.float32)
    assert G.rel_err(dz.cpu(), ref["dz"]) < TOL_BF16


def test_tensor_core_path_through_module(cuda_device):
    """bf16 z with L2-normalised rows (promised) through the drop-in class takes the tensor path; grad_out scaling."""
    x, y = O.make_inputs(1024, 256, "iso")
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    loss, dz = G.kernel_loss_and_grad(zb, y, tau=0.07, similarity="cosine", topk=15, alpha=0.0,
                                      dtype=torch.bfloat16, grad_scale=3.0, unit_rows=True)
    ref = G.oracle_for(zb.float(), y, tau=0.07, similarity="cosine", topk=15, alpha=0.0)
    assert loss == pytest.approx(ref["loss"], rel=TOL_BF16)
    assert G.rel_err(dz, ref["dz"]) < 2 * TOL_BF16      # dz additionally rounded to bf16 by autograd


def test_tensor_core_row_blocks_compose(cuda_device):
    """Row-sharded use of the tensor path: two ranks' row blocks == whole batch."""
    from wav2vec_contr_loss_b200 import functional as Fn
    x, y = O.make_inputs(1024, 256, "iso", classes=3)
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=0.07, similarity="cosine", lam=0.05, t=2.0, topk=15, alpha=0.0)
    a = G.kernel_stats(zb, y, dtype=torch.bfloat16, flags=2, row_offset=0, n_rows=384, **kw)
    b = G.kernel_stats(zb, y, dtype=torch.bfloat16, flags=2, row_offset=384, n_rows=640, **kw)
    partials = a["partials"] + b["partials"]
    stats = torch.cat([a["stats"], b["stats"]])
    whole = Fn.make_problem(1024, 256, 1, tau=0.07, similarity=0, lambda_uni=0.05, uni_t=2.0, topk=15, alpha=0.0)
    ref = G.oracle_for(zb.float(), y, **kw)
    assert float(Fn.finalize(whole, partials)) == pytest.approx(ref["loss"], rel=TOL_BF16)
    dz_b = Fn.backward_rows(b["z"], b["y"], stats, partials, None, b["prob"], out_dtype=torch.float32)
    assert G.rel_err(dz_b.cpu(), ref["dz"][384:]) < TOL_BF16


def test_large_batch_properties_bf16(cuda_device):
    """N = 16384 on the tensor path: loss against the row-blocked fp64 oracle, a sample of
    dz rows against brute force, and permutation equivariance of the gradient."""
    from wav2vec_contr_loss_b200 import functional as Fn
    n = 16384
    x, y = O.make_inputs(n, 256, "iso")
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=0.07, similarity="cosine", lam=0.0, t=2.0, topk=15, alpha=0.0)
    out = G.kernel_stats(zb, y, dtype=torch.bfloat16, flags=2, **kw)
    dz = Fn.backward_rows(out["z"], out["y"], out["stats"], out["partials"], None, out["prob"],
                          out_dtype=torch.float32).cpu()
    ref = O.closed_form(zb.float(), y, temperature=0.07, similarity="cosine", topk_neg=15, alpha=0.0,
                        dtype=torch.float32, block=2048)
    assert float(out["loss"]) == pytest.approx(ref["loss"], rel=1e-4)
    assert G.rel_err(dz, ref["dz"]) < TOL_BF16
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(3))
    out_p = G.kernel_stats(zb[perm], y[perm], dtype=torch.bfloat16, flags=2, **kw)
    dz_p = Fn.backward_rows(out_p["z"], out_p["y"], out_p["stats"], out_p["partials"], None, out_p["prob"],
                            out_dtype=torch.float32).cpu()
    assert float(out_p["loss"]) == pytest.approx(float(out["loss"]), rel=1e-5)
    assert G.rel_err(dz_p, dz[perm]) < TOL_BF16


TC_MINING_CASES = [
    # n, kind, classes, sim, tau, lam, alpha, K
    (1024, "iso", 2, "cosine", 0.07, 0.0, 0.5, 15),        # BASELINE config 3 in bf16
    (1024, "iso", 2, "cosine", 0.07, 0.0, 0.0125, 15),
    (1000, "iso", 3, "cosine", 0.07, 0.0, 1.0, 32),
    (640, "clustered", 2, "geodesic", 0.1, 0.05, 0.37, 5),
    (2048, "ties", 2, "cosine", 0.07, 0.0, 1.0, 15),       # exact duplicates: ties at the K-th boundary
]


@pytest.mark.parametrize("n,kind,classes,sim,tau,lam,alpha,k", TC_MINING_CASES)
def test_tensor_core_path_mining(cuda_device, n, kind, classes, sim, tau, lam, alpha, k):
    """Hard-negative mining fused into the tcgen05 kernels: the selected sets (threshold index per row)
    must equal the stable-sort oracle's, loss/dz within the bf16 tolerance."""
    from wav2vec_contr_loss_b200 import functional as Fn
    x, y = O.make_inputs(n, 256, kind, classes=classes)
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=tau, similarity=sim, lam=lam, t=2.0, topk=k, alpha=alpha)
    ref = G.oracle_for(zb.float(), y, **kw)
    prob_kw = dict(kw)
    z = Fn.canonical_z(zb.to(cuda_device))
    yy = Fn.canonical_labels(y.to(cuda_device), n)
    prob = Fn.make_problem(n, 256, 1, tau=tau, similarity=Fn.similarity_id(sim), lambda_uni=lam, uni_t=2.0, topk=k,
                           alpha=alpha, flags=2)
    stats, partials, loss = Fn.forward_rows(z, yy, prob, want_loss=True)
    assert float(loss) == pytest.approx(ref["loss"], rel=TOL_BF16)
    st = stats.cpu()
    if sim == "cosine":      # exact products of bf16 inputs: the ranking is identical to the fp64 oracle's
        same = (st.view(torch.int32)[:, 5].long() == ref["stats"]["thr_idx"]).float().mean()
        assert float(same) >= 0.999      # near-ties below fp32 resolution may legitimately differ from fp64
    assert float((st[:, 1].double() - ref["stats"]["lse_m"]).abs().max()) < 1e-4
    dz = Fn.backward_rows(z, yy, stats, partials, None, prob, out_dtype=torch.float32)
    assert G.rel_err(dz.cpu(), ref["dz"]) < TOL_BF16


@pytest.mark.parametrize("sim,lam,alpha,k", [("cosine", 0.0, 0.0, 15), ("geodesic", 0.05, 0.5, 7)])
def test_two_phase_forward_equals_single_phase(cuda_device, sim, lam, alpha, k):
    """supcon_forward_rows_local + _remote (own columns first, then the rest) == supcon_forward_rows."""
    from wav2vec_contr_loss_b200 import functional as Fn
    n = 2048
    x, y = O.make_inputs(n, 256, "ties", classes=3)
    z = Fn.canonical_z(F.normalize(x, dim=1).to(torch.bfloat16).to(cuda_device))
    yy = Fn.canonical_labels(y.to(cuda_device), n)
    for row_offset, n_rows in ((512, 768), (0, 256), (1792, 256)):
        prob = Fn.make_problem(n, 256, 1, tau=0.07, similarity=Fn.similarity_id(sim), lambda_uni=lam, topk=k, alpha=alpha,
                               flags=2, row_offset=row_offset, n_rows=n_rows)
        s1, p1, _ = Fn.forward_rows(z, yy, prob, want_loss=False)
        ws = Fn.forward_rows_local(z, yy, prob)
        s2, p2 = Fn.forward_rows_remote(z, yy, prob, ws)
        assert torch.equal(s1.view(torch.int32)[:, [2, 3, 5]], s2.view(torch.int32)[:, [2, 3, 5]])   # counts, threshold idx
        assert torch.allclose(s1[:, [0, 1, 6, 7]], s2[:, [0, 1, 6, 7]], rtol=1e-5, atol=1e-6)       # sums: order differs
        assert torch.allclose(p1, p2, rtol=1e-6)


@pytest.mark.parametrize("n", [100, 200])
def test_bf16_small_and_mid_batches_take_exact_kernels(cuda_device, n):
    """bf16 z below the tensor-path threshold: N=100 -> single-launch cluster kernel, N=200 -> tiled FFMA kernels."""
    x, y = O.make_inputs(n, 256, "iso", classes=3)
    zb = F.normalize(x, dim=1).to(torch.bfloat16)
    kw = dict(tau=0.07, similarity="geodesic", lam=0.05, topk=7, alpha=0.5)
    loss, dz = G.kernel_loss_and_grad(zb, y, dtype=torch.bfloat16, **kw)
    ref = G.oracle_for(zb.float(), y, **kw)
    assert loss == pytest.approx(ref["loss"], rel=1e-5)          # fp32 math on the bf16-rounded inputs
    assert G.rel_err(dz, ref["dz"]) < 2 * TOL_BF16               # dz rounded to bf16 by autograd


def test_sharded_loss_single_rank_nccl(cuda_device):
    """ShardedSupConLoss on a world_size-1 NCCL group: the GPU code path of the distributed module
    (in-place all-gather on the side stream, two-phase entry points, stats exchange) == SupConBinaryLoss."""
    import socket
    import torch.distributed as dist
    from wav2vec_contr_loss_b200 import SupConBinaryLoss
    from wav2vec_contr_loss_b200.distributed import ShardedSupConLoss
    if dist.is_initialized():
        pytest.skip("a process group already exists")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=cuda_device)
    try:
        x, y = O.make_inputs(512, 256, "iso")
        for dtype in (torch.float32, torch.bfloat16):
            z1 = F.normalize(x, dim=1).to(cuda_device).to(dtype).requires_grad_(True)
            z2 = z1.detach().clone().requires_grad_(True)
            yy = y.to(cuda_device)
            ma, mb = ShardedSupConLoss(0.07, "cosine", 0.05), SupConBinaryLoss(0.07, "cosine", 0.05)
            ma.assume_unit_rows = mb.assume_unit_rows = True      # bf16: tensor path
            a = ma(z1, yy, topk_neg=15, alpha=0.0)
            b = mb(z2, yy, topk_neg=15, alpha=0.0)
            (2.0 * a).backward(); (2.0 * b).backward()
            assert float(a) == pytest.approx(float(b), rel=1e-6)
            assert G.rel_err(z1.grad.float().cpu(), z2.grad.float().cpu()) < 1e-5
    finally:
        dist.destroy_process_group()
