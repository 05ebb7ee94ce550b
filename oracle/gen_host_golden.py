"""TEST INFRASTRUCTURE ONLY.  Regenerate tests/golden/stage1_host.json from the REAL reference.

    python -m oracle.gen_host_golden          (build container only: needs /root/reference)

Pins the host-side pieces either side of the loss by running the reference's own ``stage1_utils.py``:

* ``alpha_for_epoch`` (stage1_utils.py:84-88) over a grid of schedules and epochs;
* ``BalancedBatchSampler`` (stage1_utils.py:22-53): the exact batches for a synthetic label list, several
  epochs iterated in sequence on one sampler object (its pools are shuffled in place), world sizes 1 and 2;
* ``train_one_epoch`` / ``evaluate`` (stage1_utils.py:101-153) with the reference's ``loss.py`` on a tiny
  encoder + head (``tiny_stage1``, CPU fp32): per-epoch mean losses, alphas and the head's final weights.
"""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

from oracle.ref_loader import REFERENCE_ROOT, load_reference_module

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stage1_host.json")

SCHEDULES = [  # warmup_epochs, alpha_ramp_epochs, alpha_end
    (8, 80, 1),          # train_stage1_from_emb.py:44-46
    (100, 80, 1),        # stage1_config.py defaults
    (0, 0, 0.5),
    (3, 7, 0.3),
    (5, 1, 2.0),
]
EPOCHS = list(range(0, 12)) + [20, 48, 87, 88, 89, 100, 101, 140, 180, 181, 500]

SAMPLERS = [  # n_real, n_fake, n_other_label, batch_size, seed, epochs iterated in order
    (37, 53, 0, 8, 1337, [1, 2, 3]),
    (64, 64, 5, 16, 0, [0, 0, 7]),
    (10, 200, 0, 4, 42, [5]),
    (3, 50, 0, 8, 1, [1]),             # fewer bonafide than half a batch: no batches
]


def label_list(n_real, n_fake, n_other):
    """deterministic interleaving of labels 1 / 0 / 2 (2 = neither pool)"""
    labels, r, f, o = [], n_real, n_fake, n_other
    i = 0
    while r or f or o:
        pick = i % 3
        if pick == 0 and r:
            labels.append(1); r -= 1
        elif pick == 1 and f:
            labels.append(0); f -= 1
        elif pick == 2 and o:
            labels.append(2); o -= 1
        elif f:
            labels.append(0); f -= 1
        elif r:
            labels.append(1); r -= 1
        else:
            labels.append(2); o -= 1
        i += 1
    return labels


class TinyEncoder(nn.Module):
    """(B, T_audio) waveforms -> (B, K, F, T) hidden states, the layout the reference's encoder hands the head."""

    def __init__(self, layers=3, feat=12, hop=16):
        super().__init__()
        self.layers, self.hop = layers, hop
        self.proj = nn.Linear(hop, layers * feat)

    def forward(self, waveforms, attention_mask=None):
        b = waveforms.size(0)
        frames = (waveforms * attention_mask.to(waveforms.dtype)).view(b, -1, self.hop)      # (B, T, hop)
        hs = torch.tanh(self.proj(frames)).view(b, frames.size(1), self.layers, -1)           # (B, T, K, F)
        return hs.permute(0, 2, 3, 1).contiguous()


class TinyHead(nn.Module):
    """layer mean -> LeakyReLU -> Linear per frame: (B, K, F, T) -> (B, D, T)."""

    def __init__(self, feat=12, dim=8):
        super().__init__()
        self.act, self.fc = nn.LeakyReLU(), nn.Linear(feat, dim)

    def forward(self, hs):
        x = self.act(hs.mean(dim=1))
        return self.fc(x.transpose(1, 2)).transpose(1, 2)


def tiny_stage1(finetune=False):
    """Deterministic tiny Stage-1 problem: (encoder, head, train batches, dev batches, cfg)."""
    g = torch.Generator().manual_seed(1337)
    torch.manual_seed(1337)
    enc, head = TinyEncoder(), TinyHead()

    def batches(count, bsz):
        out = []
        for _ in range(count):
            y = torch.randperm(bsz, generator=g) % 2
            wave = torch.randn(bsz, 64, generator=g) + 0.75 * (2.0 * y.view(-1, 1) - 1.0)
            wave[:, 48:] = torch.where(torch.rand(bsz, 1, generator=g) < 0.5, 0.0, 1.0) * wave[:, 48:]   # padding
            out.append((wave, y, ["spk"] * bsz))
        return out

    cfg = SimpleNamespace(finetune_encoder=finetune, use_rawboost=False, topk_neg=3, warmup_epochs=1,
                          alpha_ramp_epochs=2, alpha_end=1.0)
    return enc, head, batches(4, 12), batches(2, 10), cfg


STEP_CASES = [  # name, similarity, tau, lambda_uni, finetune
    ("cosine_frozen", "cosine", 0.2, 0.0, False),
    ("geodesic_uni_finetune", "geodesic", 0.1, 0.1, True),
]


def run_reference_epochs(ref, ref_loss, sim, tau, lam, finetune, epochs=3):
    enc, head, train, dev, cfg = tiny_stage1(finetune)
    loss_fn = ref_loss.SupConBinaryLoss(tau, sim, lam, 2.0)
    params = list(head.parameters()) + (list(enc.parameters()) if finetune else [])
    opt = torch.optim.AdamW(params, lr=5e-3, weight_decay=3e-3)
    rec = {"train": [], "alpha": [], "dev": []}
    for epoch in range(1, epochs + 1):
        tl, alpha = ref.train_one_epoch(enc, head, loss_fn, train, opt, torch.device("cpu"), epoch, cfg)
        rec["train"].append(float(tl)); rec["alpha"].append(float(alpha))
        rec["dev"].append(float(ref.evaluate(enc, head, loss_fn, dev, torch.device("cpu"), cfg)))
    rec["head"] = torch.cat([p.detach().reshape(-1) for p in head.parameters()]).double().tolist()
    rec["encoder"] = torch.cat([p.detach().reshape(-1) for p in enc.parameters()]).double().tolist()
    return rec


def main():
    sys.path.append(REFERENCE_ROOT)      # stage1_utils imports its sibling RawBoost
    ref = load_reference_module("stage1_utils")
    ref_loss = load_reference_module("loss")
    out = {"alpha": [], "sampler": [], "epochs": {}}
    for name, sim, tau, lam, finetune in STEP_CASES:
        out["epochs"][name] = run_reference_epochs(ref, ref_loss, sim, tau, lam, finetune)
    for warm, ramp, end in SCHEDULES:
        cfg = SimpleNamespace(warmup_epochs=warm, alpha_ramp_epochs=ramp, alpha_end=end)
        out["alpha"].append({"warmup_epochs": warm, "alpha_ramp_epochs": ramp, "alpha_end": end, "epochs": EPOCHS,
                             "values": [float(ref.alpha_for_epoch(e, cfg)) for e in EPOCHS]})
    for n_real, n_fake, n_other, bs, seed, epochs in SAMPLERS:
        labels = label_list(n_real, n_fake, n_other)
        ds = SimpleNamespace(data=[(f"utt{i}", lab) for i, lab in enumerate(labels)])
        case = {"labels": labels, "batch_size": bs, "seed": seed, "epochs": epochs, "worlds": {}}
        for world in (1, 2):
            per_rank = []
            for rank in range(world):
                s = ref.BalancedBatchSampler(ds, bs, seed=seed, rank=rank, world_size=world)
                runs = []
                for e in epochs:
                    s.set_epoch(e)
                    runs.append({"len": len(s), "batches": [list(map(int, b)) for b in s]})
                per_rank.append(runs)
            case["worlds"][str(world)] = per_rank
        out["sampler"].append(case)
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
