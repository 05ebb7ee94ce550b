"""The reference's OWN callers, unedited, around the loss (north_star: "train_stage1.py, stage1_utils.py and
stage1_config.py run unchanged").

``oracle/ref_loader.run_unchanged_train_stage1`` executes the reference's train_stage1.py as __main__ from the
reference root (the source tree in the build container, the staged verbatim copies ``oracle/_ref`` on the GPU
box) with a synthetic ``data_loader`` module injected (the real one needs soundfile / librosa / ASVspoof on
disk) and a random-init Wav2Vec2 saved locally for ``--model_name``.  Which module named ``loss`` it imports
is decided by sys.path alone: the reference's own loss.py, or this repo's drop-in
(``wav2vec_contr_loss_b200/dropin/loss.py`` -> the CUDA kernels).  Both runs start from the same seed, so the
printed epoch losses and the saved head checkpoint must agree.
"""
import glob
import os

import pytest
import torch

from oracle import ref_loader as R

DROPIN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wav2vec_contr_loss_b200",
                          "dropin")

needs_ref = pytest.mark.skipif(not R.reference_available(),
                               reason="reference neither at /root/reference nor staged under oracle/_ref")


def _argv(tmp, model_dir, sub, **over):
    a = {"--model_name": model_dir, "--save_dir": os.path.join(tmp, sub), "--epochs": "2", "--batch_size": "16",
         "--num_workers": "0", "--use_rawboost": "0", "--temperature": "0.07", "--supcon_similarity": "geodesic",
         "--uniformity_weight": "0.05", "--warmup_epochs": "1", "--alpha_ramp_epochs": "1", "--topk_neg": "3",
         "--head_lr": "1e-3", "--seed": "1337"}
    a.update(over)
    return [x for kv in a.items() for x in kv]


def _best_ckpt(tmp, sub):
    files = glob.glob(os.path.join(tmp, sub, "*", "*_stage1_head_best.pt"))
    assert len(files) == 1, files
    return torch.load(files[0], map_location="cpu", weights_only=False)


@needs_ref
def test_unchanged_train_stage1_runs_on_cpu_with_the_reference_loss(tmp_path):
    """Harness sanity on the CPU: the reference's script, config parser, sampler, loops and loss, unedited."""
    tmp = str(tmp_path)
    model_dir = R.save_tiny_wav2vec2(os.path.join(tmp, "w2v"), layers=1)
    dl = R.synthetic_data_loader_module(n_items=32, samples=2000)
    recs, out, loss_file = R.run_unchanged_train_stage1(R.reference_root(), _argv(tmp, model_dir, "ref", **{"--epochs": "1"}),
                                                        dl)
    assert os.path.samefile(os.path.dirname(loss_file), R.reference_root())
    assert len(recs) == 1 and recs[0]["epoch"] == 1 and recs[0]["alpha"] == 0.0
    assert 0.0 < recs[0]["train_loss"] < 10.0 and 0.0 < recs[0]["dev_loss"] < 10.0
    assert "SUPCON_SIMILARITY=geodesic" in out          # stage1_config.print_config ran
    assert "compression_state_dict" in _best_ckpt(tmp, "ref")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("similarity,lam", [("geodesic", "0.05"), ("cosine", "0.0")])
def test_unchanged_train_stage1_with_the_dropin_matches_the_reference_loss_on_gpu(cuda_device, tmp_path, similarity, lam):
    """Same unedited script twice on the GPU: `import loss` -> reference loss.py, then -> the drop-in.  Epoch 2
    runs with alpha = 1 (hard-negative mining, top-3), epoch 1 with alpha = 0; uniformity on for geodesic."""
    tmp = str(tmp_path)
    model_dir = R.save_tiny_wav2vec2(os.path.join(tmp, "w2v"))
    over = {"--supcon_similarity": similarity, "--uniformity_weight": lam}
    ref_recs, _, ref_file = R.run_unchanged_train_stage1(R.reference_root(), _argv(tmp, model_dir, "ref", **over),
                                                         R.synthetic_data_loader_module())
    our_recs, _, our_file = R.run_unchanged_train_stage1(DROPIN_DIR, _argv(tmp, model_dir, "ours", **over),
                                                         R.synthetic_data_loader_module())
    assert os.path.samefile(os.path.dirname(our_file), DROPIN_DIR)        # the drop-in really was the `loss` module
    assert os.path.samefile(os.path.dirname(ref_file), R.reference_root())
    assert len(ref_recs) == len(our_recs) == 2 and our_recs[1]["alpha"] == 1.0
    for a, b in zip(ref_recs, our_recs):   # the script prints 4 decimals
        assert a["alpha"] == b["alpha"]
        assert abs(a["train_loss"] - b["train_loss"]) <= 2e-4, (a, b)
        assert abs(a["dev_loss"] - b["dev_loss"]) <= 2e-4, (a, b)
    ck_ref, ck_our = _best_ckpt(tmp, "ref"), _best_ckpt(tmp, "ours")
    assert ck_ref["epoch"] == ck_our["epoch"]
    for k, v in ck_ref["compression_state_dict"].items():
        w = ck_our["compression_state_dict"][k]
        assert float((v - w).norm() / v.norm()) < 2e-4, k
