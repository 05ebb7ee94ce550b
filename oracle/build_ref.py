#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference next to the oracle so it can travel to the GPU box.

    python oracle/build_ref.py [--check]

The reference is a flat directory of Python scripts (no package, no build system, nothing to compile).
This recipe copies, byte for byte, the few files of the hot path and of its direct callers from where they
lie under ``/root/reference`` into ``oracle/_ref/`` and writes ``oracle/_ref/MANIFEST.json`` (sha256 of every
file).  ``oracle/_ref/`` is listed in ``.gitignore`` (it never enters the history: the repository holds no
reference source) but NOT in ``.gpurunignore``, so the staged files ship with the gpurun snapshot exactly like
the repo's own built ``.so``.  On the GPU box ``/root/reference`` does not exist and this script is a no-op
that only verifies the manifest of what was shipped.

Who may use ``oracle/_ref``: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` / ``--workload stage1`` baseline legs -- as the checker or as the timed CPU reference,
never on the product path (``wav2vec_contr_loss_b200`` does not import ``oracle``).

Files (all read-only inputs; nothing is edited):
  loss.py                the objective itself                      (hot path)
  stage1_utils.py        train_one_epoch / evaluate / sampler / alpha schedule  (callers, a1, a9)
  stage1_config.py       the argparse config the callers read
  compression_module.py  the head that produces the loss's input   (8f N1)
  encoder.py             Wav2Vec2Encoder wrapper                   (configs[4])
  RawBoost.py            imported by stage1_utils at module level
  train_stage1.py        the entry script (imported only with data_loader stubbed)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SUPCON_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("loss.py", "stage1_utils.py", "stage1_config.py", "compression_module.py", "encoder.py", "RawBoost.py",
         "train_stage1.py")


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def staged() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "MANIFEST.json")) and os.path.isfile(os.path.join(REF_DST, "loss.py"))


def verify() -> dict:
    """Check the staged copies against their manifest (and against the source tree when it is present)."""
    with open(os.path.join(REF_DST, "MANIFEST.json")) as fh:
        man = json.load(fh)
    for name, digest in man["sha256"].items():
        got = _sha(os.path.join(REF_DST, name))
        if got != digest:
            raise RuntimeError(f"oracle/_ref/{name} does not match its manifest (staged copy was edited?)")
        src = os.path.join(REF_SRC, name)
        if os.path.isfile(src) and _sha(src) != digest:
            raise RuntimeError(f"oracle/_ref/{name} differs from {src}: re-run oracle/build_ref.py")
    return man


def build(force: bool = False) -> str:
    """Stage the reference files when the source tree is present; otherwise verify what was shipped."""
    have_src = os.path.isfile(os.path.join(REF_SRC, "loss.py"))
    if not have_src:
        if staged():
            verify()
            return REF_DST
        raise FileNotFoundError(f"neither {REF_SRC} nor a staged oracle/_ref exists")
    if staged() and not force:
        try:
            verify()
            return REF_DST
        except RuntimeError:
            pass
    os.makedirs(REF_DST, exist_ok=True)
    digests = {}
    for name in FILES:
        shutil.copyfile(os.path.join(REF_SRC, name), os.path.join(REF_DST, name))
        digests[name] = _sha(os.path.join(REF_DST, name))
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "note": "verbatim copies; git-ignored; test/baseline use only",
                   "sha256": digests}, fh, indent=1)
    return REF_DST


if __name__ == "__main__":
    if "--check" in sys.argv:
        print(json.dumps(verify(), indent=1))
    else:
        print(build(force="--force" in sys.argv))
