"""Module named ``loss`` for the reference's unchanged callers.

Put this directory ahead of the reference checkout on ``sys.path`` /
``PYTHONPATH`` and ``from loss import SupConBinaryLoss`` (train_stage1.py:14,
train_stage1_from_emb.py, train_multiclass_con.py, baseline_train.py:14) binds
to the B200 implementation.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from wav2vec_contr_loss_b200.loss import (BCEBinaryLoss, SupConBinaryLoss,  # noqa: E402,F401
                                          SupConMultiClassLoss, compute_pos_weight_from_dataset)
