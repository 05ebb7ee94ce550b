"""Drop-in replacements for the loss classes of the reference ``loss.py``.

Same class names, constructor/forward signatures, attributes and error
behaviour as the reference (loss.py:6-32, :110-114, :156-171, :213-258), so
``train_stage1.py`` / ``stage1_utils.py`` / ``stage1_config.py`` run unchanged
with ``wav2vec_contr_loss_b200/dropin`` ahead of the reference on ``sys.path``.
The SupCon classes run on libsupcon_b200.so; they hold no parameters or buffers.
"""
import torch
import torch.nn as nn

from .functional import similarity_id, supcon_loss


class SupConBinaryLoss(nn.Module):
    """main = (1 - alpha) * SupCon_full + alpha * SupCon_mined (+ lambda * uniformity).

    reference loss.py:6-153.  ``z`` is taken as given (callers L2-normalise it,
    stage1_utils.py:123); ``labels`` may be any dtype comparable with ``==``.
    """

    def __init__(self, temperature: float = 0.2, similarity: str = "geodesic",
                 uniformity_weight: float = 0.0, uniformity_t: float = 2.0):
        super().__init__()
        self.tau = temperature
        self.similarity = similarity.lower()
        self.lambda_uni = float(uniformity_weight)
        self.uni_t = float(uniformity_t)
        if self.similarity not in ("cosine", "geodesic"):
            raise ValueError(f"Unknown similarity: {similarity}")
        self.kernel_flags = 0
        # None: rows count as L2-normalised only when z came from this library's l2_normalize()/embed();
        # True: the caller promises it (bf16 cosine inputs then take the tensor-core path; checked on the device)
        self.assume_unit_rows = None

    def forward(self, z: torch.Tensor, labels: torch.Tensor, topk_neg: int = 32,
                alpha: float = 0.0) -> torch.Tensor:
        return supcon_loss(z, labels, temperature=self.tau, similarity=self.similarity,
                           uniformity_weight=self.lambda_uni, uniformity_t=self.uni_t,
                           topk_neg=topk_neg, alpha=alpha, flags=self.kernel_flags,
                           unit_rows=self.assume_unit_rows)


class SupConMultiClassLoss(nn.Module):
    """Khosla-style SupCon over arbitrary class ids (reference loss.py:156-210):
    the full-SupCon branch with cosine similarity, no mining, no uniformity."""

    takes_mining_args = False      # forward(z, labels) only: stage1.call_loss drops topk_neg / alpha

    def __init__(self, temperature: float = 0.1):
        super().__init__()
        self.tau = temperature
        self.kernel_flags = 0
        self.assume_unit_rows = None

    def forward(self, z: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        assert labels.dim() == 1 and labels.size(0) == z.size(0), "labels must be shape (B,)"
        return supcon_loss(z, labels, temperature=self.tau, similarity="cosine", uniformity_weight=0.0,
                           topk_neg=0, alpha=0.0, flags=self.kernel_flags, unit_rows=self.assume_unit_rows)


class BCEBinaryLoss(nn.Module):
    """BCE-with-logits baseline (reference loss.py:213-239); not on the SupCon
    path, kept because ``baseline_train.py:14`` imports it from this module."""

    def __init__(self, pos_weight=None):
        super().__init__()
        self.pos_weight = pos_weight
        if pos_weight is not None:
            self.register_buffer("_pos_weight_tensor", torch.tensor([float(pos_weight)], dtype=torch.float32))

    def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        weight = None
        if self.pos_weight is not None:
            self._pos_weight_tensor = self._pos_weight_tensor.to(logits.device)
            weight = self._pos_weight_tensor
        return nn.functional.binary_cross_entropy_with_logits(logits, labels.float(), pos_weight=weight)


def compute_pos_weight_from_dataset(dataset) -> float:
    """neg/pos ratio over ``dataset.data`` (label = item[1], 1 = bonafide); 1.0
    when a class is absent (reference loss.py:242-258)."""
    labels = [int(item[1]) for item in dataset.data]
    pos = sum(1 for y in labels if y == 1)
    neg = len(labels) - pos
    return float(neg) / float(pos) if pos and neg else 1.0


__all__ = ["SupConBinaryLoss", "SupConMultiClassLoss", "BCEBinaryLoss", "compute_pos_weight_from_dataset",
           "similarity_id"]
