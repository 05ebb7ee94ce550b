"""TEST-ONLY stand-in for the CUDA kernels, backed by the CPU oracle, so the
multi-rank host logic (wav2vec_contr_loss_b200.distributed) can run under gloo
without a GPU.  Never imported by the product package."""
import torch

from oracle import supcon_oracle as O

_KEYS = ("lse", "lse_m", "npos", "nneg", "thr_val", "thr_idx", "wsum", "pos_mean")


class OracleKernels:
    name = "oracle"

    @staticmethod
    def _kw(prob):
        return dict(tau=prob.tau, similarity=prob.similarity, topk=prob.topk, lambda_uni=prob.lambda_uni,
                    uni_t=prob.uni_t)

    @staticmethod
    def forward_rows(z_all, labels_all, prob):
        stats, partials = O.rowblock_forward(z_all.double(), labels_all.long(), prob.row_offset, prob.n_rows,
                                             **OracleKernels._kw(prob))
        packed = torch.stack([stats[k].double() for k in _KEYS], dim=1)
        return packed, partials

    @staticmethod
    def finalize(prob, partials):
        loss, _ = O.loss_from_partials(partials, prob.n_total, alpha=prob.alpha, lambda_uni=prob.lambda_uni)
        return torch.tensor(loss, dtype=torch.float64)

    @staticmethod
    def backward_rows(z_all, labels_all, stats_all, partials, grad_out, prob, out_dtype):
        stats = {k: stats_all[:, i] for i, k in enumerate(_KEYS)}
        for k in ("npos", "nneg", "thr_idx"):
            stats[k] = stats[k].long()
        _, coef = O.loss_from_partials(partials, prob.n_total, alpha=prob.alpha, lambda_uni=prob.lambda_uni)
        dz = O.rowblock_backward(z_all.double(), labels_all.long(), prob.row_offset, prob.n_rows, stats, coef,
                                 **OracleKernels._kw(prob))
        return (dz * float(grad_out)).to(out_dtype)
