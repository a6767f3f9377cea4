"""Drop-in for the reference's ``scripts/train/loss.py``: same classes, constructor arguments and ``get_loss``
contract; the score matrix, the loss and its gradient come from one fused call into the sm_100a kernels
(``ops.score_loss``: 2 launches in-batch, 1 for own-docs scoring).

Layout contract (reference collator): docs are query-major with the positive first, so with G = Nd // Nq the positive
of query i is row i*G.
"""
import logging

from ... import ops

logger = logging.getLogger(__name__)


class SparseTrainingLoss:
    def __init__(self, weight=1):
        self.weight = weight

    def __call__(self, q_rep, d_rep, inputs):
        raise NotImplementedError

    def get_loss(self, q_rep, d_rep, inputs):
        return self.weight * self(q_rep, d_rep, inputs)


def _teacher_scores(q_rep, inputs, in_batch, what):
    if not in_batch and q_rep.shape[0] == 1:
        # the reference squeezes the batch dimension away here and fails inside (log_)softmax(dim=1) (loss.py:33-35)
        raise IndexError(f"{what} without in-batch negatives needs a batch of at least 2 queries "
                         "(Dimension out of range in the reference)")
    return inputs["scores"]


class KLDivLoss(SparseTrainingLoss):
    """KL(softmax(teacher/T) || softmax(student/T)), summed over docs, mean over queries (reference :18-43)."""

    def __init__(self, use_in_batch_negatives=False, weight=1, temperature=1.0):
        self.use_in_batch_negatives = use_in_batch_negatives
        self.temperature = temperature
        super().__init__(weight)

    def __call__(self, q_rep, d_rep, inputs):
        teacher = _teacher_scores(q_rep, inputs, self.use_in_batch_negatives, "KLDivLoss")
        G = d_rep.shape[0] // q_rep.shape[0]
        return ops.score_loss(q_rep, d_rep, teacher, "kldiv", G, self.use_in_batch_negatives, self.temperature)


class MarginMSELoss(SparseTrainingLoss):
    """MSE between student and teacher margins against column 0 (reference :46-77)."""

    def __init__(self, use_in_batch_negatives=False, weight=1, temperature=1.0):
        self.use_in_batch_negatives = use_in_batch_negatives
        self.temperature = temperature
        super().__init__(weight)

    def __call__(self, q_rep, d_rep, inputs):
        teacher = _teacher_scores(q_rep, inputs, self.use_in_batch_negatives, "MarginMSELoss")
        G = d_rep.shape[0] // q_rep.shape[0]
        return ops.score_loss(q_rep, d_rep, teacher, "marginmse", G, self.use_in_batch_negatives, self.temperature)


class InfoNCELoss(SparseTrainingLoss):
    """Softmax cross-entropy with the positive first; in-batch negatives are every query's hard negatives, other
    queries' positives are left out (reference :80-107). Extra keyword arguments are accepted and ignored."""

    def __init__(self, weight=1, use_in_batch_negatives=False, **kwargs):
        self.use_in_batch_negatives = use_in_batch_negatives
        super().__init__(weight)

    def __call__(self, q_rep, d_rep, inputs):
        G = d_rep.shape[0] // q_rep.shape[0]
        return ops.score_loss(q_rep, d_rep, None, "infonce", G, self.use_in_batch_negatives, 1.0)


LOSS_CLS_MAP = {"infonce": InfoNCELoss, "kldiv": KLDivLoss, "marginmse": MarginMSELoss}
