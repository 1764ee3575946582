"""CTCWithoutBlankLoss -- same constructor and call signature as the reference's module
(pytorch_end2end/modules/ctc_without_blank.py:7-35): ``CTCWithoutBlankLoss(reduce=True, after_softmax=False,
space_idx=-1)(logits, targets, logits_lengths, targets_lengths)``; the lattice work runs in csrc/ctc_noblank.cu."""
import torch.nn as nn
import torch.nn.functional as F

from ..functions.ctc_without_blank import CTCWithoutBlankLossFunction


class CTCWithoutBlankLoss(nn.Module):
    """``after_softmax``: the input holds probabilities (their log is taken) instead of raw logits (log_softmax);
    ``reduce``: return the sum over the batch instead of the per-utterance losses; ``space_idx``: -1 for the plain
    target lattice, else the symbol that may be inserted before and after the targets."""

    def __init__(self, reduce=True, after_softmax=False, space_idx=-1):
        super().__init__()
        self._reduce, self._after_softmax, self._space_idx = bool(reduce), bool(after_softmax), int(space_idx)

    def forward(self, logits, targets, logits_lengths, targets_lengths):
        log_probs = logits.log() if self._after_softmax else F.log_softmax(logits, dim=2)
        per_utterance = CTCWithoutBlankLossFunction.apply(log_probs, targets, logits_lengths, targets_lengths, self._space_idx)
        return per_utterance.sum() if self._reduce else per_utterance
