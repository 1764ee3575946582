"""AlignedTargetsLoss -- frame-level NLL against the Viterbi alignment of the targets, the consumer of
``get_alignment_3d`` in the reference (pytorch_end2end/modules/alignment_loss.py:7-33).  Same constructor
(``is_ctc``, ``ignore_blank``) and call signature; the alignment is one GPU kernel launch
(end2end_b200.utils.alignment) and never leaves the device when ``log_probs`` is a CUDA tensor."""
import torch.nn as nn
import torch.nn.functional as F

from ..utils.alignment import get_alignment_3d, get_alignment_3d_device

_IGNORE = -100      # the fill value of get_alignment_3d past an utterance's frames


class AlignedTargetsLoss(nn.Module):
    def __init__(self, is_ctc, ignore_blank=False):
        super().__init__()
        self._is_ctc, self._ignore_blank = is_ctc, ignore_blank

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        """log_probs ``[B, T, V]`` (log-probabilities), targets ``[B, Lmax]``, lengths ``[B]`` -> loss ``[B]``: the summed
        negative log-probability of the best path's labels divided by the utterance's frame count."""
        align = get_alignment_3d_device if log_probs.is_cuda else get_alignment_3d
        frame_labels = align(log_probs, targets, input_lengths, target_lengths, is_ctc=self._is_ctc)
        if self._ignore_blank:
            frame_labels = frame_labels.masked_fill(frame_labels == 0, _IGNORE)
        B, T, V = log_probs.shape
        nll = F.nll_loss(log_probs.reshape(B * T, V), frame_labels.reshape(B * T), reduction="none", ignore_index=_IGNORE)
        return nll.view(B, T).sum(dim=1) / input_lengths.to(nll.device)
