"""AlignedTargetsLoss -- NLL on the Viterbi alignment, the consumer of ``get_alignment_3d`` in the reference
(pytorch_end2end/modules/alignment_loss.py:7-33).  Same constructor and forward signature; the alignment runs on the
GPU (end2end_b200.utils.alignment) and stays there when ``log_probs`` is a CUDA tensor."""
import torch.nn as nn
import torch.nn.functional as F

from ..utils.alignment import get_alignment_3d, get_alignment_3d_device


class AlignedTargetsLoss(nn.Module):
    def __init__(self, is_ctc, ignore_blank=False):
        super().__init__()
        self._is_ctc = is_ctc
        self._ignore_blank = ignore_blank

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        """
        :param log_probs: batch_size * sequence_length * num_labels
        :param targets: batch_size * sequence_length, fill with -1 if ignored label
        :param input_lengths: batch_size
        :param target_lengths: batch_size
        :return: per-utterance loss [batch_size]
        """
        if log_probs.is_cuda:     # the reference computes on the CPU and moves the result to log_probs' device
            targets_new = get_alignment_3d_device(log_probs, targets, input_lengths, target_lengths, is_ctc=self._is_ctc)
        else:
            targets_new = get_alignment_3d(log_probs, targets, input_lengths, target_lengths, is_ctc=self._is_ctc)
        batch_size, sequence_length, _ = log_probs.shape
        if self._ignore_blank:
            targets_new[targets_new == 0] = -100
        loss = F.nll_loss(log_probs.reshape(batch_size * sequence_length, -1),
                          targets_new.reshape(batch_size * sequence_length),
                          reduction="none", ignore_index=-100).reshape(batch_size, sequence_length)
        loss = loss.sum(dim=-1) / input_lengths.to(loss.device)
        return loss
