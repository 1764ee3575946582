"""CTCDecoder -- greedy (argmax) CTC decoding on the GPU behind the reference's decoder class
(pytorch_end2end/decoders/ctc_decoder.py:25-149).

Only the greedy path is on this framework's hot path: ``decode`` with ``beam_width == 1`` and
``decode_greedy`` (decoders/ctc_decoder.py:92-93, 117-149; src/decoders/ctc_decoder.cpp:443-490).
Prefix beam search and KenLM rescoring stay the reference's CPU code and are out of scope;
asking for them here raises ``NotImplementedError`` instead of silently doing something else.
"""
from collections import namedtuple

import torch

from ..engine import CTCGreedyEngine


class CTCDecoderError(Exception):
    pass


DecoderResults = namedtuple("DecoderResults", ["decoded_targets",
                                               "decoded_targets_lengths",
                                               "decoded_sentences"])


class CTCDecoder:
    """
    :param beam_width: only ``1`` (greedy) is executed by this engine
    :param after_logsoftmax: ignored by greedy decoding (argmax is invariant to log-softmax)
    :param blank_idx: id of the blank label, default ``0``
    :param time_major: logits are ``(T, B, V)`` instead of ``(B, T, V)``
    :param labels: list of strings with labels (including the blank symbol), e.g. ``["_", "a", "b"]``
    :param lm_path, lmwt, wip, oov_penalty, case_sensitive: language-model options of the reference's
        beam search; accepted for signature compatibility, unused by greedy decoding
    """

    def __init__(self, beam_width=100, after_logsoftmax=False, blank_idx=0, time_major=False, labels=None,
                 lm_path=None, lmwt=1.0, wip=1.0, oov_penalty=-10, case_sensitive=True):
        self._beam_width = beam_width
        self._blank_idx = blank_idx
        self._after_logsoftmax = after_logsoftmax
        self._labels = list(labels or [])
        self._lm_path = lm_path or ""
        self._lmwt = lmwt
        self._wip = wip
        self._oov_penalty = oov_penalty
        self._time_major = time_major
        self._case_sensitive = case_sensitive
        self._decoder = CTCGreedyEngine(self._blank_idx)

    def decode(self, logits, logits_lengths=None):
        """Greedy decoding when ``beam_width == 1`` (decoders/ctc_decoder.py:92-93)."""
        if self._beam_width == 1:
            return self.decode_greedy(logits, logits_lengths)
        raise NotImplementedError(
            "prefix beam search / LM decoding is outside this engine's scope (it stays the reference's "
            "CPU code); construct CTCDecoder(beam_width=1) or call decode_greedy()")

    def decode_greedy(self, logits, logits_lengths=None):
        """
        :param logits: ``(T, B, V)`` if ``time_major`` else ``(B, T, V)``; CPU or CUDA tensor
        :param logits_lengths: ``(B,)`` or ``None`` (decode all frames)
        :return: ``DecoderResults(decoded_targets [B,T] int64 zero padded, decoded_targets_lengths [B]
            int64, decoded_sentences list[str])`` -- CPU tensors, as in the reference
        """
        if self._time_major:
            logits = logits.transpose(1, 0)
        decoded_targets, decoded_targets_lengths = self._decoder.decode_greedy(
            logits_=logits, logits_lengths_=logits_lengths)
        decoded_sentences = self._sentences(decoded_targets, decoded_targets_lengths)
        return DecoderResults(decoded_targets, decoded_targets_lengths, decoded_sentences)

    def _sentences(self, targets, lengths):
        # src/decoders/ctc_decoder.cpp:212-220: concatenate labels; "" for every utterance without labels
        if not self._labels:
            return [""] * targets.size(0)
        rows, lens = targets.tolist(), lengths.tolist()
        return ["".join(self._labels[i] for i in row[:n]) for row, n in zip(rows, lens)]
