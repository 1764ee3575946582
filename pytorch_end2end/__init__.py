"""Drop-in name for the reference package: ``from pytorch_end2end import CTCLoss, CTCDecoder,
CTCEncoder`` (reference pytorch_end2end/__init__.py:1-6) resolves to the B200-native engine."""
import torch  # noqa: F401

from end2end_b200 import CTCDecoder, CTCEncoder, CTCLoss

__all__ = ["CTCLoss", "CTCDecoder", "CTCEncoder"]
