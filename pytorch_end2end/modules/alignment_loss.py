"""Drop-in for pytorch_end2end/modules/alignment_loss.py (reference :7-33)."""
from end2end_b200.modules.alignment_loss import AlignedTargetsLoss  # noqa: F401
