"""Drop-in for pytorch_end2end/modules/ctc_loss.py (reference :16-75)."""
from end2end_b200.modules.ctc_loss import CTCLoss, ForwardBackwardLossBase  # noqa: F401
