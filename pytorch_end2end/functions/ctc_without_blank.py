"""Drop-in for pytorch_end2end/functions/ctc_without_blank.py (reference :120-143)."""
from end2end_b200.functions.ctc_without_blank import CTCWithoutBlankLossFunction, ctc_without_blank_3d_loss  # noqa: F401
