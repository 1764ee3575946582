"""Drop-in for pytorch_end2end/modules/ctc_without_blank.py (reference :7-35)."""
from end2end_b200.modules.ctc_without_blank import CTCWithoutBlankLoss  # noqa: F401
