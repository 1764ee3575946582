"""Drop-in for pytorch_end2end/utils/alignment.py (reference :109-138): ``get_alignment_3d`` on the B200-native engine."""
from end2end_b200.utils.alignment import get_alignment_3d, get_alignment_3d_device  # noqa: F401
