"""Reference-compatible import path ``pytorch_end2end.functions`` (reference: pytorch_end2end/functions/)."""
from end2end_b200.functions.forward_backward import ForwardBackwardLossFunction  # noqa: F401
