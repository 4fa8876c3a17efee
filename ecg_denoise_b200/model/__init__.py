"""B200-native mirrors of the reference's `model` package: `transformer`, `raletransformer`, `ralenet_12leads`."""
from . import transformer, raletransformer, ralenet_12leads  # noqa: F401
