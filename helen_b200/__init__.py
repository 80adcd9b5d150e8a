"""helen_b200 -- B200-native implementation of HELEN's call_consensus / predict hot path.

Host side mirrors the reference's Python surface for this path (TransducerGRU,
ModelHandler, SequenceDataset, DataStore, predict / predict_gpu, call_consensus,
polish_genome, the ``helen`` CLI); all arithmetic runs in the in-tree CUDA library
``helen_b200/lib/libhelen_b200.so`` through the C ABI declared in ``include/helen_b200.h``.
There is no CPU fallback: without the built library or without an sm_100 GPU the
compute entry points raise.
"""
from .options import ImageSizeOptions, TrainOptions  # noqa: F401

__version__ = "0.1.0"
