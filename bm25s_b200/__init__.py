"""bm25s_b200: the bm25s query-time hot path (CSC column sums + top-k, batched) on B200 behind the reference API."""
from . import selection, sharded, synth, tokenization
from ._lib import Builder, DeviceGroup, DeviceIndex, build_extension, device_count, topk_merge
from .bm25 import BM25, Results, __version__
from .tokenization import Tokenized

__all__ = ["BM25", "Results", "Tokenized", "DeviceIndex", "DeviceGroup", "Builder", "sharded", "selection", "tokenization", "synth", "topk_merge",
           "device_count", "build_extension", "__version__"]
