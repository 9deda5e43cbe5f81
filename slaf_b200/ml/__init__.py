"""Drop-in mirrors of the reference's ``slaf.ml`` interfaces for the hot path
(window functions, tokenizers, shuffle, prefetch batch processor, dataloader)."""

from .aggregators import GeneformerWindow, ScGPTWindow, SimpleWindow, Window, WindowType, create_window  # noqa: F401
from .dataloaders import SLAFDataLoader, get_device_info, get_optimal_device  # noqa: F401
from .datasets import (AsyncPrefetcher, PrefetchBatch, PrefetchBatchProcessor, RawPrefetchBatch, SLAFIterableDataset,  # noqa: F401
                       TokenizedPrefetchBatch)
from .samplers import RandomShuffle, Shuffle, ShuffleType, StratifiedShuffle, create_shuffle  # noqa: F401
from .tokenizers import GeneformerTokenizer, ScGPTTokenizer, SLAFTokenizer, TokenizerType  # noqa: F401
