"""slaf_b200 -- B200-native (sm_100a) implementation of SLAF's ML-dataloader hot path.

COO expression rows (cell_integer_id, gene_integer_id, value) -> per-cell token tensors
(Geneformer rank-value / scGPT dual stream), behind the reference's own Window / SLAFTokenizer /
PrefetchBatchProcessor / SLAFDataLoader interfaces.  All hot-path computation happens in the
hand-written CUDA kernels of ``csrc/`` through the C ABI in ``include/slaf_b200.h``; there is
no CPU fallback and no other backend.
"""

__version__ = "0.1.0"

from . import _native  # noqa: F401


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libslaf_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    return _native.build(force=force, verbose=verbose)
