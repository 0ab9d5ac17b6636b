"""colbert.jl_b200 -- B200-native search-time scoring path for ColBERT.jl indexes.

Importable as `colbert_jl_b200` (see colbert_jl_b200.py at the repo root).  The compute lives in
csrc/ (hand-written sm_100a CUDA behind the C ABI of include/colbert_b200.h, built in-tree into
lib/libcolbert_b200.so); this package is the host-side mirror of the reference's interface."""
from ._lib import (LIB_PATH, SIGNATURES, BoundsError, ColBERTB200Error, CudaError, DimensionMismatch, DomainError,
                   OutOfMemory, Unsupported, load)
from .searcher import (ColBERTConfig, MultiSearcher, Searcher, _build_emb2pid, compress, decompress, load_object, maxsim, merge_topk, retrieve, search)

__all__ = ["ColBERTConfig", "Searcher", "MultiSearcher", "search", "retrieve", "decompress", "compress", "maxsim", "merge_topk", "load_object",
           "_build_emb2pid", "load", "LIB_PATH", "SIGNATURES", "ColBERTB200Error", "DimensionMismatch",
           "DomainError", "CudaError", "OutOfMemory", "Unsupported", "BoundsError"]
