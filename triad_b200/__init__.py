"""triad_b200 — B200-native (sm_100a) implementation of TRIAD's dense max-mean similarity +
symmetric InfoNCE hot path, behind the reference's own method signatures.

Importing the package does not load CUDA; the first op does, and fails loudly if
``triad_b200/lib/libtriad_b200.so`` has not been built (there is no CPU / PyTorch fallback).
"""
from .model import TokenSims, TriadHotPath, TriadSimilarityMixin  # noqa: F401
from . import ops, retrieval  # noqa: F401

__all__ = ["TokenSims", "TriadHotPath", "TriadSimilarityMixin", "ops", "retrieval"]
__version__ = "0.1.0"
