"""Embedding layers for masked autoregressive flows (reference tfep/nn/embeddings/__init__.py)."""

from .mafembed import FlipInvariantEmbedding, MAFEmbedding, MixedEmbedding, PeriodicEmbedding
