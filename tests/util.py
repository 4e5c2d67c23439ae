"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import strajnet_oracle as O


def sub(weights, prefix):
    """Oracle weight dict restricted to `prefix`, with the prefix stripped."""
    return {k[len(prefix):]: v for k, v in weights.items() if k.startswith(prefix)}


def randn(shape, seed=0, scale=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))


def max_abs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


_model_cache = {}


def oracle_model(seed=0, fg_msa=True, fg=True, cfg=None):
    key = (seed, fg_msa, fg)
    if key not in _model_cache:
        _model_cache[key] = O.make_weights(cfg or O.CFG256, seed=seed, fg_msa=fg_msa, fg=fg)
    return _model_cache[key]
