"""Seeded synthetic inputs of the shapes BASELINE.json names (no datasets on disk).

Entity / relation tables ~ N(0, sigma) (LibKGE default `lookup_embedder.yaml` initialize: normal_), RotatE
relation phases ~ U(-pi, pi) (`rotate.yaml:22-26`), triples uniform over (E, R, E).  Generated on the CPU with
fixed seeds so that every consumer — bench.py, the kernel micro-benchmarks, the tests and their CPU checker —
sees the same numbers; tests/test_host_logic.py pins this module against the checker's own copy.
"""
from __future__ import annotations

import math

import torch


def relation_dim(model: str, dim: int) -> int:
    if model in ("cp", "rotate"):
        return dim // 2
    if model == "rescal":
        return dim * dim
    return dim


def make_tables(model: str, E: int, R: int, D: int, sigma: float = 1.0, seed: int = 1234, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn((E, D), generator=g, dtype=torch.float32) * sigma
    dr = relation_dim(model, D)
    if model == "rotate":
        rel = (torch.rand((R, dr), generator=g, dtype=torch.float32) * 2.0 - 1.0) * math.pi
    else:
        rel = torch.randn((R, dr), generator=g, dtype=torch.float32) * sigma
    return ent.to(dtype), rel.to(dtype)


def make_triples(E: int, R: int, n: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, E, (n,), generator=g), torch.randint(0, R, (n,), generator=g),
                        torch.randint(0, E, (n,), generator=g)], 1)
