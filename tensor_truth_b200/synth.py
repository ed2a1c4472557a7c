"""Seeded synthetic workload for the retrieval hot path.

Stands in for what the reference's index build would have produced
(/root/reference/src/tensortruth/indexing/builder.py:385-442: a hierarchical node
tree, leaves embedded with BAAI/bge-m3 -> 1024-d unit vectors, builder.py:38).

* Tree: ``tree.build_uniform_tree`` (fan-out uniform{2..6}).
* Embeddings are *clustered along the tree* so that auto-merges actually fire
  (iid vectors never share a parent inside a top-10): level-0 centroids ~ N(0, I),
  child = parent + sigma(depth) * N(0, I) with sigma = 0.6, 0.4, 0.25, 0.15 by depth;
  leaves are L2-normalised and **rounded to bf16 -- the rounded values are the
  canonical corpus** (what both the oracle and the kernels consume).
* Determinism: every level is generated in fixed blocks of ``block`` nodes, block b of
  level l seeded from (seed, l, b), so a row's bytes do not depend on how the corpus is
  sharded over GPUs.  The stream depends on the device *type* (torch's CPU and CUDA
  generators differ): tests generate on CPU and upload; the bench generates on the GPU
  and hands the CPU arm a device->host copy of the same bytes.
* Ties: rows with ``ordinal % 1024 == 1023`` are verbatim copies of the previous row.
* Queries: ``normalise(leaf_t + 0.3/sqrt(D) * N(0, I))``, t uniform, query j seeded (seed, j).
"""

from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from .tree import NodeTree, build_uniform_tree

_SIGMA_BY_DEPTH = (None, 0.6, 0.4, 0.25, 0.15)


class SynthCorpus:
    def __init__(self, n_leaf: int, dim: int = 1024, levels: int = 3, seed: int = 1234,
                 device: str | torch.device = "cpu", block: int = 65536,
                 tree: Optional[NodeTree] = None):
        if block % 1024:
            raise ValueError("block must be a multiple of 1024 (tie rule)")
        if levels > len(_SIGMA_BY_DEPTH):
            raise ValueError("too many levels")
        self.n_leaf, self.dim, self.levels, self.seed = int(n_leaf), int(dim), int(levels), int(seed)
        self.device = torch.device(device)
        self.block = int(block)
        self.tree = tree if tree is not None else build_uniform_tree(n_leaf, levels, seed)
        offs = list(self.tree.level_offsets) + [self.tree.n_nodes]
        self._lo = offs
        self._cache: dict = {}

    # ---------------------------------------------------------------- internals
    def _level_size(self, lv: int) -> int:
        return self._lo[lv + 1] - self._lo[lv]

    def _noise(self, lv: int, blk: int, rows: int) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        g.manual_seed((self.seed * 1_000_003 + lv * 7_919 + 1) * 2_147_483_629 % (2**62) + blk)
        z = torch.randn(self.block, self.dim, generator=g, device=self.device, dtype=torch.float32)
        return z[:rows]

    def _level_block(self, lv: int, blk: int) -> torch.Tensor:
        """fp32 vectors (un-normalised) of nodes [blk*block, ...) of level ``lv`` (0 = leaves)."""
        key = (lv, blk)
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        a = blk * self.block
        b = min(self._level_size(lv), a + self.block)
        z = self._noise(lv, blk, b - a)
        if lv == self.levels - 1:
            v = z
        else:
            par = self.tree.parent_of[self._lo[lv] + a:self._lo[lv] + b].astype(np.int64) - self._lo[lv + 1]
            pa, pb = int(par[0]), int(par[-1]) + 1
            pv = self._level_range(lv + 1, pa, pb)
            idx = torch.from_numpy(par - pa).to(self.device)
            sigma = _SIGMA_BY_DEPTH[(self.levels - 1) - lv]
            v = pv.index_select(0, idx) + sigma * z
        # keep at most two blocks per level (generation walks the levels in order)
        for k in [k for k in self._cache if k[0] == lv and abs(k[1] - blk) > 1]:
            del self._cache[k]
        if lv > 0:
            self._cache[key] = v
        return v

    def _level_range(self, lv: int, a: int, b: int) -> torch.Tensor:
        parts = []
        for blk in range(a // self.block, (b - 1) // self.block + 1):
            base = blk * self.block
            v = self._level_block(lv, blk)
            parts.append(v[max(a, base) - base:min(b, base + self.block) - base])
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)

    # ---------------------------------------------------------------- corpus
    def leaf_block(self, blk: int):
        """Canonical rows of leaf block ``blk``: ``(bf16 [rows, D], inv_norm fp32 [rows])``."""
        v = self._level_block(0, blk)
        v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)
        c = v.to(torch.bfloat16)
        rows = c.shape[0]
        dup = torch.arange(1023, max(rows, 1023), 1024, device=self.device)
        if dup.numel():
            c[dup] = c[dup - 1]
        cf = c.to(torch.float32)
        inv_norm = (cf * cf).sum(dim=1).clamp_min(1e-30).rsqrt()
        return c, inv_norm

    def rows(self, lo: int, hi: int, out: Optional[torch.Tensor] = None,
             out_inv_norm: Optional[torch.Tensor] = None):
        """Fill/return canonical rows ``[lo, hi)`` (bf16) and their fp32 inverse norms."""
        n = hi - lo
        if out is None:
            out = torch.empty((n, self.dim), dtype=torch.bfloat16, device=self.device)
        if out_inv_norm is None:
            out_inv_norm = torch.empty((n,), dtype=torch.float32, device=self.device)
        if n == 0:
            return out, out_inv_norm
        for blk in range(lo // self.block, (hi - 1) // self.block + 1):
            base = blk * self.block
            c, inv = self.leaf_block(blk)
            a, b = max(lo, base), min(hi, base + self.block)
            out[a - lo:b - lo] = c[a - base:b - base]
            out_inv_norm[a - lo:b - lo] = inv[a - base:b - base]
        self._cache.clear()
        return out, out_inv_norm

    # ---------------------------------------------------------------- queries
    def query_targets(self, n_q: int, first: int = 0) -> np.ndarray:
        t = np.empty(n_q, dtype=np.int64)
        for j in range(n_q):
            t[j] = np.random.default_rng([self.seed, 7, first + j]).integers(0, self.n_leaf)
        return t

    def queries(self, n_q: int, first: int = 0,
                lookup: Optional[Callable[[int], Optional[torch.Tensor]]] = None) -> torch.Tensor:
        """fp32 ``[n_q, D]`` on CPU.  ``lookup(t)`` returns canonical row t (any float dtype, any
        device) or None when this process does not own it (the row then contributes zeros and the
        caller all-reduces across shards)."""
        tgt = self.query_targets(n_q, first)
        out = torch.zeros((n_q, self.dim), dtype=torch.float32)
        for j in range(n_q):
            t = int(tgt[j])
            row = lookup(t) if lookup is not None else self.rows(t, t + 1)[0][0]
            if row is None:
                continue
            g = torch.Generator(device="cpu")
            g.manual_seed(self.seed * 7_368_787 + 11 + first + j)
            z = torch.randn(self.dim, generator=g, dtype=torch.float32)
            out[j] = row.detach().to("cpu", torch.float32) + (0.3 / math.sqrt(self.dim)) * z
        return out

    @staticmethod
    def finish_queries(q: torch.Tensor) -> torch.Tensor:
        return q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)


def make_small(n_leaf: int, n_q: int, dim: int = 1024, levels: int = 3, seed: int = 1234):
    """CPU materialisation for tests: ``(tree, corpus_bits uint16 [N,D], inv_norm f32 [N], queries f32 [B,D])``."""
    sc = SynthCorpus(n_leaf, dim, levels, seed, device="cpu")
    c, inv = sc.rows(0, n_leaf)
    cf = c.to(torch.float32)
    q = sc.finish_queries(sc.queries(n_q, lookup=lambda t: cf[t]))
    bits = c.view(torch.int16).numpy().view(np.uint16).copy()
    return sc.tree, bits, inv.numpy().copy(), q.numpy().copy()
