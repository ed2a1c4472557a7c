"""Metadata filters as a row gate (the tail of that row).

The reference can restrict a vector-store query to nodes whose metadata match a filter:
``_build_metadata_filters(filter_spec)`` (/root/reference/src/tensortruth/rag_engine.py:301-365) turns a dict like
``{"doc_type": "library", "version": {"$gte": "2.0"}, "lang": ["en", "de"]}`` into LlamaIndex ``MetadataFilters``
(AND of the clauses) for ``index.as_retriever(similarity_top_k=k, filters=...)``; ``ChromaVectorStore`` hands them to
Chroma as a ``where`` clause.  (On the retrieval path of today's reference nothing passes a filter -- the function is
dead code there -- so this is the natural place for it, not a behaviour the callers rely on yet.)

Here the clauses are evaluated ONCE per filter on the host over the leaves' metadata and become a *row gate*: a copy
of the index's ``inv_norm`` array in which the excluded rows hold NaN.  Every stage-1 kernel then never shortlists
them (a NaN score passes no comparison), the certificate's thresholds bound the dropped eligible rows only, and the
exact fp64 scan reads the same array as its gate (``tt_scan_exact_f64_gated``).  No kernel has a filter code path of
its own and a filtered search streams the corpus at the same speed as an unfiltered one.

Operator semantics are Chroma's ``where`` semantics as the LlamaIndex Chroma adapter maps them **[U: restated from the
public sources, not verifiable in this image]**: ``$eq $ne $gt $gte $lt $lte $in $nin``; a node that lacks the key
matches only ``$ne`` / ``$nin``.  ``$contains`` (substring of a string value, member of a list value) is an extension
the Chroma adapter does not offer; ``$text_match`` is refused, as that adapter refuses it.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

Clause = Tuple[str, str, Any]  # (metadata key, operator, operand)

# LlamaIndex FilterOperator values -> the reference's spelling (rag_engine.py:286-298)
_LLAMA_OPS = {"==": "$eq", "!=": "$ne", ">": "$gt", ">=": "$gte", "<": "$lt", "<=": "$lte", "in": "$in", "nin": "$nin",
              "contains": "$contains", "text_match": "$text_match"}
_KNOWN = set(_LLAMA_OPS.values())


def clauses_from_spec(filter_spec: Optional[Dict[str, Any]]) -> List[Clause]:
    """The reference's ``_build_metadata_filters`` on its own input format: a scalar is an equality, a list is ``$in``, a
    dict is ``{"$op": operand}`` of which only the FIRST key counts and an unknown operator drops the clause
    (rag_engine.py:333-358); the clauses are AND-ed."""
    out: List[Clause] = []
    for key, value in (filter_spec or {}).items():
        if isinstance(value, dict):
            for op, operand in value.items():
                if op in _KNOWN:
                    out.append((key, op, operand))
                break
        elif isinstance(value, list):
            out.append((key, "$in", value))
        else:
            out.append((key, "$eq", value))
    return out


def clauses_from_filters(filters: Any) -> List[Clause]:
    """Clauses of a filter in either form: the reference's filter-spec dict, or a LlamaIndex ``MetadataFilters`` (duck-typed:
    ``.filters`` of objects with ``key`` / ``value`` / ``operator``, ``.condition`` AND)."""
    if filters is None:
        return []
    if isinstance(filters, dict):
        return clauses_from_spec(filters)
    cond = getattr(filters, "condition", None)
    cond = getattr(cond, "value", cond)
    if cond not in (None, "and", "AND"):
        raise ValueError(f"filter condition {cond!r}: the reference builds AND filters only (rag_engine.py:361-364)")
    out: List[Clause] = []
    for f in getattr(filters, "filters", []) or []:
        if hasattr(f, "filters"):
            raise ValueError("nested MetadataFilters are not supported")
        op = getattr(f, "operator", "==")
        op = _LLAMA_OPS.get(getattr(op, "value", op), getattr(op, "value", op))
        if op not in _KNOWN:
            raise ValueError(f"filter operator {op!r} is not supported")
        out.append((f.key, op, f.value))
    return out


def _match(meta: Dict[str, Any], clause: Clause) -> bool:
    key, op, operand = clause
    if op == "$text_match":
        raise ValueError("$text_match is not supported (the Chroma adapter the reference uses refuses it as well)")
    if key not in meta or meta[key] is None:
        return op in ("$ne", "$nin")
    v = meta[key]
    try:
        if op == "$eq":
            return v == operand
        if op == "$ne":
            return v != operand
        if op == "$gt":
            return v > operand
        if op == "$gte":
            return v >= operand
        if op == "$lt":
            return v < operand
        if op == "$lte":
            return v <= operand
        if op == "$in":
            return v in operand
        if op == "$nin":
            return v not in operand
        if op == "$contains":
            return operand in v
    except TypeError:  # e.g. str against int: Chroma compares within one type only -- no match
        return op in ("$ne", "$nin")
    raise ValueError(f"unknown filter operator {op!r}")


def eligible_rows(filters: Any, leaf_metadata: Sequence[Optional[Dict[str, Any]]]) -> np.ndarray:
    """bool ``[n_rows]``: which leaves (in corpus-row order) satisfy every clause."""
    clauses = clauses_from_filters(filters)
    out = np.ones(len(leaf_metadata), dtype=bool)
    if not clauses:
        return out
    for i, meta in enumerate(leaf_metadata):
        m = meta or {}
        out[i] = all(_match(m, c) for c in clauses)
    return out


class RowFilter:
    """One filter resolved against one ``DeviceIndex``: the gated ``inv_norm`` (NaN = excluded row) resident in HBM.
    Build it once per (index, filter) -- ``DeviceIndex.row_filter(...)`` caches by key -- and pass it to ``search`` /
    ``retrieve_host``."""

    def __init__(self, index, eligible: np.ndarray, key: Any = None):
        import torch

        eligible = np.ascontiguousarray(eligible, dtype=bool)
        if eligible.shape != (index.n_rows,):
            raise ValueError(f"eligibility mask must be bool[{index.n_rows}], got {eligible.shape}")
        self.key = key
        self.n_eligible = int(eligible.sum())
        gate = index.inv_norm.clone()
        if self.n_eligible < index.n_rows:
            gate[torch.from_numpy(~eligible).to(index.device)] = float("nan")
        self.inv_norm = gate
