"""Labelled square matrix of transition counts or probabilities.

Same public surface as the reference's ``TransitionTable`` (reference transition_table.py:18-234): ``keys``,
``matrix`` (read-only), ``get``, ``to_dict``, ``reorder``, ``filter``, equality, text / HTML rendering,
immutability after construction, and the same ``ValueError`` / ``KeyError`` conditions.  It is the return type of
``Trajectory.transition_counts_by_site`` and friends (reference trajectory.py:188-321).
"""
from __future__ import annotations

import html
from typing import Sequence

import numpy as np


class TransitionTable:
    __slots__ = ("_keys", "_matrix", "_where", "_sealed")
    __hash__ = None  # type: ignore[assignment]

    def __init__(self, keys, matrix) -> None:
        m = np.array(matrix, copy=True)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError(f"matrix must be a square 2-D array, got shape {m.shape}")
        if len(keys) != m.shape[0]:
            raise ValueError(f"len(keys) ({len(keys)}) != matrix dimension ({m.shape[0]})")
        if not np.issubdtype(m.dtype, np.number):
            raise ValueError(f"matrix must have a numeric dtype, got {m.dtype}")
        where = {k: i for i, k in enumerate(keys)}
        if len(where) != len(keys):
            raise ValueError("keys must not contain duplicates")
        m.flags.writeable = False
        object.__setattr__(self, "_keys", keys)
        object.__setattr__(self, "_matrix", m)
        object.__setattr__(self, "_where", where)
        object.__setattr__(self, "_sealed", True)

    def __setattr__(self, name, value) -> None:
        raise AttributeError("TransitionTable is immutable")

    @property
    def keys(self):
        return self._keys

    @property
    def matrix(self) -> np.ndarray:
        return self._matrix

    def _row_of(self, key):
        if key not in self._where:
            raise KeyError(key)
        return self._where[key]

    def get(self, from_key, to_key):
        i = self._row_of(from_key)
        j = self._row_of(to_key)
        return self._matrix[i, j].item()

    def to_dict(self) -> dict:
        return {r: {c: self._matrix[i, j].item() for j, c in enumerate(self._keys)} for i, r in enumerate(self._keys)}

    def _take(self, keys) -> "TransitionTable":
        pick = [self._where[k] for k in keys]
        if not pick:
            return TransitionTable(keys=keys, matrix=np.empty((0, 0), dtype=self._matrix.dtype))
        return TransitionTable(keys=keys, matrix=self._matrix[np.ix_(pick, pick)])

    def reorder(self, keys: Sequence) -> "TransitionTable":
        wanted = tuple(keys)
        if len(wanted) != len(self._keys) or set(wanted) != set(self._keys):
            problems = []
            missing = [k for k in self._keys if k not in set(wanted)]
            unknown = [k for k in wanted if k not in self._where]
            if missing:
                problems.append(f"missing keys: {missing!r}")
            if unknown:
                problems.append(f"unknown keys: {unknown!r}")
            raise ValueError(f"keys must be a permutation of the current keys; {'; '.join(problems)}")
        return self._take(wanted)

    def filter(self, keys: Sequence) -> "TransitionTable":
        wanted = tuple(keys)
        if len(set(wanted)) != len(wanted):
            raise ValueError("keys must not contain duplicates")
        unknown = [k for k in wanted if k not in self._where]
        if unknown:
            raise ValueError(f"unknown keys: {unknown!r}")
        return self._take(wanted)

    def __eq__(self, other):
        if not isinstance(other, TransitionTable):
            return NotImplemented
        return self._keys == other._keys and np.array_equal(self._matrix, other._matrix)

    def _cells(self):
        spec = "d" if np.issubdtype(self._matrix.dtype, np.integer) else ".3f"
        return [str(k) for k in self._keys], [[format(v, spec) for v in row] for row in self._matrix]

    def __str__(self) -> str:
        if not self._keys:
            return ""
        names, cells = self._cells()
        w = max(len(t) for t in names + [v for row in cells for v in row])
        lines = [" " * w + "".join(t.rjust(w + 2) for t in names)]
        lines += [n.rjust(w) + "".join(v.rjust(w + 2) for v in row) for n, row in zip(names, cells)]
        return "\n".join(lines)

    def _repr_html_(self) -> str:
        if not self._keys:
            return ""
        names, cells = self._cells()
        names = [html.escape(n) for n in names]
        head = "<tr><th></th>" + "".join(f"<th>{n}</th>" for n in names) + "</tr>"
        body = "".join("<tr><th>" + n + "</th>" + "".join(f"<td>{v}</td>" for v in row) + "</tr>"
                       for n, row in zip(names, cells))
        return f"<table>{head}{body}</table>"

    def __repr__(self) -> str:
        return f"TransitionTable(keys={self._keys!r}, shape={self._matrix.shape[0]}x{self._matrix.shape[1]})"
