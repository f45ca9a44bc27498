"""Contraction-order planning for the log-semiring K contraction.

The reference delegates *only the order* of pairwise contractions to
``opt_einsum.contract_path`` (reference: src/alan/reduce_Ks.py:264-265); no
arithmetic is delegated.  ``opt_einsum`` is an unpinned third-party dependency
(setup.py:12) that is not vendored, so the order is re-derived here with a
deterministic greedy rule.  Any valid order is semantically equal up to O(eps)
(SURVEY.md Appendix A4); a *bad* order explodes memory, so the rule minimises
the size of each intermediate.

A path is a list of index tuples into the *current* operand list; contracted
operands are removed and the result is appended at the end, which is exactly
how ``collect_lps`` consumes it (reduce_Ks.py:270-281).
"""
from __future__ import annotations

from typing import Hashable, Iterable, Sequence


def _size(axes: Iterable[Hashable], sizes: dict) -> int:
    n = 1
    for a in axes:
        n *= int(sizes[a])
    return n


def greedy_path(operand_axes: Sequence[Iterable[Hashable]], sum_axes: Iterable[Hashable],
                sizes: dict) -> list[tuple[int, ...]]:
    """Greedy pairwise contraction path.

    operand_axes: axes carried by every operand (plate axes and K axes alike).
    sum_axes:     the K axes that are summed somewhere along the path.
    sizes:        axis -> extent.

    At each step pick the pair (i, j) minimising
    ``size(result) - size(i) - size(j)``; ties go to the smaller union and then
    to the lexicographically smallest (i, j), so the path is a pure function of
    its inputs (the resampling walk relies on this).
    """
    ops = [frozenset(a) for a in operand_axes]
    sum_axes = frozenset(sum_axes)
    n = len(ops)
    if n == 0:
        return []
    if n <= 2:
        return [tuple(range(n))]

    path = []
    while len(ops) > 1:
        best = None
        for i in range(len(ops)):
            for j in range(i + 1, len(ops)):
                union = ops[i] | ops[j]
                others = frozenset().union(*[ops[k] for k in range(len(ops)) if k != i and k != j]) \
                    if len(ops) > 2 else frozenset()
                removed = frozenset(a for a in union if a in sum_axes and a not in others)
                result = union - removed
                cost = _size(result, sizes) - _size(ops[i], sizes) - _size(ops[j], sizes)
                key = (cost, _size(union, sizes), i, j)
                if best is None or key < best[0]:
                    best = (key, i, j, result)
        _, i, j, result = best
        path.append((i, j))
        ops = [ops[k] for k in range(len(ops)) if k != i and k != j] + [result]
    return path
