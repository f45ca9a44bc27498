"""Named-axis tensors at the engine boundary.

The reference carries plate and K axes as ``functorch.dim`` first-class dims
(reference: src/alan/utils.py:6-11).  At the B200 boundary every tensor is a plain
contiguous ``torch.Tensor`` whose LEADING dims are named (plate names such as
``'plate_1'`` or K axes such as ``'K_z'``) and whose trailing dims are positional
(batch/event dims, e.g. the ``d=18`` of MovieLens) -- SURVEY.md Appendix B.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import torch


@dataclass
class NT:
    """A plain tensor plus the names of its leading (named) dims."""
    t: torch.Tensor
    axes: tuple

    def __post_init__(self):
        self.axes = tuple(self.axes)
        if len(set(self.axes)) != len(self.axes):
            raise Exception(f"duplicate named axes {self.axes}")
        if self.t.ndim < len(self.axes):
            raise Exception(f"tensor of shape {tuple(self.t.shape)} cannot carry named axes {self.axes}")

    @property
    def pos_shape(self) -> tuple:
        return tuple(self.t.shape[len(self.axes):])

    @property
    def named_sizes(self) -> dict:
        return {a: int(s) for a, s in zip(self.axes, self.t.shape)}

    def order(self, axes: Sequence[str]) -> "NT":
        """Permute the named dims into `axes` order (must be the same set)."""
        axes = tuple(axes)
        if axes == self.axes:
            return self
        if set(axes) != set(self.axes):
            raise Exception(f"order(): {axes} is not a permutation of {self.axes}")
        perm = [self.axes.index(a) for a in axes] + list(range(len(self.axes), self.t.ndim))
        return NT(self.t.permute(perm), axes)

    def detach(self) -> "NT":
        return NT(self.t.detach(), self.axes)

    def to(self, *args, **kwargs) -> "NT":
        return NT(self.t.to(*args, **kwargs), self.axes)


def from_torch_named(x: torch.Tensor) -> NT:
    """torch named tensor (names = plate/K names, None = positional) -> NT.

    Mirrors the user-facing convention of the reference, where data / inputs /
    parameters are named tensors (reference: src/alan/Problem.py:26-27).
    Named dims must come first, as the reference requires (utils.py named2dim).
    """
    names = list(x.names)
    k = 0
    while k < len(names) and names[k] is not None:
        k += 1
    if any(n is not None for n in names[k:]):
        raise Exception(f"named dims must precede positional dims, got {x.names}")
    return NT(x.rename(None), tuple(names[:k]))
