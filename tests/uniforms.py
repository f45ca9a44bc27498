"""Explicit-uniform source shared by the golden generator, the oracle and the GPU tests.

Resampling parity is defined against explicit uniforms: sampling
step number `i` of the top-down walk consumes one float64 tensor u[batch plates..., N],
with the batch plate axes in canonical program order (Plate.all_platenames()).
"""
import torch as t


class UniformSource:
    def __init__(self, seed: int, N: int, platesizes: dict, plate_order):
        self.seed, self.N, self.count = seed, N, 0
        self.platesizes = platesizes
        self.plate_order = list(plate_order)

    def canon(self, batch_axes):
        return [p for p in self.plate_order if p in batch_axes]

    def draw(self, batch_axes):
        """returns (tensor[batch..., N] float64, axes tuple)"""
        axes = self.canon(batch_axes)
        g = t.Generator().manual_seed(self.seed * 7919 + self.count)
        self.count += 1
        u = t.rand(*[self.platesizes[a] for a in axes], self.N, dtype=t.float64, generator=g)
        return u, tuple(axes) + ('N',)
