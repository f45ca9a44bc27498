"""`computation_strategy=` objects of the logPQ path.

Mirror of the reference's strategy module (src/alan/Split.py:7-71): the singletons `no_checkpoint` and
`checkpoint`, and `Split(platename, split_size)`; plus `B200`, the strategy object a reference maintainer passes
to route `Sample._elbo` into this engine.

What they mean on the B200 engine:
  * `no_checkpoint` / `checkpoint`: the engine never stores an intermediate at cells x event size and its adjoint
    program recomputes factor values in registers, so both select the same single pass over the plate tree
    (the reference's checkpoint exists to bound autograd's stored activations, logpq.py:62-66).
  * `Split(plate, n)`: the plate is processed in blocks of `n` elements through ONE block-sized workspace
    (engine.SplitRunner): block sizes follow the reference's rule (Split.py:84-95, including the "steal one
    element if the remainder is 1" adjustment), the `[K_parents]` tiles of the blocks are added left to right
    (logpq.py:151-153), and the forward of a block is recomputed before its adjoint -- the reference's
    per-chunk checkpoint.  Across GPUs the same blocks become shards (engine.Compiled(shard_plate=...)).
"""
from __future__ import annotations


class NoSplit:
    def split_args(self, name, sample, inputs_params, extra_log_factors, data, all_platedims):
        """Strategy protocol of the reference (Split.py:7-14): one chunk holding everything."""
        return [dict(sample=sample, inputs_params=inputs_params, extra_log_factors=extra_log_factors,
                     data=data, all_platedims=all_platedims)]


class NoCheckpoint(NoSplit):
    pass


class Checkpoint(NoSplit):
    pass


no_checkpoint = NoCheckpoint()
checkpoint = Checkpoint()


class Split:
    """Split(platename, split_size): same constructor contract as the reference (Split.py:24-42)."""
    def __init__(self, platename: str, split_size: int):
        assert isinstance(platename, str)
        assert isinstance(split_size, int)
        if split_size < 1:
            raise Exception("split_size must be a positive number of plate elements")
        self.platename = platename
        self.split_size = split_size

    def sizes(self, orig: int):
        """Chunk sizes [s, ..., s, rem] of a plate of extent `orig` (SplitDims, Split.py:84-95)."""
        size = self.split_size
        assert orig > size, f"Split: plate {self.platename} of extent {orig} is not larger than the split size {size}"
        sizes = [size] * (orig // size)
        if orig % size:
            sizes.append(orig % size)
        if size > 2 and len(sizes) > 1 and sizes[-1] == 1:
            sizes[-2] -= 1
            sizes[-1] += 1
        return sizes


def resolve(strategy):
    """None | no_checkpoint | checkpoint -> None (single pass); Split -> the Split; anything else raises."""
    if strategy is None or isinstance(strategy, NoSplit):
        return None
    if isinstance(strategy, Split):
        return strategy
    if type(strategy).__name__ in ("NoCheckpoint", "Checkpoint", "NoSplit"):      # the reference's own singletons
        return None
    if type(strategy).__name__ == "Split" and hasattr(strategy, "platename"):     # the reference's own Split
        return Split(strategy.platename, int(strategy.split_size))
    if type(strategy).__name__ == "B200":
        return resolve(getattr(strategy, "split", None))
    raise Exception(f"computation_strategy must be no_checkpoint, checkpoint, Split(...) or B200(...), got {strategy!r}")
