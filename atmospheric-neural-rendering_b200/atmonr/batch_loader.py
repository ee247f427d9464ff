"""Batch-wise loader over a dataset's ray (or voxel) index (reference: src/atmonr/batch_loader.py).

Same constructor and iteration protocol. The reference builds every batch index from a Python
list of B ints (batch_loader.py:45-49); here one permutation per epoch is drawn as a tensor
(torch's generators, drawn exactly like the reference's RandomSampler does, so `torch.manual_seed`
reproduces the reference's batch order) and sliced, which removes the O(B) Python
work per step. Under torch.distributed every rank draws the SAME permutation and
takes its own contiguous slice of each global batch (data-parallel ray sharding). Every global batch
is cut to a multiple of world_size (at most world_size - 1 rays of the ragged last batch of an epoch
are left out, and a last batch smaller than world_size is skipped on ALL ranks), so the shards are
equal-sized: no rank ever sees an empty shard or skips the gradient collective, and the mean of the
per-rank mean losses IS the mean over the global batch.
"""

from __future__ import annotations

from collections.abc import Iterator

import torch


class BatchLoader:
    def __init__(self, dataset, batch_size: int, shuffle: bool = True, drop_last: bool = False,
                 rank: int = 0, world_size: int = 1, seed: int | None = None):
        self.dataset = dataset
        self.idx = dataset.ray_idx if hasattr(dataset, "ray_idx") else dataset.idx
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.rank, self.world_size = rank, world_size
        self.generator = None
        if seed is not None or world_size > 1:
            self.generator = torch.Generator().manual_seed(0 if seed is None else seed)

    def __len__(self) -> int:
        n = self.idx.shape[0]
        full, tail = divmod(n, self.batch_size)
        if self.drop_last or tail < max(1, self.world_size):   # a tail below world_size cannot be sharded
            return full
        return full + 1

    def __iter__(self) -> Iterator[dict[str, torch.Tensor]]:
        n = self.idx.shape[0]
        if not self.shuffle:
            order = torch.arange(n)
        else:
            gen = self.generator
            if gen is None:
                # what the reference's RandomSampler does (batch_loader.py:34-38 -> torch RandomSampler): one
                # int64 from the global generator seeds a private one, so `torch.manual_seed(s)` gives the
                # SAME batch order as the reference and consumes the global stream identically
                gen = torch.Generator()
                gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            order = torch.randperm(n, generator=gen)
        # ONE host->device copy of the permutation per epoch; every batch index is a device-side slice
        # of it (a per-batch copy from pageable memory blocks the host until the device has caught up)
        order = order.to(self.idx.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            if self.world_size > 1:
                per = idx.shape[0] // self.world_size              # equal shards; < world_size rays left out
                idx = idx[self.rank * per:(self.rank + 1) * per]
            yield self.dataset.__getbatch__(idx)
