"""Drop-in `KNearestNeighbor` (reference: lib/knn/__init__.py:9-23).

`KNearestNeighbor(k)(ref (B,D,R), query (B,D,Q)) -> LongTensor (B,k,Q)`, indices 1-BASED and bit-identical
to the reference CUDA kernels.  The reference subclasses the legacy (instance-style) autograd.Function;
modern torch rejects that, so this is a plain callable object with the same constructor and call
signature -- there is nothing to differentiate (integer output)."""
from __future__ import annotations

import torch

from ... import ops


class KNearestNeighbor:
    """Compute k nearest neighbors for each query point."""

    def __init__(self, k: int):
        self.k = int(k)

    def forward(self, ref: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
        # the reference moves its inputs to the GPU itself (`.float().cuda()`, lib/knn/__init__.py:16-17)
        ref = ref.detach().float().cuda()
        query = query.detach().float().cuda()
        return ops.knn(ref, query, self.k)

    __call__ = forward
