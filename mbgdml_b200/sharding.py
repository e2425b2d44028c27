"""Fragment partition across GPUs: the rule the kernels implement (csrc/enumerate.cuh,
``global_candidate``), restated for the host so that tests can check it.

The global candidate space of a job (n_frames x tuples, frame-major) is cut into granules
of GRANULE consecutive candidates; shard ``rank`` of ``world`` owns granules
rank, rank + world, rank + 2*world, ...  No data-path communication is needed: every
shard produces partial E/F sums, combined by one all-reduce.
"""
from __future__ import annotations

import numpy as np

GRANULE = 16  # csrc/common.cuh: kShardGranule


def owner_of(candidate_index, world):
    """Shard that evaluates global candidate ``candidate_index``."""
    return (np.asarray(candidate_index, dtype=np.int64) // GRANULE) % int(world)


def shard_mask(n_candidates, rank, world):
    """Boolean mask over the global candidate space of the candidates shard ``rank`` owns."""
    return owner_of(np.arange(int(n_candidates), dtype=np.int64), world) == int(rank)
