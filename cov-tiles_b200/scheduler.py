"""One-process-per-GPU view of the batch scheduler: which contiguous tile range of ONE host batch a rank decodes.
The split is covt_partition_tiles — the same call the in-library scheduler (covt_decode_batch_multi, csrc/covt_multi.cu: one
process, N GPUs) makes; bench.py uses rank_slice under torchrun, tests/test_partition_gloo.py covers it with gloo on CPU.

Tiles share nothing (CovtParser.java:56-131 keeps no cross-layer state; every delta chain starts from 0,
DecodingUtils.java:57,97-98,396), so the data path has NO collective: rank g decodes the contiguous tile range
[starts[g], starts[g+1]) balanced by payload bytes (covt_partition_tiles: prefix sum over tile_offsets), with its own
context, stream and output arena. torch.distributed is used only for the barrier around the timed region and for
summing the per-rank counters (bench.py).
"""
import numpy as np

from . import abi


def partition(tile_offsets, world_size):
    """-> uint32[world_size + 1] tile range starts (host only; covt_partition_tiles in libcovt_b200)."""
    from . import partition_tiles
    return partition_tiles(tile_offsets, world_size)


def rank_slice(blob, tile_offsets, rank, world_size):
    """The part of a host batch rank `rank` decodes: (blob view, re-based tile_offsets, first tile index)."""
    offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
    starts = partition(offs, world_size)
    t0, t1 = int(starts[rank]), int(starts[rank + 1])
    b0, b1 = int(offs[t0]), int(offs[t1])
    return blob[b0:b1], offs[t0:t1 + 1] - np.uint64(b0), t0


def decode_partitioned(decoder, blob, tile_offsets, rank, world_size, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT):
    """Decode this rank's share of a batch on this rank's GPU. Returns (Result, first tile index)."""
    sub, offs, t0 = rank_slice(blob, tile_offsets, rank, world_size)
    return decoder.decode_batch(sub, offs, container, flags), t0
