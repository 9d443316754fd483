"""CPU, world_size 2 over gloo: the N>1 path of the batch scheduler — contiguous tile ranges by tile index,
no data-path collective; the only collectives are the barrier and the sum of per-rank counters that bench.py uses.
Each rank checks ITS share with the oracle (the product needs a GPU); the union must equal the whole-batch result."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import covt_loader
    covt = covt_loader.load()
    from cov_tiles_b200 import scheduler
    from oracle import oracle as O
    from tools.gen import gen as G
    abi = covt.abi
    blob, offs, truth = G.tiles(77, 500, G.default_params())   # every rank can regenerate the batch: seed = tile index
    sub, sub_offs, t0 = scheduler.rank_slice(blob, offs, rank, world)
    ref = O.decode_batch(sub, sub_offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    assert np.all(ref.tile_status == 0)
    counters = torch.tensor([ref.payload_bytes, ref.vertices, len(sub_offs) - 1, int(ref.layers["n_rings"].sum()),
                             int(ref.layers["n_parts"].sum())], dtype=torch.int64)
    dist.barrier()
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    # a rank's layers equal the corresponding rows of the whole-batch decode (tile index re-based by t0)
    whole = O.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    rows = whole.layers[(whole.layers["tile"] >= t0) & (whole.layers["tile"] < t0 + len(sub_offs) - 1)]
    assert len(rows) == len(ref.layers)
    assert np.array_equal(rows["tile"] - t0, ref.layers["tile"]) and np.array_equal(rows["n_vertices"], ref.layers["n_vertices"])
    o0 = int(rows["out"][0][abi.BUF_A_COORDS]) if len(rows) else 0
    n = int(ref.layers["n_coords"].sum())
    got = ref.buffer(abi.BUF_A_COORDS)
    want = whole.buffer(abi.BUF_A_COORDS)[o0:o0 + len(got)]
    assert np.array_equal(got, want) and n > 0
    if rank == 0:
        q.put((counters.tolist(), whole.payload_bytes, whole.vertices, truth))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_the_batch_without_a_data_collective():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    counters, payload, vertices, truth = q.get()
    assert counters[0] == payload and counters[1] == vertices == truth["vertices"]
    assert counters[2] == 500 and counters[3] == truth["rings"] and counters[4] == truth["parts"]
