"""GPU: parity at BASELINE.json's FULL sizes (configs 3 and 5), through the C ABI, against the oracle on the same inputs.

The per-layer comparison of test_gpu_batch.py is a Python loop; at 2 M layers the same check runs vectorised
(util.compare_results_bulk: every decoded stream and every assembled buffer, alignment padding masked out), followed by the
size-independent properties of the generator (vertex / part / ring totals and the closed-ring coordinate sums it recorded while
encoding). Needs ~45 GB of host memory for the two result sets; skipped (loudly) on a smaller box."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def test_config5_one_million_tiles_bit_exact(covt, oracle, gen, decoder):
    if _mem_available_gb() < 64:
        pytest.skip("needs 64 GB of free host memory for the oracle's and the GPU's 17 GB result sets")
    abi = covt.abi
    n_tiles = 1 << 20
    blob, offs, truth = gen.tiles(0, n_tiles, gen.default_params())
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    st, _ = res.tile_status()
    assert np.array_equal(st, ref.tile_status) and not st.any()
    n_layers, n_elems = util.compare_results_bulk(abi, res, ref)
    assert n_layers == 2 * n_tiles and n_elems > 4_000_000_000
    L = res.layers
    assert int(L["n_vertices"].sum()) == truth["vertices"]
    assert int(L["n_parts"].sum()) == truth["parts"] and int(L["n_rings"].sum()) == truth["rings"]
    assert int(L["num_features"].sum()) == truth["features"]
    # order-insensitive checksum of all assembled coordinates (closing vertices included): slices are 16-byte aligned, so the
    # padding between them (<= 3 ints) is masked through the layer table
    coords = res.buffer(abi.BUF_A_COORDS)
    o = L["out"][:, abi.BUF_A_COORDS].astype(np.int64)
    n2 = 2 * L["n_coords"].astype(np.int64)
    sx = sy = 0
    step = 1 << 16
    for l0 in range(0, len(L), step):
        lo, hi = int(o[l0]), int((o[l0:l0 + step] + n2[l0:l0 + step]).max())
        d = np.bincount(o[l0:l0 + step] - lo, minlength=hi - lo + 1) - np.bincount(o[l0:l0 + step] + n2[l0:l0 + step] - lo, minlength=hi - lo + 1)
        inside = np.cumsum(d[:-1]) > 0
        c = np.where(inside, coords[lo:hi], 0).astype(np.int64)
        # every slice starts at an even element (16-byte aligned), so x sits at even positions of the buffer
        assert lo % 2 == 0
        sx += int(c[0::2].sum())
        sy += int(c[1::2].sum())
    assert (sx, sy) == (truth["sum_x_closed"], truth["sum_y_closed"])
    res.free()


def test_config3_one_gibibyte_stream_bit_exact(covt, oracle, gen, decoder):
    if _mem_available_gb() < 24:
        pytest.skip("needs 24 GB of free host memory")
    abi = covt.abi
    enc, n = gen.varint_stream(1 << 30, seed=0xC0717)
    assert len(enc) == 1 << 30
    got, st, cons = decoder.decode_stream(enc, abi.OP_VARINT_ZZ_DELTA_XY, num_values=n, byte_length=len(enc))
    assert st == 0 and cons == len(enc) and len(got) == n
    want, wst, wcons = oracle.decode_stream(enc, abi.OP_VARINT_ZZ_DELTA_XY, byte_offset=0, byte_length=len(enc), num_values=n)
    assert wst == 0 and wcons == len(enc)
    assert np.array_equal(got, want)
    # size-independent property: the last vertex is the sum of all zigzag deltas, i.e. decode is linear in the deltas
    assert int(got[-2]) == int(np.int32(np.diff(want[0::2].astype(np.int64), prepend=0).sum() & 0xFFFFFFFF))
