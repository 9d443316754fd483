"""The GeoArrow / Arrow materialiser (cov-tiles_b200/arrow.py) on the CPU: it wraps result buffers of the layout the product
and the oracle share, so the oracle's results are enough to check it — nested geometry lists against the canonical form of the
assembled buffers, property columns against the List<Optional> view (CovtParser.decodePropertyColumn)."""
import numpy as np
import pytest

import canon
import util

pa = pytest.importorskip("pyarrow")


def test_layer_tables_wrap_the_result_buffers(oracle, fixtures):
    import covt_loader
    covt = covt_loader.load()
    abi = oracle.abi
    from importlib import import_module
    arrow = import_module(covt.__name__ + ".arrow")
    tiles = [b for n, b in fixtures if n in ("omt/5_16_21", "omt/14_8298_10748", "amazon/10_518_352", "omt/2_2_2")]
    assert len(tiles) >= 3
    blob, offs = util.concat_tiles(tiles)
    flags = abi.FLAG_ID_DVZZ_IS_RLE  # rings not closed: the canonical form strips closing vertices anyway
    res = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    props = oracle.decode_properties(blob, offs, abi.CONTAINER_GEN2B, flags)
    buffers = {b: res.buffer(b) for b in range(abi.NUM_BUFFERS - 1)}
    tables = arrow.layer_tables(blob, res.layers, buffers, props)
    ok_layers = [L for L in res.layers if L["status"] == 0]
    assert len(tables) == len(ok_layers) > 20
    n_prop_cols = n_values = 0
    for (tile, name, T), L in zip(tables, ok_layers):
        F = int(L["num_features"])
        assert tile == L["tile"] and name == util.layer_name(blob, L) and T.num_rows == F
        types, g, p, r, c = canon.layer_slices(L, buffers, abi)
        assert T.column("geometry_type").to_numpy().tolist() == types.tolist()
        if L["has_id"]:
            o = int(L["out"][abi.BUF_S_IDS])
            assert np.array_equal(T.column("id").to_numpy(), buffers[abi.BUF_S_IDS][o:o + F])
        # the nested lists, flattened level by level, are the assembled buffers
        geom = T.column("geometry").chunk(0)
        parts = geom.flatten()
        rings = parts.flatten()
        xy = rings.flatten()
        assert len(parts) == L["n_parts"] and len(rings) == L["n_rings"] and len(xy) == L["n_coords"]
        assert np.array_equal(np.asarray(geom.offsets), g) and np.array_equal(np.asarray(parts.offsets), p) and np.array_equal(np.asarray(rings.offsets), r)
        assert np.array_equal(xy.flatten().to_numpy(), c)
        if F:  # one feature spelled out: list of parts of rings of [x, y]
            f = int(F // 2)
            want = [[[list(map(int, c[2 * v:2 * v + 2])) for v in range(r[q], r[q + 1])] for q in range(p[k], p[k + 1])] for k in range(g[f], g[f + 1])]
            assert geom[f].as_py() == want
        # property columns of this layer: Arrow nulls and values = the List<Optional> view
        mine = [cc for cc in props.columns if cc["tile"] == L["tile"] and cc["layer"] == L["layer_index"] and cc["status"] == 0]
        assert T.num_columns == (1 + int(L["has_id"]) + 1) + len(mine)
        for k, cc in enumerate(mine):
            col = T.column(1 + int(L["has_id"]) + 1 + k)
            want = props.column_values(blob, cc)
            got = col.to_pylist()
            if cc["value_kind"] in (abi.PV_F32, abi.PV_F64):
                assert [None if v is None else np.float32(v) for v in got] == [None if v is None else np.float32(v) for v in want]
            else:
                assert got == want, T.column_names[1 + int(L["has_id"]) + 1 + k]
            n_prop_cols += 1
            n_values += sum(v is not None for v in got)
    assert n_prop_cols > 100 and n_values > 10000
