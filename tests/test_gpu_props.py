"""GPU: property columns on the batch path (COVT_FLAG_DECODE_PROPERTIES) — CovtParser.decodePropertyColumn
(J/decoder/CovtParser.java:276-390) as validity bitmaps + dense values + dictionary offsets. The CUDA path is compared with the C
oracle column by column (bit-exact: statuses, bitmaps, values, offsets), with the property values of the partner MVT tiles
(tests/golden/mvt_property_digests.json, generated from the reference's fixtures by tests/golden/make_golden.py), on the gen-2b
fixtures, on their gen-3 (HEAD container) re-wraps, and on mutated tiles."""
import json
import os

import numpy as np
import pytest

import canon
import util
from test_gpu_batch import _mutants

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _decode_props_both(covt, oracle, decoder, blob, offs, container, flags, n_fields=None):
    abi = covt.abi
    res = decoder.decode_batch(blob, offs, container, flags | abi.FLAG_DECODE_PROPERTIES, n_fields=n_fields)
    got = util.GpuProps(abi, res)
    want = oracle.decode_properties(blob, offs, container, flags, n_fields=n_fields)
    return res, got, want


def test_property_columns_of_all_fixtures_vs_oracle(covt, oracle, decoder, fixtures):
    """All 129 gen-2b fixture tiles in one batch: 13 000+ property columns (INT_64 RLE / varint, FLOAT, BOOLEAN with and without
    present stream, dictionary strings, localized dictionaries) equal to the oracle's; the geometry path of the same call is
    unchanged by the flag."""
    abi = covt.abi
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    blob, offs = util.concat_tiles([b for _, b in fixtures])
    res, got, want = _decode_props_both(covt, oracle, decoder, blob, offs, abi.CONTAINER_GEN2B, flags)
    n, n_ok = util.compare_props(abi, blob, got, want)
    assert n >= 13000 and n_ok == n and not got.dictionaries["status"].any()
    kinds = set(int(k) for k in got.columns["value_kind"])
    assert {abi.PV_I64, abi.PV_F32, abi.PV_BOOL, abi.PV_DICT_INDEX} <= kinds
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    util.compare_results(abi, res, ref)
    # the metric's unit: compressed bytes of every decoded stream, geometry + properties
    assert res.timing()["payload_bytes"] == ref.payload_bytes + want.payload_bytes and want.payload_bytes > 2_000_000
    res.free()
    # without the flag: no property columns, no property buffers
    res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    assert len(res.prop_columns()) == 0 and res.prop_device_buffer(abi.PBUF_I64)[1] == 0
    res.free()


def test_property_values_equal_partner_mvt(covt, decoder, fixtures):
    """The decoded columns against the reference's own data: the property values of the partner .mvt/.pbf tiles of the 102 OMT +
    Amazon fixtures (committed digests) — the GPU result pinned without the oracle in between."""
    abi = covt.abi
    with open(os.path.join(HERE, "golden", "mvt_property_digests.json")) as fh:
        gold = json.load(fh)
    named = [(n, b) for n, b in fixtures if not n.startswith("bing/")]
    blob, offs = util.concat_tiles([b for _, b in named])
    res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE | abi.FLAG_DECODE_PROPERTIES)
    got = util.GpuProps(abi, res)
    layers = res.layers
    first = res.tile_status()[1]
    kind_buf = {abi.PV_I64: got.i64, abi.PV_F32: got.f32, abi.PV_F64: got.f64, abi.PV_BOOL: got.bools, abi.PV_DICT_INDEX: got.dict_index}
    matched = values = 0
    for c in got.columns:
        assert c["status"] == 0
        name = named[int(c["tile"])][0]
        src, tile = name.split("/")
        L = layers[int(first[int(c["tile"])]) + int(c["layer"])]
        key = util.prop_column_key(blob, c)
        g = gold.get("%s/%s/%s/%s" % (src, tile, util.layer_name(blob, L), key.replace(":", "_")))
        col = abi.prop_column_values(blob, c, got.validity, kind_buf[int(c["value_kind"])], got.dict_offsets, got.dictionaries)
        if g is None:
            assert all(v is None for v in col), (name, key)
            continue
        assert sum(v is not None for v in col) == g["present"], (name, key)
        assert canon.property_digest(col) == g["digest"], "%s/%s: values differ from the MVT" % (name, key)
        matched += 1
        values += g["present"]
    assert matched >= 9000 and values >= 2_500_000, (matched, values)
    res.free()


@pytest.mark.parametrize("optimized", [False, True])
def test_property_columns_gen3(covt, oracle, gen, decoder, fixtures, optimized):
    """The HEAD container (CovtParser.decodeLayerMetadata + decodePropertyColumn): unlisted Byte-RLE present streams, BOOLEAN data as
    one bit per feature, dictionary strings; optimised metadata names its columns through the TileJSON side-car."""
    abi = covt.abi
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    tiles, nf_all = [], None
    names = [n for n, _ in fixtures if n.startswith(("omt/5_", "omt/10_", "omt/14_", "amazon/6_", "bing/"))][:24]
    fx = dict(fixtures)
    n_fields = None
    for n in names:
        t, nf = util.rewrap_gen3(abi, oracle, fx[n], optimized=optimized, props=True, gen=gen)
        if optimized:
            # one TileJSON for the batch: tile-local layer ids share it, so it must hold the largest field count per layer id
            n_fields = nf if n_fields is None else [max(a, b) for a, b in zip(n_fields + [0] * (len(nf) - len(n_fields)), nf + [0] * (len(n_fields) - len(nf)))]
        tiles.append(t)
    blob, offs = util.concat_tiles(tiles)
    res, got, want = _decode_props_both(covt, oracle, decoder, blob, offs, abi.CONTAINER_GEN3, flags, n_fields=n_fields)
    assert not want.tile_status.any()
    n, n_ok = util.compare_props(abi, blob, got, want)
    assert n >= 700 and n_ok == n
    assert {abi.PV_I64, abi.PV_F32, abi.PV_BOOL, abi.PV_DICT_INDEX} <= set(int(k) for k in got.columns["value_kind"])
    # the same columns as the gen-2b tiles hold (localized dictionaries, which HEAD cannot write, are dropped by the re-wrap)
    blob2, offs2 = util.concat_tiles([fx[n] for n in names])
    w2 = oracle.decode_properties(blob2, offs2, abi.CONTAINER_GEN2B, flags)
    keep = [c for c in w2.columns if c["column_type"] != abi.CT_LOCALIZED_DICTIONARY]
    assert len(keep) == n
    for a, c in zip(got.columns, keep):
        va = abi.prop_column_values(blob, a, got.validity, got.buffers[int(a["value_kind"])], got.dict_offsets, got.dictionaries)  # PV_x == PBUF_x for x = 1..5
        vc = w2.column_values(blob2, c)
        if a["value_kind"] == abi.PV_BOOL:  # HEAD booleans have no nulls: the re-wrap writes an absent value as false
            vc = [bool(v) for v in vc]
        assert va == vc, (util.prop_column_key(blob2, c), va[:8], vc[:8])
    res.free()


@pytest.mark.parametrize("corpus", ["gen2b", "gen3"])
def test_property_mutation_fuzz_against_oracle(covt, oracle, gen, decoder, fixtures, corpus):
    """Mutants of fixture tiles with their property columns (byte flips in metadata and payload, truncations, junk) in one batch:
    the call survives, the column list, every column status and every accepted column's bitmap / values / dictionary offsets are
    the oracle's."""
    abi = covt.abi
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    clean = [(n, b) for n, b in fixtures if n.startswith(("omt/", "bing/")) and not n.startswith("omt/8_")
             and not any(k.startswith(n + "/") for k in util.KNOWN_MISLABELLED)]
    base = [b for _, b in sorted(clean, key=lambda t: len(t[1]))[:8]]
    container = abi.CONTAINER_GEN2B
    if corpus == "gen3":
        base = [util.rewrap_gen3(abi, oracle, b, props=True, gen=gen)[0] for b in base]
        container = abi.CONTAINER_GEN3
    tiles, good = _mutants(np.random.default_rng(77), base, 1600, 40)
    blob, offs = util.concat_tiles(tiles)
    res, got, want = _decode_props_both(covt, oracle, decoder, blob, offs, container, flags)
    assert not got.tile_status[good].any()
    n, n_ok = util.compare_props(abi, blob, got, want)
    assert n > 2000 and n_ok > n // 2 and n_ok < n  # accepted and rejected columns are both exercised
    ref = oracle.decode_batch(blob, offs, container, flags)
    util.compare_results(abi, res, ref)
    res.free()


def test_property_columns_through_the_pipelined_host_path(covt, oracle, fixtures, monkeypatch):
    """covt_decode_batch uploads and decodes in segments; the property pass runs once over the layer table of the whole batch.
    The OMT fixture tiles x3 in many small segments (or, if the first segment was not representative, after the capacity retry)."""
    abi = covt.abi
    monkeypatch.setenv("COVT_SEG_BYTES", str(1 << 20))
    monkeypatch.setenv("COVT_MAX_SEGMENTS", "64")
    monkeypatch.setenv("COVT_SEG_MIN_TILES", "16")
    dec = covt.Decoder(0)
    try:
        flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
        tiles = [b for n, b in fixtures if n.startswith("omt/")] * 3
        blob, offs = util.concat_tiles(tiles)
        res, got, want = _decode_props_both(covt, oracle, dec, blob, offs, abi.CONTAINER_GEN2B, flags)
        t = res.timing()
        assert t["segments"] >= 4 or t["capacity_retries"] == 1
        n, n_ok = util.compare_props(abi, blob, got, want)
        assert n > 20000 and n_ok == n
        util.compare_results(abi, res, oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags))
        res.free()
    finally:
        dec.close()


def test_property_columns_two_gpus_one_call(covt, oracle, fixtures):
    """The library's batch scheduler passes the flag through: every part of the multi-GPU result carries the property columns of
    its tile range (tile indices relative to the part)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    abi = covt.abi
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    tiles = [b for n, b in fixtures if n.startswith("omt/")]
    blob, offs = util.concat_tiles(tiles)
    md = covt.MultiDecoder([0, 1])
    try:
        mres = md.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags | abi.FLAG_DECODE_PROPERTIES)
        total = 0
        for p in mres.parts:
            t0, n = p["first_tile"], p["n_tiles"]
            sub_blob = blob[int(offs[t0]):int(offs[t0 + n])]
            sub_offs = offs[t0:t0 + n + 1] - offs[t0]
            want = oracle.decode_properties(sub_blob, sub_offs, abi.CONTAINER_GEN2B, flags)
            k, k_ok = util.compare_props(abi, sub_blob, util.GpuProps(abi, p["result"]), want)
            assert k_ok == k
            total += k
        assert total > 9000
        mres.free()
    finally:
        md.close()


def test_result_to_arrow_equals_the_oracle_tables(covt, oracle, decoder, fixtures):
    """Result.to_arrow: the decoded layers as pyarrow Tables (GeoArrow-nested geometry, ids, property columns) wrapped around host copies
    of the GPU buffers = the same tables built from the oracle's results (tests/test_arrow.py checks those against the canonical forms)."""
    pa = pytest.importorskip("pyarrow")
    from importlib import import_module
    arrow = import_module(covt.__name__ + ".arrow")
    abi = covt.abi
    tiles = [b for n, b in fixtures if n.startswith(("omt/5_", "omt/10_", "amazon/"))]
    blob, offs = util.concat_tiles(tiles)
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags | abi.FLAG_DECODE_PROPERTIES)
    got = res.to_arrow(blob)
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    want = arrow.layer_tables(blob, ref.layers, {b: ref.buffer(b) for b in range(abi.NUM_BUFFERS - 1)},
                              oracle.decode_properties(blob, offs, abi.CONTAINER_GEN2B, flags))
    assert len(got) == len(want) > 100
    n_cols = 0
    for (t1, n1, T1), (t2, n2, T2) in zip(got, want):
        assert (t1, n1) == (t2, n2) and T1.column_names == T2.column_names
        assert T1.equals(T2), (t1, n1)
        n_cols += T1.num_columns
    assert n_cols > 1500
    res.free()
