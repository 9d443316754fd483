"""CPU: pins the property-column oracle (oracle/properties.py, SURVEY §8 f1 — the next row after the geometry + id path) against
the reference's own data: the property values of the partner .mvt/.pbf tiles of the 102 OMT + Amazon fixtures, committed as
digests by tests/golden/make_golden.py (the JVM is absent, the Java decoder cannot run here).

Every stream of every property column must consume exactly its declared byteLength, the gen-2b walk must end at EOF, and
  * every COVT property column whose key exists in the MVT layer must equal the MVT value column feature by feature;
  * a COVT column whose key the MVT layer does not have must be entirely null (the converter's localized-name splitting
    creates such empty sub-columns, e.g. `name:xx` for Amazon's `_name_xx` keys).
The 27 Bing tiles have no partner: their FLOAT and BOOLEAN columns are only checked for clean, byte-exact consumption."""
import json
import os

import numpy as np

import canon
import util

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "mvt_property_digests.json")) as fh:
        return json.load(fh)


def test_property_columns_equal_partner_mvt(oracle, fixtures):
    from oracle import properties as P
    gold = _golden()
    matched = empty = values = 0
    kinds = set()
    for name, data in fixtures:
        src, tile = name.split("/")
        if src == "bing":
            continue
        for layer_name, props in P.decode_property_columns(data):
            for key, col in props.items():
                # golden keys use the `_` spelling of localized keys (`name:de` and `name_de` are one sub-column in COVT)
                g = gold.get("%s/%s/%s/%s" % (src, tile, layer_name, key.replace(":", "_")))
                if g is None:
                    assert all(v is None for v in col), "%s/%s: column %s has values but the MVT layer has no such key" % (name, layer_name, key)
                    empty += 1
                    continue
                assert sum(v is not None for v in col) == g["present"], "%s/%s/%s: present count" % (name, layer_name, key)
                assert canon.property_digest(col) == g["digest"], "%s/%s/%s: values differ from the MVT" % (name, layer_name, key)
                matched += 1
                values += g["present"]
                kinds.update(type(v).__name__ for v in col if v is not None)
    assert matched >= 9000 and values >= 2_500_000, (matched, values)
    assert {"str", "int", "bool"} <= kinds
    # the converter drops a few MVT keys (SURVEY §4.4-style census, facts about the fixture bytes): they must stay few
    assert empty <= 3600


def test_bing_property_columns_decode_cleanly(oracle, fixtures):
    from oracle import properties as P
    n_cols = 0
    kinds = set()
    for name, data in fixtures:
        if not name.startswith("bing/"):
            continue
        for layer_name, props in P.decode_property_columns(data):
            for key, col in props.items():
                n_cols += 1
                kinds.update(type(v).__name__ for v in col if v is not None)
    assert n_cols >= 800 and {"str", "int", "float", "bool"} <= kinds


def test_c_property_oracle_equals_the_pinned_statement(oracle, fixtures):
    """covt_oracle_decode_properties (C, columnar: validity bitmap + dense values + dictionary offsets — the layout meant for
    the GPU path) over all 129 fixture tiles in one batch == oracle/properties.py, which the tests above pin on the MVT."""
    from oracle import properties as P
    blob, offs = util.concat_tiles([b for _, b in fixtures])
    res = oracle.decode_properties(blob, offs)
    assert not res.tile_status.any()
    by_tile = {}
    for c in res.columns:
        by_tile.setdefault(int(c["tile"]), []).append(c)
    n_cols = 0
    kinds = set()
    for t, (name, data) in enumerate(fixtures):
        want = P.decode_property_columns(data)
        layers = P.walk_gen2b(bytes(data))
        got = {}
        for c in by_tile.get(t, []):
            assert c["status"] == 0, (name, c["layer"], c["status"])
            key = util.prop_column_key(blob, c)
            got.setdefault(int(c["layer"]), {})[key] = res.column_values(blob, c)
            kinds.add(int(c["value_kind"]))
        for li, (layer_name, props) in enumerate(want):
            assert layers[li]["name"] == layer_name
            g = got.get(li, {})
            assert set(g) == set(props), (name, layer_name, sorted(set(g) ^ set(props))[:5])
            for key, col in props.items():
                a = g[key]
                assert len(a) == len(col) and all((x is None and y is None) or (x is not None and y is not None and (x == y or abs(x - y) < 1e-6 * max(1.0, abs(y))))
                                                  if not isinstance(y, (str, bool)) else x == y for x, y in zip(a, col)), (name, layer_name, key)
                n_cols += 1
    assert n_cols >= 13000 and {oracle.PV_I64, oracle.PV_F32, oracle.PV_BOOL, oracle.PV_DICT_INDEX} <= kinds
    assert not res.dictionaries["status"].any()


def test_gen3_property_columns_equal_gen2b(oracle, gen, fixtures):
    """The HEAD container (CovtParser.decodeLayerMetadata :574-652 + decodePropertyColumn :276-390: unlisted Byte-RLE present streams,
    BOOLEAN data as one bit per feature): gen-3 re-wraps of fixture tiles, optimised metadata and not, decode to the values of the
    gen-2b tiles they were made from (localized dictionaries, which HEAD cannot write, are dropped by the re-wrap)."""
    import covt_loader
    abi = covt_loader.load().abi
    names = [n for n, _ in fixtures if n.startswith(("omt/5_", "omt/10_", "amazon/6_", "bing/"))][:16]
    fx = dict(fixtures)
    blob2, offs2 = util.concat_tiles([fx[n] for n in names])
    w2 = oracle.decode_properties(blob2, offs2, abi.CONTAINER_GEN2B)
    keep = [c for c in w2.columns if c["column_type"] != abi.CT_LOCALIZED_DICTIONARY]
    for optimized in (False, True):
        tiles, n_fields = [], None
        for n in names:
            t, nf = util.rewrap_gen3(abi, oracle, fx[n], optimized=optimized, props=True, gen=gen)
            if optimized:
                n_fields = nf if n_fields is None else [max(a, b) for a, b in zip(n_fields + [0] * (len(nf) - len(n_fields)), nf + [0] * (len(n_fields) - len(nf)))]
            tiles.append(t)
        blob, offs = util.concat_tiles(tiles)
        w = oracle.decode_properties(blob, offs, abi.CONTAINER_GEN3, n_fields=n_fields)
        assert not w.tile_status.any() and not w.columns["status"].any() and len(w.columns) == len(keep) > 400
        for a, c in zip(w.columns, keep):
            va, vc = w.column_values(blob, a), w2.column_values(blob2, c)
            if a["value_kind"] == abi.PV_BOOL:  # HEAD booleans have no nulls: the re-wrap writes an absent value as false
                vc = [bool(v) for v in vc]
            assert va == vc, (util.prop_column_key(blob2, c), va[:6], vc[:6])


def test_property_columns_exist_for_complete_layers_only(oracle, fixtures):
    """A tile cut in the middle of a layer keeps the property columns of the layers before it, nothing of the broken one."""
    name, data = next((n, b) for n, b in fixtures if n == "omt/5_16_21")
    full = oracle.decode_properties(*util.concat_tiles([data]))
    rc, layers, _ = oracle.parse_tile(np.frombuffer(data, np.uint8), 0, flags=0)
    cut = int(layers[3]["streams"][1]["byte_offset"]) + 5  # inside the payload of layer 3
    part = oracle.decode_properties(*util.concat_tiles([data[:cut]]))
    assert part.tile_status[0] != 0
    n_before = int((full.columns["layer"] < 3).sum())
    assert len(part.columns) == n_before > 0 and set(int(x) for x in part.columns["layer"]) == {0, 1, 2}
    assert np.array_equal(part.columns["status"], full.columns["status"][:n_before])
