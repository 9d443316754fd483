"""The tile writer over the GPU stream encoders (cov-tiles_b200/converter.py): tiles byte-identical to those of the CPU
restatement of the reference converter (tools/gen/covt_gen.c: covt_gen_append_layer — stream selection "encode both ways, keep the
shorter", ICE Morton dictionary, gen-2b and gen-3 metadata), and back through the GPU decoder to the layers they were written from."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

POINT, LINESTRING, POLYGON, MULTIPOINT, MULTILINESTRING, MULTIPOLYGON = range(6)


def random_layer(rng, n_features, name="layer", extent=4096, with_ids=True, step=60, id_kind="seq"):
    """A valid layer in the converter's input form: types + the three count streams + vertices (no closing vertices)."""
    types, geom, part, ring, xy = [], [], [], [], []
    cur = np.array([extent // 2, extent // 2])

    def walk(n):
        nonlocal cur
        for _ in range(n):
            cur = np.clip(cur + rng.integers(-step, step + 1, 2), 0, extent - 1)
            xy.extend(int(c) for c in cur)

    def polygon():
        nr = 1 if rng.random() < 0.8 else int(rng.integers(2, 4))
        part.append(nr)
        for _ in range(nr):
            nv = int(rng.integers(3, 12))
            ring.append(nv)
            walk(nv)

    for _ in range(n_features):
        t = int(rng.choice([POINT, LINESTRING, POLYGON, MULTILINESTRING, MULTIPOLYGON], p=[0.15, 0.5, 0.2, 0.08, 0.07]))
        types.append(t)
        if t == POINT:
            walk(1)
        elif t == LINESTRING:
            nv = int(rng.integers(2, 20))
            part.append(nv)
            walk(nv)
        elif t == POLYGON:
            polygon()
        elif t == MULTILINESTRING:
            np_ = int(rng.integers(1, 4))
            geom.append(np_)
            for _ in range(np_):
                nv = int(rng.integers(2, 9))
                part.append(nv)
                walk(nv)
        else:
            np_ = int(rng.integers(1, 3))
            geom.append(np_)
            for _ in range(np_):
                polygon()
    L = {"name": name, "extent": extent, "types": types, "geom": geom, "part": part, "ring": ring, "xy": xy}
    if with_ids:
        if id_kind == "seq":
            L["ids"] = np.arange(1, n_features + 1, dtype=np.int64) * 3
        elif id_kind == "big":
            L["ids"] = rng.integers(0, 1 << 50, n_features).astype(np.int64)
        else:
            L["ids"] = np.cumsum(rng.integers(1, 5000, n_features)).astype(np.int64)
    return L


OPTION_SETS = [3, 0, 3 | 4, 4, 3 | 8, 3 | 16, 3 | 32, 3 | 4 | 16]  # ALLOW_PFOR_TOPOLOGY 1, _VERTEX 2, ICE_MORTON 4, ID_DELTA_VARINT 8, FORCE_VARINT_VERTEX 16, FORCE_RLE_TOPOLOGY 32


@pytest.mark.parametrize("container", [0, 1])
def test_converted_tiles_equal_the_reference_converter_restatement(covt, gen, decoder, container):
    rng = np.random.default_rng(21 + container)
    conv = covt.CovtConverter(decoder)
    tiles, opts = [], []
    for k, options in enumerate(OPTION_SETS * 3):
        layers = [random_layer(rng, int(n), name="l%d" % i, extent=int(rng.choice([4096, 8192])), with_ids=bool((k + i) % 3),
                               id_kind=["seq", "big", "walk"][(k + i) % 3], step=int(rng.choice([3, 60, 900])))
                  for i, n in enumerate(rng.choice([0, 1, 3, 40, 300, 1500], size=int(rng.integers(1, 4))))]
        for L in layers:
            L["options"] = options
        tiles.append(layers)
        opts.append(options)
    big = random_layer(rng, 30000, name="big", step=40)  # FastPFOR pages / long RLE streams
    big["options"] = 3
    tiles.append([big])
    got = conv.convert_tiles(tiles, container)
    for layers, g in zip(tiles, got):
        want = bytes(gen.make_tile(layers, container, layers[0]["options"]))
        assert g == want, "tile with options %#x: %d bytes vs %d, first difference at %d" % (
            layers[0]["options"], len(g), len(want), next((i for i, (a, b) in enumerate(zip(g, want)) if a != b), min(len(g), len(want))))


def test_converted_tiles_decode_back(covt, decoder):
    """write -> decode on the GPU: types, ids, the decoded vertex streams and the assembled coordinates equal what went in."""
    abi = covt.abi
    rng = np.random.default_rng(23)
    conv = covt.CovtConverter(decoder)
    for container in (0, 1):
        for options in (3, 3 | 4):
            tiles = [[random_layer(rng, int(n), name="a%d" % n) for n in (5, 200)] for _ in range(6)]
            blobs = conv.convert_tiles(tiles, container, options)
            blob, offs = util.concat_tiles(blobs)
            res = decoder.decode_batch(blob, offs, container, abi.FLAG_DEFAULT & ~abi.FLAG_CLOSE_RINGS)
            st, first = res.tile_status()
            assert not st.any()
            layers = res.layers
            flat = [L for t in tiles for L in t]
            assert len(layers) == len(flat)
            types, ids, coords = res.buffer(abi.BUF_S_GEOMETRY_TYPES), res.buffer(abi.BUF_S_IDS), res.buffer(abi.BUF_A_COORDS)
            for R, L in zip(layers, flat):
                F = len(L["types"])
                assert R["status"] == 0 and R["num_features"] == F
                o = int(R["out"][abi.BUF_S_GEOMETRY_TYPES])
                assert np.array_equal(types[o:o + F], np.asarray(L["types"], np.uint8))
                o = int(R["out"][abi.BUF_S_IDS])
                assert np.array_equal(ids[o:o + F], L["ids"])
                o = int(R["out"][abi.BUF_A_COORDS])
                assert R["n_coords"] * 2 == len(L["xy"])  # rings not closed: the assembled coordinates are the input vertices
                assert np.array_equal(coords[o:o + len(L["xy"])], np.asarray(L["xy"], np.int32))
            res.free()
