"""Shared helpers of the test-suite (test infrastructure)."""
import json
import os
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_fixture_tiles():
    """-> list of (name 'omt/5_16_21', bytes) from the committed bundle of the reference's gen-2b fixtures."""
    out = []
    with tarfile.open(os.path.join(GOLDEN, "fixtures_covt.tar.xz"), "r:xz") as tf:
        for m in tf.getmembers():
            out.append((m.name[:-5], tf.extractfile(m).read()))
    out.sort()
    return out


def concat_tiles(blobs):
    offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in blobs])
    blob = np.frombuffer(b"".join(bytes(b) for b in blobs), dtype=np.uint8) if blobs else np.zeros(0, np.uint8)
    return blob, offs


def mvt_digests():
    with open(os.path.join(GOLDEN, "mvt_geometry_digests.json")) as fh:
        return json.load(fh)


def fixture_flags(abi, name):
    """Quirk switches that reproduce the .mvt ground truth for a fixture tile."""
    f = abi.FLAG_ID_DVZZ_IS_RLE
    if name.startswith("omt/8_"):
        f |= abi.FLAG_MORTON_NO_SHIFT
    return f


# Layers whose committed .covt is known not to reproduce the partner MVT
KNOWN_MVT_MISMATCH = {"omt/11_1063_1368/landcover", "omt/6_34_41/water", "omt/8_134_171/park"}
# ICE vertex buffers labelled FAST_PFOR but varint-coded: decode error expected (COUNT_MISMATCH)
KNOWN_MISLABELLED = {"omt/4_8_10/water_name", "amazon/6_33_21/Graticule", "amazon/8_136_89/Colormap"}


def layer_name(blob, L):
    return bytes(blob[int(L["name_offset"]):int(L["name_offset"]) + int(L["name_length"])]).decode("utf-8", "replace")


def compare_results(abi, got, want, blob=None, check_assembled=True, same_container=True):
    """Bit-exact comparison of two batch results (product Result-like vs oracle BatchResult-like).
    `got`/`want` expose .layers (structured array), .buffer(which) -> ndarray, tile status arrays.
    Returns the number of layers compared; raises AssertionError with a useful message."""
    gl, wl = got.layers, want.layers
    assert len(gl) == len(wl), "layer count %d != %d" % (len(gl), len(wl))
    fields = ["tile", "layer_index", "extent", "num_features", "geom_column_type", "num_bits", "has_id", "cap_parts", "cap_rings"]
    if same_container:  # a re-wrapped tile has other metadata offsets and no property columns
        fields += ["num_columns", "name_length", "name_offset"]
    for f in fields:
        assert np.array_equal(gl[f], wl[f]), "layer field %s differs" % f
    assert np.array_equal(gl["out"], wl["out"]), "result layout (out offsets) differs"
    for f in (("byte_offset",) if same_container else ()) + ("byte_length", "num_values", "encoding", "op"):
        assert np.array_equal(gl["streams"][f], wl["streams"][f]), "stream field %s differs" % f
    ok_g = gl["status"] == 0
    ok_w = wl["status"] == 0
    assert np.array_equal(ok_g, ok_w), "layer status OK-ness differs: got %s want %s" % (gl["status"][ok_g != ok_w][:8], wl["status"][ok_g != ok_w][:8])
    sg = gl["streams"]["status"] == 0
    sw = wl["streams"]["status"] == 0
    assert np.array_equal(sg, sw), "stream status OK-ness differs"
    gb = [got.buffer(b) for b in range(abi.NUM_BUFFERS - 1)]
    wb = [want.buffer(b) for b in range(abi.NUM_BUFFERS - 1)]
    for b in range(abi.NUM_BUFFERS - 1):
        assert len(gb[b]) == len(wb[b]), "buffer %s length %d != %d" % (abi.BUF_NAMES[b], len(gb[b]), len(wb[b]))
    n = 0
    for i in range(len(gl)):
        G, W = gl[i], wl[i]
        for s in range(abi.NUM_SLOTS):
            if W["streams"][s]["encoding"] == abi.ENC_ABSENT or W["streams"][s]["status"] != 0:
                continue
            b = abi.SLOT_BUF[s]
            cnt = int(W["streams"][s]["num_values"])
            if s == abi.SLOT_VBUF and W["geom_column_type"] in (abi.CT_ICE, abi.CT_ICE_MORTON_CODE):
                cnt *= 2
            o = int(W["out"][b])
            if not np.array_equal(gb[b][o:o + cnt], wb[b][o:o + cnt]):
                bad = np.nonzero(gb[b][o:o + cnt] != wb[b][o:o + cnt])[0]
                raise AssertionError("layer %d (tile %d) stream %s (op %s, %d values) differs at %s: got %s want %s" % (
                    i, W["tile"], abi.SLOT_NAMES[s], abi.OP_NAMES[W["streams"][s]["op"]], cnt, bad[:5],
                    gb[b][o:o + cnt][bad[:5]], wb[b][o:o + cnt][bad[:5]]))
        if W["status"] != 0 or not check_assembled:
            continue
        for f in ("n_parts", "n_rings", "n_vertices", "n_coords"):
            assert G[f] == W[f], "layer %d %s %d != %d" % (i, f, G[f], W[f])
        F = int(W["streams"][abi.SLOT_TYPES]["num_values"])
        for b, cnt in ((abi.BUF_A_GEOM_OFFSETS, F + 1), (abi.BUF_A_PART_OFFSETS, int(W["n_parts"]) + 1),
                       (abi.BUF_A_RING_OFFSETS, int(W["n_rings"]) + 1), (abi.BUF_A_COORDS, 2 * int(W["n_coords"]))):
            o = int(W["out"][b])
            if not np.array_equal(gb[b][o:o + cnt], wb[b][o:o + cnt]):
                bad = np.nonzero(gb[b][o:o + cnt] != wb[b][o:o + cnt])[0]
                raise AssertionError("layer %d (tile %d) assembled %s differs at %s: got %s want %s" % (
                    i, W["tile"], abi.BUF_NAMES[b], bad[:5], gb[b][o:o + cnt][bad[:5]], wb[b][o:o + cnt][bad[:5]]))
        n += 1
    return n


# ---- gen-2b -> gen-3 metadata-only re-wrap (SURVEY §8c "gen-3 inputs") ---------------------------------
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def rewrap_gen3(abi, oracle, tile_bytes, optimized=False):
    """Re-wraps one gen-2b tile as gen-3 (CovtParser.decodeLayerMetadata grammar, CovtParser.java:574-652):
    same stream payload bytes and order, id + geometry columns only (property columns dropped).
    Returns (gen-3 bytes, n_fields list for the TileJSON side-car or None)."""
    blob = np.frombuffer(tile_bytes, dtype=np.uint8)
    rc, layers, ep = oracle.parse_tile(blob, abi.CONTAINER_GEN2B, flags=0)
    assert rc == 0 and ep == len(blob)
    out = bytearray()
    stream_type_of_slot = [abi.ST_DATA, abi.ST_GEOMETRY_TYPES, abi.ST_GEOMETRY_OFFSETS, abi.ST_PART_OFFSETS,
                           abi.ST_RING_OFFSETS, abi.ST_VERTEX_OFFSETS, abi.ST_VERTEX_BUFFER, abi.ST_INDEX_BUFFER]
    for li, L in enumerate(layers):
        name = bytes(blob[int(L["name_offset"]):int(L["name_offset"]) + int(L["name_length"])])
        out.append((1 << 1) | (1 if optimized else 0))
        out += _varint(li) if optimized else _varint(len(name)) + name
        out += _varint(int(L["extent"])) + _varint(int(L["num_features"])) + _varint(2 if L["has_id"] else 1)
        col = 0
        if L["has_id"]:
            s = L["streams"][abi.SLOT_ID]
            out += _varint(0) + bytes([(abi.DT_UINT_64 << 3) | abi.CT_PLAIN])
            out += bytes([(abi.ST_DATA << 4) | int(s["encoding"])]) + _varint(int(s["num_values"])) + _varint(int(s["byte_length"]))
            col += 1
        out += _varint(1) if (optimized or col == 0) else _varint(8) + b"geometry"
        out.append((abi.DT_GEOMETRY << 3) | int(L["geom_column_type"]))
        order = [abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VOFF, abi.SLOT_INDEX, abi.SLOT_VBUF]
        for slot in order:
            s = L["streams"][slot]
            if s["encoding"] == abi.ENC_ABSENT:
                continue
            out += bytes([(stream_type_of_slot[slot] << 4) | int(s["encoding"])]) + _varint(int(s["num_values"])) + _varint(int(s["byte_length"]))
        for slot in range(abi.NUM_SLOTS):  # payload order = slot order
            s = L["streams"][slot]
            if s["encoding"] == abi.ENC_ABSENT:
                continue
            out += bytes(blob[int(s["byte_offset"]):int(s["byte_offset"]) + int(s["byte_length"])])
    return bytes(out), ([0] * len(layers) if optimized else None)


# ---- the same comparison without a per-layer Python loop (batches of 10^6 layers) ---------------------------------
def compare_results_bulk(abi, got, want, chunk_layers=1 << 16, same_container=True):
    """Bit-exact comparison of every decoded stream and every assembled buffer of two batch results, vectorised: elements that
    belong to no valid layer slice (alignment padding between slices, slices of failed streams / layers) are masked out.
    Returns (layers compared, elements compared)."""
    gl, wl = got.layers, want.layers
    assert len(gl) == len(wl), "layer count %d != %d" % (len(gl), len(wl))
    # (same_container=False: the same tiles with re-encoded streams — offsets, lengths and encodings legitimately differ)
    for f in ("tile", "layer_index", "extent", "num_features", "geom_column_type", "num_bits", "has_id", "cap_parts", "cap_rings",
              "num_columns", "name_length", "status", "n_parts", "n_rings", "n_vertices", "n_coords") + (("name_offset",) if same_container else ()):
        assert np.array_equal(gl[f], wl[f]), "layer field %s differs" % f
    assert np.array_equal(gl["out"], wl["out"]), "result layout (out offsets) differs"
    for f in ("num_values", "status") + (("byte_offset", "byte_length", "encoding", "op") if same_container else ()):
        assert np.array_equal(gl["streams"][f], wl["streams"][f]), "stream field %s differs" % f
    layer_ok = wl["status"] == 0
    n_elems = 0
    # (buffer, start element, element count) of every valid slice, as arrays over layers
    slices = {b: [] for b in range(abi.NUM_BUFFERS - 1)}
    for s in range(abi.NUM_SLOTS):
        b = abi.SLOT_BUF[s]
        st = wl["streams"][:, s]
        ok = (st["encoding"] != abi.ENC_ABSENT) & (st["status"] == 0)
        cnt = st["num_values"].astype(np.int64)
        if s == abi.SLOT_VBUF:
            cnt = np.where(np.isin(wl["geom_column_type"], (abi.CT_ICE, abi.CT_ICE_MORTON_CODE)), 2 * cnt, cnt)
        slices[b].append((wl["out"][:, b].astype(np.int64), np.where(ok, cnt, 0)))
    F = wl["streams"][:, abi.SLOT_TYPES]["num_values"].astype(np.int64)
    for b, cnt in ((abi.BUF_A_GEOM_OFFSETS, F + 1), (abi.BUF_A_PART_OFFSETS, wl["n_parts"].astype(np.int64) + 1),
                   (abi.BUF_A_RING_OFFSETS, wl["n_rings"].astype(np.int64) + 1), (abi.BUF_A_COORDS, 2 * wl["n_coords"].astype(np.int64))):
        slices[b].append((wl["out"][:, b].astype(np.int64), np.where(layer_ok, cnt, 0)))
    for b in range(abi.NUM_BUFFERS - 1):
        g, w = got.buffer(b), want.buffer(b)
        assert len(g) == len(w), "buffer %s length %d != %d" % (abi.BUF_NAMES[b], len(g), len(w))
        for start, cnt in slices[b]:
            end = start + cnt
            assert (end <= len(w)).all(), "buffer %s: a slice ends outside the buffer" % abi.BUF_NAMES[b]
            for l0 in range(0, len(wl), chunk_layers):
                s_, e_ = start[l0:l0 + chunk_layers], end[l0:l0 + chunk_layers]
                have = e_ > s_
                if not have.any():
                    continue
                lo, hi = int(s_[have].min()), int(e_[have].max())
                d = np.bincount(s_[have] - lo, minlength=hi - lo + 1) - np.bincount(e_[have] - lo, minlength=hi - lo + 1)
                inside = np.cumsum(d[:-1]) > 0
                neq = (g[lo:hi] != w[lo:hi]) & inside
                if neq.any():
                    at = lo + int(np.nonzero(neq)[0][0])
                    li = l0 + int(np.nonzero((s_ <= at) & (e_ > at))[0][0])
                    raise AssertionError("buffer %s differs at element %d (layer %d, tile %d): got %s want %s" % (
                        abi.BUF_NAMES[b], at, li, wl["tile"][li], g[at:at + 4], w[at:at + 4]))
                n_elems += int(inside.sum())
        del g, w
    return int(layer_ok.sum()), n_elems
