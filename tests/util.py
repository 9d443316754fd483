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


def prop_column_key(blob, c):
    """MVT-style key of a decoded property column record (covt_prop_column): `<column>` or `<column>:<sub-key>`."""
    key = bytes(blob[int(c["name_offset"]):int(c["name_offset"]) + int(c["name_length"])]).decode("utf-8")
    if c["sub_length"]:
        sub = bytes(blob[int(c["sub_offset"]):int(c["sub_offset"]) + int(c["sub_length"])]).decode("utf-8")
        key = key if sub == key else key + ":" + sub
    return key


# ---- gen-2b -> gen-3 metadata-only re-wrap (SURVEY §8c "gen-3 inputs") ---------------------------------
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _gen3_property_column(abi, P, tile_arr, tile, col, F, gen):
    """One gen-2b property column as the HEAD converter lays it out (CovtConverter.java:1062-1195): (descriptor byte, listed stream
    metadata bytes, payload bytes) or None for column kinds HEAD cannot write (localized dictionaries, :1180-1182)."""
    dt2, ct = col["data_type"], col["column_type"]
    S = {s["name"]: s for s in col["streams"]}

    def raw(s):
        return bytes(tile[s["offset"]:s["offset"] + s["byte_length"]])

    def meta(stream_type, s, byte_length=None, encoding=None):
        return (bytes([(stream_type << 4) | (s["encoding"] if encoding is None else encoding)]) + _varint(s["num_values"]) +
                _varint(s["byte_length"] if byte_length is None else byte_length))
    if dt2 == P.DT2_BOOLEAN and ct == abi.CT_PLAIN:
        # HEAD: data = one bit per FEATURE, Byte-RLE, no present stream (CovtParser.java:280-290, CovtConverter.java:1062-1076)
        present = P._bitset(tile_arr, S["present"], F) if "present" in S else np.ones(F, bool)
        dense = P._bitset(tile_arr, S["data"], S["data"]["num_values"])
        bits = np.zeros(F, bool)
        bits[np.nonzero(present)[0][:len(dense)]] = dense[:int(present.sum())]
        payload = bytes(gen.encode_byte_rle(np.packbits(bits, bitorder="little")))
        return (abi.DT_BOOLEAN << 3) | abi.CT_PLAIN, meta(abi.ST_DATA, S["data"], len(payload), abi.ENC_BOOLEAN_RLE), payload
    if ct == abi.CT_LOCALIZED_DICTIONARY or "present" not in S or "data" not in S:
        return None
    head_dt = {P.DT2_STRING: abi.DT_STRING, P.DT2_FLOAT: abi.DT_FLOAT, P.DT2_DOUBLE: abi.DT_DOUBLE, P.DT2_INT_64: abi.DT_INT_64,
               P.DT2_UINT_64: abi.DT_UINT_64}.get(dt2)
    if head_dt is None:
        return None
    if dt2 == P.DT2_STRING:
        if ct != abi.CT_DICTIONARY:
            return None
        # payload order present, data, length, dictionary (CovtConverter.java:1152-1167); the present stream is NOT listed (:434-436)
        return ((head_dt << 3) | ct, meta(abi.ST_DATA, S["data"]) + meta(abi.ST_LENGTH, S["length"]) + meta(abi.ST_DICTIONARY, S["dictionary"]),
                raw(S["present"]) + raw(S["data"]) + raw(S["length"]) + raw(S["dictionary"]))
    if ct != abi.CT_PLAIN:
        return None
    return (head_dt << 3) | ct, meta(abi.ST_DATA, S["data"]), raw(S["present"]) + raw(S["data"])


def rewrap_gen3(abi, oracle, tile_bytes, optimized=False, props=False, gen=None, id_last=False):
    """Re-wraps one gen-2b tile as gen-3 (CovtParser.decodeLayerMetadata grammar, CovtParser.java:574-652): same stream payload
    bytes and order. props=False: id + geometry columns only. props=True (needs gen = tools.gen.gen): the property columns are
    kept the way the HEAD converter writes them — unlisted Byte-RLE present streams, BOOLEAN data as one bit per feature,
    payloads in column order after the geometry (localized dictionaries, which HEAD cannot write, are dropped).
    id_last: the id column FOLLOWS the geometry column (metadata and payload), the order CovtParser.java:64-85 also accepts.
    Returns (gen-3 bytes, n_fields list for the TileJSON side-car or None)."""
    blob = np.frombuffer(tile_bytes, dtype=np.uint8)
    rc, layers, ep = oracle.parse_tile(blob, abi.CONTAINER_GEN2B, flags=0)
    assert rc == 0 and ep == len(blob)
    pl = None
    if props:
        from oracle import properties as P
        pl = P.walk_gen2b(bytes(tile_bytes))
        tile_arr = np.frombuffer(bytes(tile_bytes) + bytes(64), dtype=np.uint8)
    out = bytearray()
    n_fields = []
    stream_type_of_slot = [abi.ST_DATA, abi.ST_GEOMETRY_TYPES, abi.ST_GEOMETRY_OFFSETS, abi.ST_PART_OFFSETS,
                           abi.ST_RING_OFFSETS, abi.ST_VERTEX_OFFSETS, abi.ST_VERTEX_BUFFER, abi.ST_INDEX_BUFFER]
    for li, L in enumerate(layers):
        name = bytes(blob[int(L["name_offset"]):int(L["name_offset"]) + int(L["name_length"])])
        pcols = []
        if props:
            for c in pl[li]["columns"]:
                if c["data_type"] == P.DT2_GEOMETRY or c["name"] == "id":
                    continue
                pc = _gen3_property_column(abi, P, tile_arr, bytes(tile_bytes), c, int(L["num_features"]), gen)
                if pc is not None:
                    pcols.append((c["name"].encode(),) + pc)
        n_fields.append(len(pcols))
        out.append((1 << 1) | (1 if optimized else 0))
        out += _varint(li) if optimized else _varint(len(name)) + name
        out += _varint(int(L["extent"])) + _varint(int(L["num_features"])) + _varint((2 if L["has_id"] else 1) + len(pcols))

        def id_column(first):
            s = L["streams"][abi.SLOT_ID]
            return ((_varint(0) if (optimized or first) else _varint(2) + b"id") + bytes([(abi.DT_UINT_64 << 3) | abi.CT_PLAIN]) +
                    bytes([(abi.ST_DATA << 4) | int(s["encoding"])]) + _varint(int(s["num_values"])) + _varint(int(s["byte_length"])))

        def geometry_column(first):
            m = (_varint(1) if (optimized or first) else _varint(8) + b"geometry") + bytes([(abi.DT_GEOMETRY << 3) | int(L["geom_column_type"])])
            for slot in [abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VOFF, abi.SLOT_INDEX, abi.SLOT_VBUF]:
                s = L["streams"][slot]
                if s["encoding"] != abi.ENC_ABSENT:
                    m += bytes([(stream_type_of_slot[slot] << 4) | int(s["encoding"])]) + _varint(int(s["num_values"])) + _varint(int(s["byte_length"]))
            return m

        def payload_of(slots):
            b = bytearray()
            for slot in slots:
                s = L["streams"][slot]
                if s["encoding"] != abi.ENC_ABSENT:
                    b += bytes(blob[int(s["byte_offset"]):int(s["byte_offset"]) + int(s["byte_length"])])
            return b
        geom_slots = list(range(abi.SLOT_TYPES, abi.NUM_SLOTS))  # payload order = slot order
        if L["has_id"] and not id_last:
            out += id_column(True) + geometry_column(False)
            payload = payload_of([abi.SLOT_ID]) + payload_of(geom_slots)
        elif L["has_id"]:
            out += geometry_column(True) + id_column(False)
            payload = payload_of(geom_slots) + payload_of([abi.SLOT_ID])
        else:
            out += geometry_column(True)
            payload = payload_of(geom_slots)
        for k, (cname, desc, smeta, spayload) in enumerate(pcols):
            out += (_varint(2 + k) if optimized else _varint(len(cname)) + cname) + bytes([desc]) + smeta
            payload += spayload
        out += payload
    return bytes(out), (n_fields if optimized else None)


def swap_id_and_geometry_gen2b(P, tile):
    """The same gen-2b tile with the id column moved behind the geometry column, metadata and payload (test input for the
    metadata-order placement rule)."""
    tile = bytes(tile)
    layers = P.walk_gen2b(tile)

    def col_meta(c):
        m = _varint(len(c["name"].encode())) + c["name"].encode() + bytes([c["data_type"], c["column_type"]]) + _varint(len(c["streams"]))
        for s in c["streams"]:
            m += _varint(len(s["name"].encode())) + s["name"].encode() + _varint(s["num_values"]) + _varint(s["byte_length"]) + bytes([s["encoding"]])
        return m

    def col_payload(c):
        ss = c["streams"]
        if c["name"] == "geometry":
            ss = sorted(ss, key=lambda s: s["offset"])
        return b"".join(tile[s["offset"]:s["offset"] + s["byte_length"]] for s in ss)
    out = bytearray(_varint(1) + _varint(len(layers)))
    for L in layers:
        cols = list(L["columns"])
        if len(cols) >= 2 and cols[0]["name"] == "id" and cols[1]["name"] == "geometry":
            cols[0], cols[1] = cols[1], cols[0]
        out += _varint(len(L["name"].encode())) + L["name"].encode() + _varint(L["extent"]) + _varint(L["num_features"]) + _varint(len(cols))
        for c in cols:
            out += col_meta(c)
        for c in cols:
            out += col_payload(c)
    return bytes(out)


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


# ---- property columns ------------------------------------------------------------------------------------------
class GpuProps:
    """Host copy of the property-column part of a product Result, with the oracle PropsResult's attribute names."""

    def __init__(self, abi, res):
        self.columns = res.prop_columns()
        self.dictionaries = res.prop_dictionaries()
        self.buffers = [res.prop_buffer(b) for b in range(abi.NUM_PROP_BUFFERS)]
        self.validity, self.i64, self.f32, self.f64, self.bools, self.dict_index, self.dict_offsets = self.buffers
        self.tile_status = res.tile_status()[0]


def compare_props(abi, blob, got, want):
    """Property columns of the product (GpuProps) vs the oracle (oracle.PropsResult): records, statuses, validity bitmaps, dense
    values and dictionary offsets, column by column (slices sit at the same offsets: both sides pad every slice to 16 bytes).
    Returns (columns compared, columns both sides accept)."""
    # (tile statuses are compared by compare_results: the product's include the first layer error, the oracle's property walk reports
    # the container walk alone)
    gc, wc = got.columns, want.columns
    assert len(gc) == len(wc), "column count %d != %d" % (len(gc), len(wc))
    gd, wd = got.dictionaries, want.dictionaries
    assert len(gd) == len(wd), "dictionary count %d != %d" % (len(gd), len(wd))
    for f in ("tile", "layer", "name_offset", "sub_offset", "name_length", "sub_length", "data_type", "column_type", "value_kind",
              "num_features", "dictionary"):
        assert np.array_equal(gc[f], wc[f]), "column field %s differs at %s" % (f, np.nonzero(gc[f] != wc[f])[0][:5])
    if not np.array_equal(gc["status"], wc["status"]):
        bad = np.nonzero(gc["status"] != wc["status"])[0]
        raise AssertionError("column status differs at %s: got %s want %s (kinds %s)" % (bad[:8], gc["status"][bad[:8]], wc["status"][bad[:8]], wc["value_kind"][bad[:8]]))
    for f in ("tile", "layer"):
        assert np.array_equal(gd[f], wd[f]), "dictionary field %s differs" % f
    if not np.array_equal(gd["status"], wd["status"]):
        bad = np.nonzero(gd["status"] != wd["status"])[0]
        raise AssertionError("dictionary status differs at %s: got %s want %s" % (bad[:8], gd["status"][bad[:8]], wd["status"][bad[:8]]))
    ok_d = wd["status"] == 0
    for f in ("n_entries", "offsets_offset", "bytes_offset", "n_bytes"):
        assert np.array_equal(gd[f][ok_d], wd[f][ok_d]), "dictionary field %s differs" % f
    for i in np.nonzero(ok_d)[0]:
        o, n = int(wd["offsets_offset"][i]), int(wd["n_entries"][i]) + 1
        assert np.array_equal(got.dict_offsets[o:o + n], want.dict_offsets[o:o + n]), "offsets of dictionary %d differ" % i
    ok = wc["status"] == 0
    for f in ("num_values", "validity_offset", "values_offset", "data_num_values"):
        assert np.array_equal(gc[f][ok], wc[f][ok]), "column field %s differs at %s" % (f, np.nonzero(gc[f][ok] != wc[f][ok])[0][:5])
    kind_buf = {abi.PV_I64: abi.PBUF_I64, abi.PV_F32: abi.PBUF_F32, abi.PV_F64: abi.PBUF_F64, abi.PV_BOOL: abi.PBUF_BOOL,
                abi.PV_DICT_INDEX: abi.PBUF_DICT_INDEX}
    for i in np.nonzero(ok)[0]:
        c = wc[i]
        F, n = int(c["num_features"]), int(c["num_values"])
        vo = int(c["validity_offset"])
        gv, wv = got.validity[vo:vo + (F + 7) // 8].copy(), want.validity[vo:vo + (F + 7) // 8].copy()
        if F & 7 and len(gv):  # bits behind the last feature are don't-care
            gv[-1] &= (1 << (F & 7)) - 1
            wv[-1] &= (1 << (F & 7)) - 1
        assert np.array_equal(gv, wv), "validity of column %d differs" % i
        b = kind_buf[int(c["value_kind"])]
        o = int(c["values_offset"])
        if b == abi.PBUF_BOOL:  # one slot (bit) per feature, zero where the feature has no value
            ga = np.unpackbits(got.buffers[b][o:o + (F + 7) // 8], bitorder="little")[:F]
            wa = np.unpackbits(want.buffers[b][o:o + (F + 7) // 8], bitorder="little")[:F]
        else:
            ga, wa = got.buffers[b][o:o + F], want.buffers[b][o:o + F]
        if b in (abi.PBUF_F32, abi.PBUF_F64):  # bit patterns (NaNs included)
            ga, wa = ga.view(np.uint32 if b == abi.PBUF_F32 else np.uint64), wa.view(np.uint32 if b == abi.PBUF_F32 else np.uint64)
        assert np.array_equal(ga, wa), "values of column %d (kind %d, tile %d) differ" % (i, c["value_kind"], c["tile"])
    return len(wc), int(ok.sum())
