"""CPU oracle for property columns of gen-2b tiles — TEST INFRASTRUCTURE ONLY, like the rest of oracle/.

Restates CovtParser.decodePropertyColumn (J/decoder/CovtParser.java:276-390) on top of the C stream codecs of covt_oracle.c,
for the container the committed fixtures use (gen-2b: every stream, the present stream included, is listed with name,
numValues, byteLength and StreamEncoding ordinal). Pinned by tests/test_oracle_properties.py against the property values of the
partner .mvt tiles. Not on the GPU path yet: the product decodes geometry + id columns only. The same decode with a
columnar result (validity bitmap + dense values + dictionary offsets: the layout meant for the GPU path) is
covt_oracle_decode_properties in covt_oracle.c; tests/test_oracle_properties.py holds the two against each other.

Layout established on the 129 fixtures (no gen-2b reader or writer exists in the reference at HEAD):
  * property payloads follow the geometry payload of their layer IN COLUMN-METADATA ORDER, stream after stream in listed order;
  * `present`  BOOLEAN_RLE: Byte-RLE of ceil(numValues / 8) bitset bytes, bit i = byte[i >> 3] >> (i & 7) (java.util.BitSet);
  * INT_64 / UINT_64 `data`: RLE (signed / unsigned), VARINT_ZIG_ZAG or VARINT_DELTA_ZIG_ZAG over the PRESENT values only;
  * BOOLEAN `data`: BOOLEAN_RLE bitset; HEAD reads one bit per feature and no present stream (CovtParser.java:280-290), the
    fixtures list a present stream for most boolean columns and `data` then holds one bit per PRESENT feature;
  * FLOAT / DOUBLE `data`: little-endian IEEE values of the present features (DecodingUtils.decodeFloatsLE :446-453);
  * STRING, ColumnType.DICTIONARY: `data` = RLE dictionary indices of the present features, `length` = RLE byte lengths of the
    dictionary entries, `dictionary` = the UTF-8 bytes back to back;
  * STRING, ColumnType.LOCALIZED_DICTIONARY: pairs (`present_<s>`, `<s>`) per sub-key, then ONE shared `length` + `dictionary`;
    sub-key `<column>` is the plain key, any other `<s>` is the MVT key `<column>:<s>` or `<column>_<s>`.
"""
import numpy as np

from . import oracle as O

abi = O.abi  # enums only (cov-tiles_b200/abi.py: constants + ctypes structs, no GPU code)

# gen-2 data type byte of the column header
DT2_STRING, DT2_FLOAT, DT2_DOUBLE, DT2_INT_64, DT2_UINT_64, DT2_BOOLEAN, DT2_GEOMETRY = range(7)


def _varint(b, p):
    v = s = 0
    for _ in range(4):  # DecodingUtils.decodeVarint: at most 4 bytes
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            break
    return v, p


def _string(b, p):
    n, p = _varint(b, p)
    return bytes(b[p:p + n]).decode("utf-8"), p + n


def walk_gen2b(tile):
    """-> [layer dict(name, extent, num_features, columns=[dict(name, data_type, column_type, streams=[dict(name, num_values,
    byte_length, encoding, offset)])])] with absolute payload offsets (payload order = metadata order)."""
    b = memoryview(tile)
    p = 0
    _, p = _varint(b, p)
    n_layers, p = _varint(b, p)
    layers = []
    for _ in range(n_layers):
        name, p = _string(b, p)
        extent, p = _varint(b, p)
        num_features, p = _varint(b, p)
        n_cols, p = _varint(b, p)
        cols = []
        for _ in range(n_cols):
            cname, p = _string(b, p)
            dt, ct = b[p], b[p + 1]
            p += 2
            n_streams, p = _varint(b, p)
            streams = []
            for _ in range(n_streams):
                sname, p = _string(b, p)
                nv, p = _varint(b, p)
                bl, p = _varint(b, p)
                streams.append({"name": sname, "num_values": nv, "byte_length": bl, "encoding": b[p]})
                p += 1
            cols.append({"name": cname, "data_type": dt, "column_type": ct, "streams": streams})
        # payload: [id] | geometry in the fixed order types, geometry_offsets, part_offsets, ring_offsets, vertex_offsets,
        # vertex_buffer | property columns in metadata order
        geom_order = ["geometry_types", "geometry_offsets", "part_offsets", "ring_offsets", "vertex_offsets", "vertex_buffer", "index_buffer"]
        for c in cols:
            ss = c["streams"]
            if c["name"] == "geometry":
                ss = sorted(ss, key=lambda s: geom_order.index(s["name"]))
            for s in ss:
                s["offset"] = p
                p += s["byte_length"]
        layers.append({"name": name, "extent": extent, "num_features": num_features, "columns": cols})
    if p != len(tile):
        raise ValueError("gen-2b walk ended at %d of %d" % (p, len(tile)))
    return layers


def _decode(tile_arr, s, op, num_values):
    vals, st, cons = O.decode_stream(tile_arr, op, byte_offset=s["offset"], byte_length=s["byte_length"], num_values=num_values)
    if st != 0:
        raise ValueError("stream %s: status %d" % (s["name"], st))
    if cons != s["byte_length"]:
        raise ValueError("stream %s: consumed %d of %d bytes" % (s["name"], cons, s["byte_length"]))
    return vals


def _bitset(tile_arr, s, n_bits):
    by = _decode(tile_arr, s, abi.OP_BYTE_RLE, (n_bits + 7) // 8)
    return np.unpackbits(by, bitorder="little")[:n_bits].astype(bool)


def _int_data(tile_arr, s, n, signed):
    enc = s["encoding"]
    if enc == abi.ENC_RLE:
        return _decode(tile_arr, s, abi.OP_RLE_S64 if signed else abi.OP_RLE_U64, n).astype(np.int64)
    if enc == abi.ENC_VARINT_ZIG_ZAG:
        return _decode(tile_arr, s, abi.OP_VARINT_ZZ, n).astype(np.int64)  # Java: int varints widened (CovtParser.java:303-306)
    if enc == abi.ENC_VARINT_DELTA_ZIG_ZAG:
        return _decode(tile_arr, s, abi.OP_VARINT_ZZ_DELTA, n).astype(np.int64)
    if enc == abi.ENC_VARINT:
        return _decode(tile_arr, s, abi.OP_VARINT_U32, n).astype(np.int64)
    raise ValueError("INT_64 data stream: unsupported encoding %d" % enc)  # CovtParser.java:313-315


def _dictionary(tile_arr, tile, length_s, dict_s):
    lens = _decode(tile_arr, length_s, abi.OP_RLE_U64, length_s["num_values"]).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    if offs[-1] != dict_s["byte_length"]:
        raise ValueError("dictionary: lengths sum to %d, stream has %d bytes" % (offs[-1], dict_s["byte_length"]))
    raw = bytes(tile[dict_s["offset"]:dict_s["offset"] + dict_s["byte_length"]])
    return [raw[offs[i]:offs[i + 1]].decode("utf-8") for i in range(len(lens))]


def _expand(present, dense):
    """List<Optional> of CovtParser.decodePropertyColumn: value i of the present features, None elsewhere."""
    out = [None] * len(present)
    it = iter(dense)
    for i in np.nonzero(present)[0]:
        out[i] = next(it)
    return out


def decode_property_columns(tile):
    """-> [(layer name, {property key: [value or None per feature]})] for one gen-2b tile."""
    tile = bytes(tile)
    arr = np.frombuffer(tile + bytes(64), dtype=np.uint8)
    result = []
    for L in walk_gen2b(tile):
        F = L["num_features"]
        props = {}
        for c in L["columns"]:
            if c["data_type"] == DT2_GEOMETRY or (c["name"] == "id" and c is L["columns"][0]):
                continue  # the id and geometry columns are the product's path (covt_oracle.c)
            S = {s["name"]: s for s in c["streams"]}
            dt, ct = c["data_type"], c["column_type"]
            if dt == DT2_BOOLEAN:
                present = _bitset(arr, S["present"], F) if "present" in S else np.ones(F, bool)
                data = _bitset(arr, S["data"], S["data"]["num_values"])
                props[c["name"]] = _expand(present, [bool(x) for x in data]) if len(data) != F else [bool(v) if p else None for v, p in zip(data, present)]
            elif dt in (DT2_INT_64, DT2_UINT_64):
                present = _bitset(arr, S["present"], F)
                data = _int_data(arr, S["data"], S["data"]["num_values"], signed=(dt == DT2_INT_64))
                props[c["name"]] = _expand(present, [int(x) for x in data])
            elif dt in (DT2_FLOAT, DT2_DOUBLE):
                present = _bitset(arr, S["present"], F)
                d = S["data"]
                data = np.frombuffer(tile, dtype="<f4" if dt == DT2_FLOAT else "<f8", count=d["num_values"], offset=d["offset"])
                props[c["name"]] = _expand(present, [float(x) for x in data])
            elif dt == DT2_STRING and ct == abi.CT_DICTIONARY:
                present = _bitset(arr, S["present"], F)
                words = _dictionary(arr, tile, S["length"], S["dictionary"])
                idx = _decode(arr, S["data"], abi.OP_RLE_U64, S["data"]["num_values"]).astype(np.int64)
                props[c["name"]] = _expand(present, [words[i] for i in idx])
            elif dt == DT2_STRING and ct == abi.CT_LOCALIZED_DICTIONARY:
                words = _dictionary(arr, tile, S["length"], S["dictionary"])
                for s in c["streams"]:
                    if s["name"].startswith("present_"):
                        sub = s["name"][len("present_"):]
                        present = _bitset(arr, s, F)
                        idx = _decode(arr, S[sub], abi.OP_RLE_U64, S[sub]["num_values"]).astype(np.int64)
                        key = c["name"] if sub == c["name"] else c["name"] + ":" + sub
                        props[key] = _expand(present, [words[i] for i in idx])
            else:
                raise ValueError("column %s: data type %d / column type %d not supported" % (c["name"], dt, ct))
        result.append((L["name"], props))
    return result
