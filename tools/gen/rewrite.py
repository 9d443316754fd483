"""gen-2b tile re-writer (workload tooling, like the rest of tools/gen): the "RLE topology streams" variant of BASELINE config 2.

The committed fixture tiles mix FastPFOR and ORC RLE for their topology streams (the converter picks the shorter). The README's
"Reduction 1" column and BASELINE config 2 are about tiles written with allowFastPforForTopologyStreams = false, so this module
rewrites a gen-2b tile with every geometry_offsets / part_offsets / ring_offsets stream as ORC RLE v1: the FastPFOR streams
are decoded by a caller-supplied function and re-encoded with the restated EncodingUtils.encodeRle (covt_gen.c); metadata
(numValues, byteLength, encoding ordinal) and payload are re-serialised, everything else is copied byte for byte.
Container grammar: SURVEY §A.1 (no gen-2b reader or writer exists in the reference at HEAD)."""
import numpy as np

from . import gen

ENC_RLE, ENC_FAST_PFOR_DELTA_ZIG_ZAG = 5, 9   # StreamEncoding ordinals (J/converter/StreamEncoding.java:3-16)
DT2_GEOMETRY = 6                              # gen-2 data type byte
_GEOM_ORDER = ["geometry_types", "geometry_offsets", "part_offsets", "ring_offsets", "vertex_offsets", "vertex_buffer", "index_buffer"]
TOPOLOGY = ("geometry_offsets", "part_offsets", "ring_offsets")


def _varint(b, p):
    v = s = 0
    for _ in range(4):
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            break
    return v, p


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def walk(tile):
    """-> (version, [layer dict(name, extent, num_features, columns=[dict(name, data_type, column_type, streams=[dict(name,
    num_values, byte_length, encoding, offset)])])]); payload order: [id] | geometry streams in their fixed order | the rest."""
    b = memoryview(tile)
    p = 0
    version, p = _varint(b, p)
    n_layers, p = _varint(b, p)
    layers = []
    for _ in range(n_layers):
        n, p = _varint(b, p)
        name = bytes(b[p:p + n])
        p += n
        extent, p = _varint(b, p)
        num_features, p = _varint(b, p)
        n_cols, p = _varint(b, p)
        cols = []
        for _ in range(n_cols):
            n, p = _varint(b, p)
            cname = bytes(b[p:p + n])
            p += n
            dt, ct = b[p], b[p + 1]
            p += 2
            n_streams, p = _varint(b, p)
            streams = []
            for _ in range(n_streams):
                n, p = _varint(b, p)
                sname = bytes(b[p:p + n]).decode("utf-8")
                p += n
                nv, p = _varint(b, p)
                bl, p = _varint(b, p)
                streams.append({"name": sname, "num_values": nv, "byte_length": bl, "encoding": b[p]})
                p += 1
            cols.append({"name": cname, "data_type": dt, "column_type": ct, "streams": streams})
        for c in cols:
            ss = c["streams"]
            if c["data_type"] == DT2_GEOMETRY:
                ss = sorted(ss, key=lambda s: _GEOM_ORDER.index(s["name"]))
            for s in ss:
                s["offset"] = p
                p += s["byte_length"]
        layers.append({"name": name, "extent": extent, "num_features": num_features, "columns": cols})
    if p != len(tile):
        raise ValueError("gen-2b walk ended at %d of %d" % (p, len(tile)))
    return version, layers


def topology_pfor_streams(tile):
    """[(offset, byte_length, num_values)] of the FastPFOR topology streams of one tile (what the caller has to decode)."""
    out = []
    for L in walk(tile)[1]:
        for c in L["columns"]:
            if c["data_type"] == DT2_GEOMETRY:
                out += [(s["offset"], s["byte_length"], s["num_values"]) for s in c["streams"]
                        if s["name"] in TOPOLOGY and s["encoding"] == ENC_FAST_PFOR_DELTA_ZIG_ZAG]
    return out


def transcode_topology_to_rle(tile, decoded, keep_pfor=False):
    """decoded: {offset: int32 ndarray} for every entry of topology_pfor_streams(tile). Returns the rewritten tile bytes.
    keep_pfor=True re-serialises the tile without touching any stream (must give back the input byte for byte)."""
    tile = bytes(tile)
    version, layers = walk(tile)
    out = bytearray(_enc_varint(version) + _enc_varint(len(layers)))
    for L in layers:
        payloads = []
        for c in L["columns"]:
            for s in c["streams"]:
                data = tile[s["offset"]:s["offset"] + s["byte_length"]]
                if not keep_pfor and c["data_type"] == DT2_GEOMETRY and s["name"] in TOPOLOGY and s["encoding"] == ENC_FAST_PFOR_DELTA_ZIG_ZAG:
                    vals = np.asarray(decoded[s["offset"]])
                    assert len(vals) == s["num_values"]
                    data = bytes(gen.encode_rle(vals.astype(np.int64), signed=False))
                    s["encoding"] = ENC_RLE
                    s["byte_length"] = len(data)
                payloads.append((s["offset"], data))
        out += _enc_varint(len(L["name"])) + L["name"] + _enc_varint(L["extent"]) + _enc_varint(L["num_features"]) + _enc_varint(len(L["columns"]))
        for c in L["columns"]:
            out += _enc_varint(len(c["name"])) + c["name"] + bytes([c["data_type"], c["column_type"]]) + _enc_varint(len(c["streams"]))
            for s in c["streams"]:
                sn = s["name"].encode("utf-8")
                out += _enc_varint(len(sn)) + sn + _enc_varint(s["num_values"]) + _enc_varint(s["byte_length"]) + bytes([s["encoding"]])
        for _, data in sorted(payloads, key=lambda t: t[0]):  # the original payload order
            out += data
    return bytes(out)
