"""Minimal Mapbox Vector Tile (protobuf) geometry reader — test infrastructure for the .mvt/.pbf ground truth.

Returns, per layer, the three canonical arrays the parity tests hash (see canonical_from_rings):
  rings_per_feature int32[F], vertices_per_ring int32[R], coords int32[2V]  (no closing vertices).
"""
import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf, pos, end):
    """Yields (field_number, wire_type, value) where value is an int or a (start, end) span."""
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fn, wt, v
        elif wt == 2:
            n, pos = _varint(buf, pos)
            yield fn, wt, (pos, pos + n)
            pos += n
        elif wt == 5:
            yield fn, wt, (pos, pos + 4)
            pos += 4
        elif wt == 1:
            yield fn, wt, (pos, pos + 8)
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)


def _packed_uint32(buf, start, end):
    out = []
    pos = start
    while pos < end:
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _decode_geometry(cmds, gtype):
    """MVT command stream -> list of rings/lines/points (lists of (x, y)). ClosePath adds no vertex."""
    rings = []
    cur = None
    x = y = 0
    i = 0
    n = len(cmds)
    while i < n:
        c = cmds[i]
        i += 1
        cid, cnt = c & 7, c >> 3
        if cid == 1:  # MoveTo
            for _ in range(cnt):
                dx, dy = cmds[i], cmds[i + 1]
                i += 2
                x += (dx >> 1) ^ -(dx & 1)
                y += (dy >> 1) ^ -(dy & 1)
                if gtype == 1:
                    rings.append([(x, y)])
                else:
                    cur = [(x, y)]
                    rings.append(cur)
        elif cid == 2:  # LineTo
            for _ in range(cnt):
                dx, dy = cmds[i], cmds[i + 1]
                i += 2
                x += (dx >> 1) ^ -(dx & 1)
                y += (dy >> 1) ^ -(dy & 1)
                cur.append((x, y))
        elif cid == 7:  # ClosePath
            pass
        else:
            raise ValueError("bad MVT command %d" % cid)
    return rings


def _value(buf, start, end):
    """vector_tile.Value -> (kind, python value); kind in string/float/double/int/uint/sint/bool."""
    import struct
    for fn, wt, v in _fields(buf, start, end):
        if fn == 1:
            return "string", bytes(buf[v[0]:v[1]]).decode("utf-8")
        if fn == 2:
            return "float", struct.unpack("<f", bytes(buf[v[0]:v[1]]))[0]
        if fn == 3:
            return "double", struct.unpack("<d", bytes(buf[v[0]:v[1]]))[0]
        if fn == 4:
            return "int", v - (1 << 64) if v >= (1 << 63) else v
        if fn == 5:
            return "uint", v
        if fn == 6:
            return "sint", (v >> 1) ^ -(v & 1)
        if fn == 7:
            return "bool", bool(v)
    return "none", None


def read_layers(data, with_properties=False):
    """-> list of dict(name, extent, features=[(id, gtype, rings)]); with_properties adds
    properties=[{key: (kind, value)} per feature] (vector_tile.Layer keys / values / Feature.tags)."""
    buf = memoryview(data)
    layers = []
    for fn, wt, v in _fields(buf, 0, len(buf)):
        if fn != 3 or wt != 2:
            continue
        name, extent, feats, keys, values, tags = None, 4096, [], [], [], []
        for lfn, lwt, lv in _fields(buf, v[0], v[1]):
            if lfn == 1:
                name = bytes(buf[lv[0]:lv[1]]).decode("utf-8")
            elif lfn == 5:
                extent = lv
            elif lfn == 3 and with_properties:
                keys.append(bytes(buf[lv[0]:lv[1]]).decode("utf-8"))
            elif lfn == 4 and with_properties:
                values.append(_value(buf, lv[0], lv[1]))
            elif lfn == 2:
                fid, gtype, cmds, ftags = 0, 0, [], []
                for ffn, fwt, fv in _fields(buf, lv[0], lv[1]):
                    if ffn == 1:
                        fid = fv
                    elif ffn == 2 and with_properties:
                        ftags = _packed_uint32(buf, fv[0], fv[1])
                    elif ffn == 3:
                        gtype = fv
                    elif ffn == 4:
                        cmds = _packed_uint32(buf, fv[0], fv[1])
                feats.append((fid, gtype, _decode_geometry(cmds, gtype)))
                tags.append(ftags)
        layer = {"name": name, "extent": extent, "features": feats}
        if with_properties:
            layer["properties"] = [{keys[t[i]]: values[t[i + 1]] for i in range(0, len(t) - 1, 2)} for t in tags]
        layers.append(layer)
    return layers


def _strip_closing(ring):
    if len(ring) > 1 and ring[0] == ring[-1]:
        return ring[:-1]
    return ring


def canonical_from_features(features, strip_closing=True):
    """features: [(id, gtype, rings)] -> (rings_per_feature, vertices_per_ring, coords) int32 arrays."""
    rpf, vpr, xy = [], [], []
    for _, gtype, rings in features:
        rpf.append(len(rings))
        for r in rings:
            if strip_closing and gtype == 3:
                r = _strip_closing(r)
            vpr.append(len(r))
            for p in r:
                xy.append(p[0])
                xy.append(p[1])
    return (np.asarray(rpf, dtype=np.int32), np.asarray(vpr, dtype=np.int32), np.asarray(xy, dtype=np.int32))
