"""Canonical geometry form shared by the MVT ground truth and the assembled COVT buffers (test infrastructure).

canonical = (rings_per_feature int32[F], vertices_per_ring int32[R], coords int32[2V]) with polygon rings stripped of a
duplicated closing vertex ("after dropping a duplicated closing vertex on either side").
"""
import hashlib

import numpy as np


def canonical_from_assembled(types, geom_off, part_off, ring_off, coords, strip_closing=True):
    """Inputs are one layer's slices decoded WITHOUT COVT_FLAG_CLOSE_RINGS:
    types u8[F], geom_off i32[F+1], part_off i32[P+1], ring_off i32[R+1], coords i32[2V]."""
    types = np.asarray(types)
    F = len(types)
    geom_off = np.asarray(geom_off, dtype=np.int64)
    part_off = np.asarray(part_off, dtype=np.int64)
    ring_off = np.asarray(ring_off, dtype=np.int64)
    ring_start = part_off[geom_off]
    rpf = np.diff(ring_start)
    vpr = np.diff(ring_off)
    xy = np.asarray(coords, dtype=np.int32).reshape(-1, 2)
    if strip_closing and len(vpr):
        fidx = np.repeat(np.arange(F), rpf)
        is_poly = np.isin(types[fidx], (2, 5))
        first = ring_off[:-1]
        last = np.maximum(ring_off[1:] - 1, 0)
        cand = is_poly & (vpr > 1)
        dup = np.zeros(len(vpr), dtype=bool)
        idx = np.nonzero(cand)[0]
        dup[idx] = (xy[first[idx]] == xy[last[idx]]).all(axis=1)
        keep = np.ones(len(xy), dtype=bool)
        keep[last[dup]] = False
        xy = xy[keep]
        vpr = vpr - dup
    return rpf.astype(np.int32), vpr.astype(np.int32), np.ascontiguousarray(xy, dtype=np.int32).ravel()


def digest(canon):
    h = hashlib.blake2b(digest_size=16)
    for a in canon:
        h.update(np.ascontiguousarray(a, dtype=np.int32).tobytes())
        h.update(b"|")
    return h.hexdigest()


def layer_slices(layers_row, buffers, abi):
    """Returns (types, geom_off, part_off, ring_off, coords) numpy views of one covt_layer row."""
    L = layers_row
    F = int(L["streams"][abi.SLOT_TYPES]["num_values"])
    o = L["out"]
    types = buffers[abi.BUF_S_GEOMETRY_TYPES][int(o[abi.BUF_S_GEOMETRY_TYPES]):][:F]
    g = buffers[abi.BUF_A_GEOM_OFFSETS][int(o[abi.BUF_A_GEOM_OFFSETS]):][:F + 1]
    p = buffers[abi.BUF_A_PART_OFFSETS][int(o[abi.BUF_A_PART_OFFSETS]):][:int(L["n_parts"]) + 1]
    r = buffers[abi.BUF_A_RING_OFFSETS][int(o[abi.BUF_A_RING_OFFSETS]):][:int(L["n_rings"]) + 1]
    c = buffers[abi.BUF_A_COORDS][int(o[abi.BUF_A_COORDS]):][:2 * int(L["n_coords"])]
    return types, g, p, r, c


def property_token(v):
    """Canonical token of one property value (None = the feature has no such property). Numbers: integers in decimal, reals as
    the hex of their float32 image (COVT FLOAT columns are 32-bit; MVT may carry the same value as float or double)."""
    import struct
    if v is None:
        return "~"
    if isinstance(v, bool):
        return "b1" if v else "b0"
    if isinstance(v, int):
        return "i%d" % v
    if isinstance(v, float):
        return "f" + struct.pack("<f", v).hex()
    return "s" + str(v)


def property_digest(values):
    """blake2b digest of one property column: the tokens of all features in feature order."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for v in values:
        t = property_token(v).encode("utf-8")
        h.update(len(t).to_bytes(4, "little"))
        h.update(t)
    return h.hexdigest()
