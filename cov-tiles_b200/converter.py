"""CovtConverter — the tile WRITER over the GPU stream encoders.

Mirrors the stream selection of the reference converter (J/converter/CovtConverter.java): every stream is encoded BOTH ways on the GPU
(one covt_encode_streams call for all candidates of all layers of all tiles of a batch) and the shorter one is kept —
  geometry_types                 Byte-RLE                                                   (:876-879)
  geometry / part / ring offsets ORC RLE, or FastPFOR(zigzag, delta) if not longer          (addOffsets :899-920)
  PLAIN vertex buffer            zigzag-delta coordinates as varints, or FastPFOR if not longer   (convertUnorderedGeometryColumn :641-668)
  ICE_MORTON_CODE                sorted dictionary of Morton codes (deltas without zigzag) + vertex_offsets (zigzag delta), each as
                                 varints if strictly shorter than FastPFOR                   (:671-769, :771-856, :939-948)
  ids                            the shortest of RLE / plain varints / zigzag-delta varints  (convertIdColumn :546-569, without its label bug)
— then the layer metadata (gen-2b: the grammar of the committed fixtures; gen-3: convertLayerMetadata :383-426,
addOptimizedStreamMetadata :478-483) and the payloads in the order the decoder reads them are put together on the host (a few bytes
per stream). Property columns are not written (the decode side reads them; the reference converter's property writer is
CovtConverter.java:988-1221). tests/test_gpu_converter.py: tiles equal those of the CPU restatement (tools/gen) byte for byte and
decode back to the layers they were written from.

A layer is a dict: name, extent (default 4096), types (u8 GeometryType ordinals), geom / part / ring (count streams, may be empty),
xy (flat x, y ints, no closing vertices), ids (optional int64), index_buffer (optional int32: the INDEX_BUFFER extension stream)."""
import ctypes as C

import numpy as np

from . import abi

ALLOW_PFOR_TOPOLOGY, ALLOW_PFOR_VERTEX, ICE_MORTON, ID_DELTA_VARINT, FORCE_VARINT_VERTEX, FORCE_RLE_TOPOLOGY = 1, 2, 4, 8, 16, 32
OPTIMIZED_METADATA = 0x40

_GEN2B_NAMES = {abi.SLOT_TYPES: b"geometry_types", abi.SLOT_GEOM: b"geometry_offsets", abi.SLOT_PART: b"part_offsets",
                abi.SLOT_RING: b"ring_offsets", abi.SLOT_VOFF: b"vertex_offsets", abi.SLOT_VBUF: b"vertex_buffer", abi.SLOT_INDEX: b"index_buffer"}
_GEN3_TYPES = {abi.SLOT_TYPES: 4, abi.SLOT_GEOM: 5, abi.SLOT_PART: 6, abi.SLOT_RING: 7, abi.SLOT_VOFF: 8, abi.SLOT_VBUF: 9, abi.SLOT_INDEX: 12}


def _varint(v):
    out = bytearray()
    v = int(v)
    while True:
        if v < 0x80:
            out.append(v)
            return bytes(out)
        out.append(0x80 | (v & 0x7F))
        v >>= 7


def _string(b):
    return _varint(len(b)) + b


def _compact_even_bits(v):
    v = v & 0x55555555
    v = (v | (v >> 1)) & 0x33333333
    v = (v | (v >> 2)) & 0x0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF
    return v


def _morton_codes(xy, num_bits):
    """GeometryUtils.encodeMorton :23-32 on arrays (host side: only to build the sorted dictionary; the bytes come from the GPU)."""
    half = (2 << (num_bits - 2)) // 2
    x = (xy[:, 0].astype(np.int64) + half) & 0xFFFFFFFF
    y = (xy[:, 1].astype(np.int64) + half) & 0xFFFFFFFF
    code = np.zeros(len(xy), dtype=np.int64)
    for i in range(num_bits):
        code |= ((x >> i) & 1) << (2 * i) | ((y >> i) & 1) << (2 * i + 1)
    return code & 0xFFFFFFFF


def _morton_xy(codes, num_bits):
    half = (2 << (num_bits - 2)) // 2
    mask = (1 << num_bits) - 1
    x = (_compact_even_bits(codes) & mask) - half
    y = (_compact_even_bits(codes >> 1) & mask) - half
    return np.stack([x, y], axis=1).astype(np.int32)


class _Plan:
    """Candidate encodings of one batch: add() queues (op, values) and returns a handle; run() encodes them all in one GPU call."""

    def __init__(self, decoder):
        self.dec = decoder
        self.buf = bytearray()
        self.req = []
        self.out = None

    def add(self, op, values, num_bits=0):
        morton = op in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON)
        a = np.ascontiguousarray(values, dtype=np.int32 if morton else abi.op_dtype(op))
        self.buf += bytes((-len(self.buf)) % 8)
        self.req.append((len(self.buf), a.size // 2 if morton else a.size, op, num_bits))
        self.buf += a.tobytes()
        return len(self.req) - 1

    def run(self):
        self.out = []
        if not self.req:
            return
        descs = (abi.EncodeDesc * len(self.req))()
        for i, (off, n, op, nb) in enumerate(self.req):
            descs[i] = abi.EncodeDesc(value_offset=off, num_values=n, op=op, num_bits=nb)
        self.buf += bytes(8)
        res = self.dec.encode_streams(np.frombuffer(bytes(self.buf), np.uint8), descs)
        arena = res.buffer(abi.BUF_STREAM_ARENA)
        for i in range(len(self.req)):
            if descs[i].status != abi.OK:
                res.free()
                raise RuntimeError("stream encode failed (status %d)" % descs[i].status)
            self.out.append(bytes(arena[descs[i].out_offset:descs[i].out_offset + descs[i].byte_length]))
        res.free()


class CovtConverter:
    def __init__(self, decoder):
        self.dec = decoder

    # ---- what to encode for one layer (handles into the plan) + how to choose afterwards ------------------------------------
    @staticmethod
    def _plan_layer(plan, L, options):
        extent = int(L.get("extent", 4096))
        num_bits = extent.bit_length()  # 32 - numberOfLeadingZeros(extent), CovtParser.java:77
        types = np.ascontiguousarray(L["types"], dtype=np.uint8)
        xy = np.ascontiguousarray(L["xy"], dtype=np.int32).reshape(-1, 2)
        allow_t = bool(options & ALLOW_PFOR_TOPOLOGY) and not (options & FORCE_RLE_TOPOLOGY)
        allow_v = bool(options & ALLOW_PFOR_VERTEX) and not (options & FORCE_VARINT_VERTEX)
        P = {"extent": extent, "n_features": len(types), "name": L.get("name", "layer"), "streams": {}}
        P["streams"][abi.SLOT_TYPES] = ("fixed", abi.ENC_BYTE_RLE, len(types), plan.add(abi.OP_BYTE_RLE, types))
        for slot, key in ((abi.SLOT_GEOM, "geom"), (abi.SLOT_PART, "part"), (abi.SLOT_RING, "ring")):
            v = np.ascontiguousarray(L.get(key, []), dtype=np.int32)
            if len(v):  # addOffsets: FastPFOR if pl <= rl
                rle = plan.add(abi.OP_RLE_U32, v)
                pf = plan.add(abi.OP_PFOR_ZZ_DELTA, v) if allow_t else None
                P["streams"][slot] = ("le", (abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, pf), (abi.ENC_RLE, rle), len(v))
        if options & ICE_MORTON:
            P["column_type"] = abi.CT_ICE_MORTON_CODE
            codes = _morton_codes(xy, num_bits)
            dictionary, offs = np.unique(codes, return_inverse=True)
            offs = offs.astype(np.int32)
            dict_xy = _morton_xy(dictionary, num_bits)
            vi = plan.add(abi.OP_VARINT_ZZ_DELTA, offs)
            pf = plan.add(abi.OP_PFOR_ZZ_DELTA, offs) if allow_v else None
            P["streams"][abi.SLOT_VOFF] = ("lt", (abi.ENC_VARINT_DELTA_ZIG_ZAG, vi), (abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, pf), len(offs))
            vi = plan.add(abi.OP_VARINT_DELTA_MORTON, dict_xy, num_bits)
            pf = plan.add(abi.OP_PFOR_DELTA_MORTON, dict_xy, num_bits) if allow_v else None
            P["streams"][abi.SLOT_VBUF] = ("lt", (abi.ENC_VARINT_DELTA_ZIG_ZAG, vi), (abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, pf), len(dictionary))
        else:
            P["column_type"] = abi.CT_PLAIN
            flat = xy.reshape(-1)
            vi = plan.add(abi.OP_VARINT_ZZ_DELTA_XY, flat)
            pf = plan.add(abi.OP_PFOR_ZZ_DELTA_XY, flat) if allow_v else None
            P["streams"][abi.SLOT_VBUF] = ("le", (abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, pf), (abi.ENC_VARINT_DELTA_ZIG_ZAG, vi), len(flat))
        if L.get("ids") is not None:
            ids = np.ascontiguousarray(L["ids"], dtype=np.int64)
            dv = plan.add(abi.OP_VARINT_ZZ_DELTA_64, ids)
            if options & ID_DELTA_VARINT:
                P["streams"][abi.SLOT_ID] = ("fixed", abi.ENC_VARINT_DELTA_ZIG_ZAG, len(ids), dv)
            else:
                P["streams"][abi.SLOT_ID] = ("ids", plan.add(abi.OP_RLE_U64, ids), plan.add(abi.OP_VARINT_U64, ids), dv, len(ids))
        if L.get("index_buffer") is not None and len(L["index_buffer"]):
            idx = np.ascontiguousarray(L["index_buffer"], dtype=np.int32)
            P["streams"][abi.SLOT_INDEX] = ("fixed", abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, len(idx), plan.add(abi.OP_PFOR_ZZ_DELTA, idx))
        return P

    @staticmethod
    def _choose(plan, spec):
        """-> (encoding ordinal, numValues, bytes)"""
        kind = spec[0]
        if kind == "fixed":
            return spec[1], spec[2], plan.out[spec[3]]
        if kind == "ids":  # the shortest of RLE / plain / zigzag-delta
            rl, vl, dl = plan.out[spec[1]], plan.out[spec[2]], plan.out[spec[3]]
            if len(rl) < len(vl) and len(rl) < len(dl):
                return abi.ENC_RLE, spec[4], rl
            if len(dl) < len(vl):
                return abi.ENC_VARINT_DELTA_ZIG_ZAG, spec[4], dl
            return abi.ENC_VARINT, spec[4], vl
        (e1, h1), (e2, h2), n = spec[1], spec[2], spec[3]
        if h1 is None:
            return e2, n, plan.out[h2]
        if h2 is None:
            return e1, n, plan.out[h1]
        a, b = plan.out[h1], plan.out[h2]
        first = len(a) <= len(b) if kind == "le" else len(a) < len(b)  # "le": the first wins ties, "lt": only if strictly shorter
        return (e1, n, a) if first else (e2, n, b)

    @staticmethod
    def _write_layer(plan, P, container, options):
        S = {slot: CovtConverter._choose(plan, spec) for slot, spec in P["streams"].items()}
        name = P["name"].encode() if isinstance(P["name"], str) else bytes(P["name"])
        have_id = abi.SLOT_ID in S
        m = bytearray()
        geom_slots = [s for s in (abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VOFF, abi.SLOT_VBUF, abi.SLOT_INDEX) if s in S]
        if container == abi.CONTAINER_GEN2B:
            def stream(nm, s):
                enc, nv, data = S[s]
                return _string(nm) + _varint(nv) + _varint(len(data)) + bytes([enc])
            m += _string(name) + _varint(P["extent"]) + _varint(P["n_features"]) + _varint(1 + have_id)
            if have_id:
                m += _string(b"id") + bytes([4, 0]) + _varint(1) + stream(b"data", abi.SLOT_ID)
            m += _string(b"geometry") + bytes([6, P["column_type"]]) + _varint(len(geom_slots))
            order = [abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VBUF, abi.SLOT_INDEX]
            if abi.SLOT_VOFF in S:  # ICE columns list vertex_offsets, vertex_buffer first (insertion order of the converter's map)
                order = [abi.SLOT_VOFF, abi.SLOT_VBUF, abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_INDEX]
            for s in order:
                if s in S:
                    m += stream(_GEN2B_NAMES[s], s)
        else:
            def stream(s):
                enc, nv, data = S[s]
                return bytes([_GEN3_TYPES.get(s, 1) << 4 | enc]) + _varint(nv) + _varint(len(data))
            optimized = bool(options & OPTIMIZED_METADATA)
            m += bytes([1 << 1 | int(optimized)])
            m += _varint(int(name)) if optimized else _string(name)
            m += _varint(P["extent"]) + _varint(P["n_features"]) + _varint(1 + have_id)
            if have_id:
                m += _varint(0) + bytes([4 << 3 | 0]) + bytes([1 << 4 | S[abi.SLOT_ID][0]]) + _varint(S[abi.SLOT_ID][1]) + _varint(len(S[abi.SLOT_ID][2]))
            m += _varint(1) if (optimized or not have_id) else _string(b"geometry")
            m += bytes([8 << 3 | P["column_type"]])
            for s in (abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VOFF, abi.SLOT_INDEX, abi.SLOT_VBUF):  # VERTEX_BUFFER ends the column
                if s in S:
                    m += stream(s)
        payload = b"".join(S[s][2] for s in (abi.SLOT_ID, abi.SLOT_TYPES, abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING, abi.SLOT_VOFF, abi.SLOT_VBUF,
                                             abi.SLOT_INDEX) if s in S)
        return bytes(m) + payload

    # ---- public ---------------------------------------------------------------------------------------------------------------
    def convert_tiles(self, tiles, container=abi.CONTAINER_GEN2B, options=ALLOW_PFOR_TOPOLOGY | ALLOW_PFOR_VERTEX):
        """tiles: list of tiles, each a list of layer dicts (a layer may carry its own 'options'). Returns a list of tile byte strings.
        All streams of all tiles are encoded in ONE covt_encode_streams call."""
        plan = _Plan(self.dec)
        plans = [[self._plan_layer(plan, L, L.get("options", options)) for L in layers] for layers in tiles]
        plan.run()
        out = []
        for layers, lp in zip(tiles, plans):
            t = bytearray()
            if container == abi.CONTAINER_GEN2B:
                t += _varint(1) + _varint(len(layers))
            for L, P in zip(layers, lp):
                t += self._write_layer(plan, P, container, L.get("options", options))
            out.append(bytes(t))
        return out

    def convert_tile(self, layers, container=abi.CONTAINER_GEN2B, options=ALLOW_PFOR_TOPOLOGY | ALLOW_PFOR_VERTEX):
        return self.convert_tiles([layers], container, options)[0]
