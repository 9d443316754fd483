"""GPU stream encoders (covt_encode_streams): byte-identical to the reference's encoders.

Two anchors: (1) the fixture tiles themselves — every stream the reference converter wrote decodes on the GPU and re-encodes on the GPU
to its own bytes (no oracle, no CPU encoder in between); (2) the CPU restatement of EncodingUtils (tools/gen/covt_gen.c, itself pinned
by re-encoding the fixtures byte for byte in tests/test_oracle_fixtures.py) on random inputs of every size class. All calls go
through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 5, 31, 32, 33, 127, 128, 129, 130, 131, 255, 256, 257, 300, 511, 512, 513, 1000, 4095, 4096, 4097, 9000, 65535, 65536,
         65537, 65536 + 255, 65536 + 256, 70000, 140000]


def _encode_many(covt, decoder, cases, flags=None):
    """cases: list of (op, values ndarray, num_bits). ONE covt_encode_streams call. Returns the list of encoded byte arrays."""
    abi = covt.abi
    buf = bytearray()
    descs = (abi.EncodeDesc * len(cases))()
    for i, (op, v, nbits) in enumerate(cases):
        morton = op in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON)
        a = np.ascontiguousarray(v, dtype=np.int32 if morton else abi.op_dtype(op))
        buf += bytes((-len(buf)) % 8)
        descs[i] = abi.EncodeDesc(value_offset=len(buf), num_values=(a.size // 2 if morton else a.size), op=op, num_bits=nbits)
        buf += a.tobytes()
    buf += bytes(8)
    res = decoder.encode_streams(np.frombuffer(bytes(buf), np.uint8), descs, abi.FLAG_DEFAULT if flags is None else flags)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    out = []
    for i in range(len(cases)):
        assert descs[i].status == 0, "case %d op %s: status %d" % (i, abi.OP_NAMES[cases[i][0]], descs[i].status)
        assert descs[i].out_offset % 16 == 0
        out.append(arena[descs[i].out_offset:descs[i].out_offset + descs[i].byte_length].copy())
    t = res.timing()
    assert t["payload_bytes"] == sum(len(o) for o in out)
    res.free()
    return out


def _check(abi, cases, got, want):
    for i, ((op, v, nbits), g, w) in enumerate(zip(cases, got, want)):
        w = np.asarray(w, dtype=np.uint8)
        if len(g) != len(w) or not np.array_equal(g, w):
            bad = np.nonzero(g[:min(len(g), len(w))] != w[:min(len(g), len(w))])[0]
            raise AssertionError("case %d op %s n=%d: %d bytes vs %d, first difference at %s" % (
                i, abi.OP_NAMES[op], len(v), len(g), len(w), bad[:4]))


def test_varint_encoders_equal_the_reference_encoders(covt, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(11)
    cases, want = [], []
    for n in SIZES:
        for span in (40, 5000, 1 << 20, 1 << 27):
            v = rng.integers(-span, span, n).astype(np.int32)
            cases.append((abi.OP_VARINT_U32, np.abs(v), 0)); want.append(gen.encode_varints(np.abs(v).astype(np.int64)))
            cases.append((abi.OP_VARINT_ZZ, v, 0)); want.append(gen.encode_varints(v.astype(np.int64), zigzag=True))
            w = np.cumsum(rng.integers(-span, span, n)).astype(np.int32)
            cases.append((abi.OP_VARINT_ZZ_DELTA, w, 0)); want.append(gen.encode_varints(w.astype(np.int64), zigzag=True, delta=True))
            xy = np.cumsum(rng.integers(-span, span, n & ~1)).astype(np.int32)
            cases.append((abi.OP_VARINT_ZZ_DELTA_XY, xy, 0))
            want.append(gen.encode_varints(gen.encode_zigzag_delta_coordinates(xy).astype(np.int64) & 0xFFFFFFFF))
        big = rng.integers(0, 1 << 62, n).astype(np.int64)
        cases.append((abi.OP_VARINT_U64, big, 0)); want.append(gen.encode_varints(big))
        walk = np.cumsum(rng.integers(-(1 << 40), 1 << 40, n)).astype(np.int64)
        cases.append((abi.OP_VARINT_ZZ_DELTA_64, walk, 0)); want.append(gen.encode_varints(walk, zigzag=True, delta=True))
    ext = np.array([0, 1, -1, (1 << 63) - 1, -(1 << 63), 1 << 35, 127, 128], dtype=np.int64)
    cases.append((abi.OP_VARINT_ZZ_DELTA_64, ext, 0)); want.append(gen.encode_varints(ext, zigzag=True, delta=True))
    cases.append((abi.OP_VARINT_U64, ext, 0)); want.append(gen.encode_varints(ext))
    i32 = np.array([0, 1, -1, (1 << 31) - 1, -(1 << 31), 5, -(1 << 31), (1 << 31) - 1], dtype=np.int32)  # int overflow in the deltas
    cases.append((abi.OP_VARINT_ZZ_DELTA, i32, 0)); want.append(gen.encode_varints(i32.astype(np.int64), zigzag=True, delta=True))
    cases.append((abi.OP_VARINT_ZZ_DELTA_XY, i32, 0))
    want.append(gen.encode_varints(gen.encode_zigzag_delta_coordinates(i32).astype(np.int64) & 0xFFFFFFFF))
    _check(abi, cases, _encode_many(covt, decoder, cases), want)


def _rle_inputs(rng, n):
    """values that exercise runs (constant and arithmetic, |delta| <= 127 and beyond), literal groups and their 128 / 130 limits"""
    yield rng.integers(0, 5, n)                                                     # short runs between literals
    yield np.repeat(rng.integers(0, 1000, n // 7 + 1), 7)[:n]                       # runs of 7
    yield np.arange(n) * 3 + 10                                                     # one long arithmetic run (130-value cuts)
    yield np.arange(n)[::-1] * 127                                                  # delta -127
    yield np.arange(n) * 128                                                        # delta 128: not a run
    yield rng.integers(0, 1 << 40, n)                                               # literals only (128-value cuts)
    yield np.concatenate([np.arange(k, k + ln) for k, ln in zip(rng.integers(0, 9999, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))])[:n]
    yield np.concatenate([np.full(int(ln), int(k)) for k, ln in zip(rng.integers(0, 3, n // 40 + 1), rng.integers(1, 300, n // 40 + 1))])[:n]


def test_rle_encoders_equal_the_reference_encoders(covt, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(12)
    cases, want = [], []
    for n in SIZES:
        for v in _rle_inputs(rng, n):
            v = np.asarray(v, dtype=np.int64)
            cases.append((abi.OP_RLE_U64, v, 0)); want.append(gen.encode_rle(v, signed=False))
            s = v - (v.max() // 2 if len(v) else 0)
            cases.append((abi.OP_RLE_S64, s, 0)); want.append(gen.encode_rle(s, signed=True))
            u = (v & 0x7FFFFFFF).astype(np.int32)
            cases.append((abi.OP_RLE_U32, u, 0)); want.append(gen.encode_rle(u.astype(np.int64), signed=False))
            b = (v & 0xFF).astype(np.uint8)
            cases.append((abi.OP_BYTE_RLE, b, 0)); want.append(gen.encode_byte_rle(b))
    _check(abi, cases, _encode_many(covt, decoder, cases), want)


def test_fastpfor_encoder_equals_the_reference_encoder(covt, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(13)
    cases, want = [], []
    for n in SIZES:
        for span, p_exc in ((3, 0.0), (60, 0.02), (4000, 0.08), (1 << 20, 0.3), (1 << 30, 0.0)):
            d = rng.integers(-span, span, n)
            d = np.where(rng.random(n) < p_exc, d * 977, d)  # outliers -> exceptions of several widths
            v = np.cumsum(d).astype(np.int32)
            cases.append((abi.OP_PFOR_ZZ_DELTA, v, 0)); want.append(gen.encode_fastpfor(v, zigzag=True, delta=True))
            xy = v[:n & ~1]
            cases.append((abi.OP_PFOR_ZZ_DELTA_XY, xy, 0))
            want.append(gen.encode_fastpfor(gen.encode_zigzag_delta_coordinates(xy), zigzag=False, delta=False))
    cases.append((abi.OP_PFOR_ZZ_DELTA, np.zeros(700, np.int32), 0)); want.append(gen.encode_fastpfor(np.zeros(700, np.int32), zigzag=True, delta=True))
    full = rng.integers(-(1 << 31), 1 << 31, 1024).astype(np.int32)  # b = 32
    cases.append((abi.OP_PFOR_ZZ_DELTA, full, 0)); want.append(gen.encode_fastpfor(full, zigzag=True, delta=True))
    _check(abi, cases, _encode_many(covt, decoder, cases), want)


@pytest.mark.parametrize("nbits,no_shift", [(13, False), (14, False), (13, True)])
def test_morton_encoders(covt, gen, decoder, nbits, no_shift):
    """encodeMorton + delta (no zigzag) + varint / FastPFOR vs the CPU restatement, and back through the GPU decoder."""
    abi = covt.abi
    rng = np.random.default_rng(14 + nbits)
    flags = abi.FLAG_DEFAULT | (abi.FLAG_MORTON_NO_SHIFT if no_shift else 0)
    cases, want, verts = [], [], []
    half = (1 << (nbits - 1)) if no_shift else (2 << (nbits - 2)) // 2  # no_shift: the decoder sign-extends the num_bits-bit value
    for n in [0, 1, 2, 33, 255, 256, 257, 4097, 70000]:
        x = rng.integers(0, 1 << nbits, n) - half
        y = rng.integers(0, 1 << nbits, n) - half
        if no_shift:  # the older converter: no extent/2 shift, the low num_bits bits of the (possibly negative) coordinate are interleaved
            codes = np.zeros(n, np.int64)
            for i in range(nbits):
                codes |= ((x >> i) & 1) << (2 * i) | ((y >> i) & 1) << (2 * i + 1)
        else:
            codes = np.array([gen.encode_morton(int(a), int(b), nbits) for a, b in zip(x, y)], dtype=np.int64)
        order = np.argsort(codes, kind="stable")  # the converter writes a SORTED vertex dictionary
        x, y, codes = x[order], y[order], codes[order]
        xy = np.stack([x, y], axis=1).astype(np.int32)
        verts.append(xy)
        cases.append((abi.OP_VARINT_DELTA_MORTON, xy, nbits)); want.append(gen.encode_varints(codes, delta=True))
        cases.append((abi.OP_PFOR_DELTA_MORTON, xy, nbits)); want.append(gen.encode_fastpfor(codes.astype(np.int32), zigzag=False, delta=True))
    got = _encode_many(covt, decoder, cases, flags)
    _check(abi, cases, got, want)
    for (op, xy, _), enc in zip(cases, got):  # and back
        if len(xy) == 0:
            continue
        vals, st, cons = decoder.decode_stream(np.concatenate([enc, np.zeros(64, np.uint8)]), op, byte_length=len(enc), num_values=len(xy),
                                               num_bits=nbits, flags=flags)
        assert st == 0 and cons == len(enc) and np.array_equal(vals.reshape(-1, 2), xy)


def test_every_fixture_stream_reencodes_to_its_own_bytes(covt, decoder, fixtures):
    """All 129 fixture tiles: decode on the GPU (batch path), encode every decoded stream on the GPU in one call, compare with the
    bytes the reference converter wrote. Exceptions: FastPFOR streams of more than 65 536 values (the Java encoder leaks
    stale array contents into don't-care padding bits of pages >= 2) and the three mislabelled ICE layers."""
    abi = covt.abi
    blob, offs = util.concat_tiles([b for _, b in fixtures])
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    layers = res.layers
    bufs = {b: res.buffer(b) for b in set(abi.SLOT_BUF)}
    cases, where = [], []
    for L in layers:
        key = "%s/%s" % (fixtures[int(L["tile"])][0], util.layer_name(blob, L))
        for s in range(abi.NUM_SLOTS):
            S = L["streams"][s]
            if S["encoding"] == abi.ENC_ABSENT or S["op"] == abi.OP_NONE or S["status"] != 0:
                continue
            if key in util.KNOWN_MISLABELLED and s == abi.SLOT_VBUF:
                continue
            op, nv = int(S["op"]), int(S["num_values"])
            ice = s == abi.SLOT_VBUF and L["geom_column_type"] in (abi.CT_ICE, abi.CT_ICE_MORTON_CODE)
            cnt = 2 * nv if ice else nv
            b = abi.SLOT_BUF[s]
            o = int(L["out"][b])
            cases.append((op, bufs[b][o:o + cnt], int(L["num_bits"])))
            where.append((key, s, int(S["byte_offset"]), int(S["byte_length"]), nv))
    assert len(cases) > 6000
    got = _encode_many(covt, decoder, cases, flags)
    same = {}
    for (op, v, _), enc, (key, s, off, bl, nv) in zip(cases, got, where):
        fam = abi.OP_NAMES[op]
        ok = len(enc) == bl and np.array_equal(enc, blob[off:off + bl])
        tot, good = same.get(fam, (0, 0))
        same[fam] = (tot + 1, good + int(ok))
        if not ok and not (fam.startswith("pfor") and nv > 65536):
            raise AssertionError("%s slot %d op %s (%d values): re-encoding differs (%d vs %d bytes)" % (key, s, fam, nv, len(enc), bl))
    for fam in ("byte_rle", "rle_u32", "rle_u64", "varint_zz_delta", "varint_delta_morton", "pfor_zz_delta"):
        assert same.get(fam, (0, 0))[0] > 0, (fam, same)
    pf = [v for k, v in same.items() if k.startswith("pfor")]
    assert sum(t for t, _ in pf) >= 900 and sum(t - g for t, g in pf) <= 30, same
    res.free()


def test_every_property_stream_of_the_fixtures_reencodes_to_its_own_bytes(covt, decoder, fixtures):
    """The other 26 000+ codec streams of the fixture tiles — present bitsets, BOOLEAN / INT_64 data, dictionary indices and entry
    lengths of the property columns: each one decodes through covt_decode_streams and re-encodes through covt_encode_streams to
    the bytes the reference converter wrote (one call each for all of them). The stream list comes from the gen-2b walker of the
    test infrastructure (oracle/properties.py: names, encodings, offsets); no CPU codec is involved."""
    from oracle import properties as P
    abi = covt.abi
    tiles = [b for _, b in fixtures]
    blob, offs = util.concat_tiles(tiles)
    want = []  # (op, absolute offset, byte length, values to decode)
    for t, tile in enumerate(tiles):
        base = int(offs[t])
        for L in P.walk_gen2b(bytes(tile)):
            for c in L["columns"]:
                if c["data_type"] == P.DT2_GEOMETRY or (c["name"] == "id" and c is L["columns"][0]):
                    continue
                dt = c["data_type"]
                for st in c["streams"]:
                    name, enc, nv = st["name"], st["encoding"], st["num_values"]
                    op, n = None, nv
                    if enc == abi.ENC_BOOLEAN_RLE:
                        op, n = abi.OP_BYTE_RLE, (nv + 7) // 8  # a java.util.BitSet as bytes
                    elif name == "dictionary" or dt in (P.DT2_FLOAT, P.DT2_DOUBLE) or enc == abi.ENC_PLAIN:
                        continue  # raw bytes, no codec
                    elif enc == abi.ENC_RLE:
                        op = abi.OP_RLE_S64 if (dt == P.DT2_INT_64 and name == "data") else abi.OP_RLE_U64
                    elif enc == abi.ENC_VARINT_ZIG_ZAG:
                        op = abi.OP_VARINT_ZZ
                    elif enc == abi.ENC_VARINT_DELTA_ZIG_ZAG:
                        op = abi.OP_VARINT_ZZ_DELTA
                    elif enc == abi.ENC_VARINT:
                        op = abi.OP_VARINT_U32
                    assert op is not None, (c["name"], name, enc)
                    want.append((op, base + st["offset"], st["byte_length"], n))
    assert len(want) > 20000
    descs = (abi.StreamDesc * len(want))()
    for i, (op, off, bl, n) in enumerate(want):
        descs[i] = abi.StreamDesc(byte_offset=off, byte_length=bl, num_values=n, op=op)
    padded = np.concatenate([blob, np.zeros(64, np.uint8)])
    res = decoder.decode_streams(padded, descs)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    cases = []
    for i, (op, off, bl, n) in enumerate(want):
        d = descs[i]
        assert d.status == 0 and d.bytes_consumed == bl, "stream %d op %s: status %d, consumed %d of %d" % (i, abi.OP_NAMES[op], d.status, d.bytes_consumed, bl)
        dt = np.dtype(abi.op_dtype(op))
        cases.append((op, arena[d.out_offset:d.out_offset + d.out_count * dt.itemsize].view(dt).copy(), 0))
    res.free()
    got = _encode_many(covt, decoder, cases)
    seen = {}
    for (op, off, bl, n), enc in zip(want, got):
        assert len(enc) == bl and np.array_equal(enc, blob[off:off + bl]), "op %s, %d values at %d: re-encoding differs (%d vs %d bytes)" % (
            abi.OP_NAMES[op], n, off, len(enc), bl)
        seen[abi.OP_NAMES[op]] = seen.get(abi.OP_NAMES[op], 0) + 1
    assert seen.get("byte_rle", 0) > 9000 and seen.get("rle_u64", 0) > 9000 and seen.get("rle_s64", 0) > 100, seen


def test_encoding_utils_mirror_and_round_trip(covt, decoder):
    """The host-side mirror of EncodingUtils (same names) and DecodingUtils, there and back."""
    E, D = covt.EncodingUtils, covt.DecodingUtils
    rng = np.random.default_rng(15)
    ids = np.cumsum(rng.integers(1, 9, 5000)).astype(np.int64)
    enc = E.encodeRle(ids, False)
    assert np.array_equal(D.decodeRle(np.concatenate([enc, np.zeros(16, np.uint8)]), len(ids), [0], False), ids)
    types = rng.integers(0, 3, 3000).astype(np.uint8)
    enc = E.encodeByteRle(types)
    assert np.array_equal(D.decodeByteRle(np.concatenate([enc, np.zeros(16, np.uint8)]), len(types), [0], len(enc)), types)
    xy = np.cumsum(rng.integers(-300, 300, 20000)).astype(np.int32)
    enc = E.encodeZigZagDeltaCoordinates(xy)
    assert np.array_equal(D.decodeZigZagDeltaVarintCoordinates(np.concatenate([enc, np.zeros(16, np.uint8)]), [0], len(xy)), xy)
    off = np.cumsum(rng.integers(0, 40, 7000)).astype(np.int32)
    enc = E.encodeFastPfor128(off, True, True)
    assert np.array_equal(D.decodeFastPfor128ZigZagDelta(np.concatenate([enc, np.zeros(16, np.uint8)]), len(off), len(enc), [0]), off)


def test_encode_request_validation(covt, decoder):
    abi = covt.abi
    descs = (abi.EncodeDesc * 3)()
    v = np.arange(10, dtype=np.int32)
    descs[0] = abi.EncodeDesc(value_offset=0, num_values=11, op=abi.OP_VARINT_ZZ_DELTA)       # reads past the values
    descs[1] = abi.EncodeDesc(value_offset=0, num_values=4, op=abi.OP_VARINT_ZZ_DELTA_AS_I64)  # a decode-side emulation
    descs[2] = abi.EncodeDesc(value_offset=2, num_values=2, op=abi.OP_VARINT_ZZ)               # misaligned
    res = decoder.encode_streams(v, descs)
    assert [d.status for d in descs] == [abi.ERR_INVALID_ARG, abi.ERR_UNSUPPORTED_ENCODING, abi.ERR_INVALID_ARG]
    assert all(d.byte_length == 0 for d in descs)
    res.free()
