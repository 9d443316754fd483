"""GPU parity tests of the stream path (covt_decode_streams = the static codecs of DecodingUtils.java)
against the CPU oracle: every op, lengths straddling the 16/32/256/512/65 536 boundaries, every byte alignment,
the reference's known-answer vectors, and large streams through the multi-CTA two-pass kernels."""
import ctypes as C

import numpy as np
import pytest

import util
import vectors

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 15, 16, 17, 31, 32, 33, 255, 256, 257, 300, 511, 512, 513, 1023, 1024, 4097, 65535, 65536, 65537, 70000, 140000]


def _rand_values(rng, n, kind):
    if kind == "small":
        return rng.integers(0, 100, n)
    if kind == "walk":
        return np.cumsum(rng.integers(-300, 300, n))
    if kind == "wide":
        return rng.integers(-(1 << 26), 1 << 26, n)
    if kind == "runs":
        out = []
        while len(out) < n:
            if rng.random() < 0.5:
                base, d, k = int(rng.integers(0, 1000)), int(rng.integers(-3, 4)), int(rng.integers(3, 200))
                out += [base + d * i for i in range(k)]
            else:
                out += list(rng.integers(0, 1 << 20, int(rng.integers(1, 40))))
        return np.asarray(out[:n])
    raise ValueError(kind)


def _batch_check(covt, oracle, decoder, cases, flags=None):
    """cases: list of (op, payload bytes ndarray, num_values, num_bits, exact). Runs them all in ONE covt_decode_streams
    call (each at a different byte alignment) and compares every output with the oracle."""
    abi = covt.abi
    if flags is None:
        flags = abi.FLAG_DEFAULT
    blob = bytearray()
    descs = (abi.StreamDesc * len(cases))()
    for i, (op, payload, n, nbits, exact) in enumerate(cases):
        blob += bytes((i * 7 + 3) % 16 + (1 if i % 3 == 0 else 0))  # vary the alignment
        off = len(blob)
        blob += bytes(payload)
        slack = 0 if exact else 37  # varint/RLE calls only know where the stream starts (DecodingUtils "pos")
        descs[i] = abi.StreamDesc(byte_offset=off, byte_length=len(payload) + slack, num_values=n, num_bits=nbits, op=op)
        if not exact:
            blob += bytes([0x80] * 3) + bytes(slack - 3)  # trailing garbage that must not be touched
    blob += bytes(64)
    blob = np.frombuffer(bytes(blob), dtype=np.uint8)
    res = decoder.decode_streams(blob, descs, flags)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    for i, (op, payload, n, nbits, exact) in enumerate(cases):
        d = descs[i]
        want, wst, wcons = oracle.decode_stream(blob, op, byte_offset=d.byte_offset, byte_length=d.byte_length, num_values=n,
                                                num_bits=nbits, flags=flags)
        assert (d.status == 0) == (wst == 0), "case %d op %s: status %d vs oracle %d" % (i, abi.OP_NAMES[op], d.status, wst)
        if wst != 0:
            continue
        dt = np.dtype(abi.op_dtype(op))
        got = arena[d.out_offset:d.out_offset + d.out_count * dt.itemsize].view(dt)
        assert d.out_count == len(want), "case %d op %s count %d vs %d" % (i, abi.OP_NAMES[op], d.out_count, len(want))
        if not np.array_equal(got, want):
            bad = np.nonzero(got != want)[0]
            raise AssertionError("case %d op %s n=%d differs at %s: got %s want %s" % (i, abi.OP_NAMES[op], n, bad[:6], got[bad[:6]], want[bad[:6]]))
        assert d.bytes_consumed == wcons, "case %d op %s consumed %d vs %d" % (i, abi.OP_NAMES[op], d.bytes_consumed, wcons)
    res.free()


def test_known_answer_vectors(covt, oracle, decoder):
    """The reference's own vectors (parser/js/test/unit/decoder/decodingUtils.spec.ts) through the GPU path."""
    abi = covt.abi
    for v in vectors.VECTORS:
        op = getattr(abi, v["op"])
        got, st, cons = decoder.decode_stream(np.asarray(v["bytes"], dtype=np.uint8), op, byte_offset=v.get("offset", 0),
                                              num_values=v["n"])
        assert st == 0, v["name"]
        assert list(got) == v["expect"], v["name"]
        assert cons == v["consumed"], v["name"]


def test_varint32_ops(covt, oracle, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(1)
    cases = []
    for n in SIZES:
        for kind in ("small", "walk", "wide"):
            v = _rand_values(rng, n, kind).astype(np.int64)
            cases.append((abi.OP_VARINT_U32, gen.encode_varints(np.abs(v) & 0x0FFFFFFF), n, 0, False))
            cases.append((abi.OP_VARINT_ZZ, gen.encode_varints(v, zigzag=True), n, 0, False))
            cases.append((abi.OP_VARINT_ZZ_DELTA, gen.encode_varints(v, zigzag=True, delta=True), n, 0, False))
            n2 = n & ~1
            xy = _rand_values(rng, n2, kind).astype(np.int32)
            zz = gen.encode_zigzag_delta_coordinates(xy).astype(np.int64) & 0xFFFFFFFF
            cases.append((abi.OP_VARINT_ZZ_DELTA_XY, gen.encode_varints(zz), n2, 0, False))
        codes = np.sort(rng.integers(0, 1 << 26, n)).astype(np.int64)
        for nbits in (13, 14):
            cases.append((abi.OP_VARINT_DELTA_MORTON, gen.encode_varints(codes, delta=True), n, nbits, False))
    _batch_check(covt, oracle, decoder, cases)
    _batch_check(covt, oracle, decoder, [c for c in cases if c[0] == abi.OP_VARINT_DELTA_MORTON][:12], flags=abi.FLAG_MORTON_NO_SHIFT)


def test_varint64_and_id_ops(covt, oracle, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(2)
    cases = []
    for n in SIZES[:20]:
        big = rng.integers(0, 1 << 62, n).astype(np.int64)
        cases.append((abi.OP_VARINT_U64, gen.encode_varints(big), n, 0, False))
        walk = np.cumsum(rng.integers(-(1 << 40), 1 << 40, n)).astype(np.int64)
        cases.append((abi.OP_VARINT_ZZ_DELTA_64, gen.encode_varints(walk, zigzag=True, delta=True), n, 0, False))
        small = rng.integers(0, 1 << 27, n).astype(np.int64)
        cases.append((abi.OP_VARINT_U32_AS_I64, gen.encode_varints(small), n, 0, False))
        cases.append((abi.OP_VARINT_ZZ_DELTA_AS_I64, gen.encode_varints(np.cumsum(rng.integers(-99, 99, n)), zigzag=True, delta=True), n, 0, False))
        cases.append((abi.OP_VARINT_ZZ_AS_I64, gen.encode_varints(rng.integers(-(1 << 20), 1 << 20, n), zigzag=True, delta=False), n, 0, False))
    # extremes
    ext = np.array([0, 1, -1, (1 << 63) - 1, -(1 << 63), 1 << 35, 127, 128], dtype=np.int64)
    cases.append((abi.OP_VARINT_U64, gen.encode_varints(ext), len(ext), 0, False))
    cases.append((abi.OP_VARINT_ZZ_DELTA_64, gen.encode_varints(ext, zigzag=True, delta=True), len(ext), 0, False))
    _batch_check(covt, oracle, decoder, cases)


def test_rle_ops(covt, oracle, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(3)
    cases = []
    for n in SIZES[:21]:
        for kind in ("runs", "small", "wide"):
            v = _rand_values(rng, n, kind).astype(np.int64)
            cases.append((abi.OP_RLE_U32, gen.encode_rle(np.abs(v)), n, 0, False))
            cases.append((abi.OP_RLE_U64, gen.encode_rle(np.abs(v) << 20), n, 0, False))
            cases.append((abi.OP_RLE_S64, gen.encode_rle(v, signed=True), n, 0, False))
        b = _rand_values(rng, n, "runs").astype(np.uint8) % 6
        cases.append((abi.OP_BYTE_RLE, gen.encode_byte_rle(b), n, 0, False))
        cases.append((abi.OP_BYTE_RLE, gen.encode_byte_rle(rng.integers(0, 256, n).astype(np.uint8)), n, 0, True))
    _batch_check(covt, oracle, decoder, cases)


def test_fastpfor_ops(covt, oracle, gen, decoder):
    abi = covt.abi
    rng = np.random.default_rng(4)
    cases = []
    for n in SIZES:
        for kind in ("small", "walk", "wide"):
            v = _rand_values(rng, n, kind).astype(np.int32)
            cases.append((abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(v, zigzag=True, delta=True), n, 0, True))
            n2 = n & ~1
            zz = gen.encode_zigzag_delta_coordinates(v[:n2])
            cases.append((abi.OP_PFOR_ZZ_DELTA_XY, gen.encode_fastpfor(zz), n2, 0, True))
        # exception-heavy: mostly small values with outliers of many widths
        v = rng.integers(0, 8, n).astype(np.int64)
        idx = rng.random(n) < 0.12
        v[idx] = rng.integers(0, 1 << 30, int(idx.sum())) >> rng.integers(0, 29, int(idx.sum()))
        cases.append((abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(np.cumsum(v).astype(np.int32), zigzag=True, delta=True), n, 0, True))
        codes = np.sort(rng.integers(0, 1 << 27, n)).astype(np.int32)
        cases.append((abi.OP_PFOR_DELTA_MORTON, gen.encode_fastpfor(codes, delta=True), n, 13, True))
        cases.append((abi.OP_PFOR_DELTA_MORTON, gen.encode_fastpfor(codes, delta=True), n, 14, True))
    # full 32-bit values (b = 32) and all-zero blocks (b = 0)
    cases.append((abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(rng.integers(-(1 << 31), 1 << 31, 1024).astype(np.int32), zigzag=True), 1024, 0, True))
    cases.append((abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(np.zeros(700, np.int32), zigzag=True, delta=True), 700, 0, True))
    _batch_check(covt, oracle, decoder, cases)


def test_malformed_streams(covt, oracle, gen, decoder):
    """Truncated / overlong / miscounted streams: the GPU flags exactly what the oracle flags."""
    abi = covt.abi
    rng = np.random.default_rng(6)
    v = np.cumsum(rng.integers(-500, 500, 2000)).astype(np.int64)
    enc = gen.encode_varints(v, zigzag=True, delta=True)
    cases = [
        (abi.OP_VARINT_ZZ_DELTA, enc[:1000], 2000, 0, True),                                   # truncated
        (abi.OP_VARINT_ZZ_DELTA, np.array([0x80, 0x80, 0x80, 0x80, 0x01] * 10, np.uint8), 10, 0, True),  # 5-byte varints
        (abi.OP_VARINT_ZZ_DELTA_XY, enc, 1999, 0, True),                                       # odd coordinate count
        (abi.OP_RLE_U32, gen.encode_rle(np.arange(500))[:-1], 500, 0, True),
        (abi.OP_BYTE_RLE, gen.encode_byte_rle(np.arange(300).astype(np.uint8))[:-5], 300, 0, True),
        (abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(v.astype(np.int32), True, True), 2100, 0, True),  # numValues too large
        (abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(v.astype(np.int32), True, True), 1900, 0, True),  # numValues too small
        (abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(v.astype(np.int32), True, True)[:2000], 2000, 0, True),
        (abi.OP_PFOR_ZZ_DELTA, rng.integers(0, 256, 3000).astype(np.uint8), 1000, 0, True),      # garbage
        (abi.OP_VARINT_ZZ_DELTA, enc, 2000, 0, True),                                          # and a good one
    ]
    _batch_check(covt, oracle, decoder, cases)


def test_overlong_varints_follow_the_java_reader(covt, oracle, gen, decoder):
    """A value with four continuation bytes: the Java reader ends it after the 4th byte (DecodingUtils.java:157-186), so it counts
    more values than the stream has terminators. Status (OVERLONG vs TRUNCATED) and bytes consumed equal the oracle's exactly,
    in the warp-per-stream decoder and in the large-stream kernels (found by the property-column fuzzer, seed 307)."""
    abi = covt.abi
    five = [0x80, 0x80, 0x80, 0x80, 0x01]
    small = [
        ([0xCE, 0xB9, 0xA9, 0xB8, 0x01], 2), ([0xCE, 0xB9, 0xA9, 0xB8], 2), ([0xCE, 0xB9, 0xA9, 0xB8], 1), (five * 10, 10), (five * 10, 20),
        (five * 10, 21), ([1, 2, 3] + five + [4, 5], 6), ([1, 2, 3] + five + [4, 5], 7), ([1, 2, 3] + five + [4, 5], 8),
        ([0x80] * 9 + [1, 2], 3), ([0x80] * 9 + [1, 2], 4), ([0x80] * 9 + [1, 2], 5), ([0x80] * 700 + [1], 176), ([0x80] * 700 + [1], 177),
    ]
    enc, vals = gen.varint_stream((1 << 18) + 4096, seed=5)
    enc = np.asarray(enc, dtype=np.uint8)
    for at in (100, 200_001):  # an overlong run early / late in a large stream (decoded over many CTAs)
        while enc[at - 1] & 0x80:
            at += 1
        big = np.concatenate([enc[:at], np.array([0x81] * 6 + [0x01], np.uint8), enc[at:]])
        for n in (vals, vals + 1, vals + 2, vals + 3, vals // 2):
            small.append((big, n))
    flagged = 0
    for op in (abi.OP_VARINT_ZZ_DELTA, abi.OP_VARINT_U32, abi.OP_VARINT_ZZ_DELTA_AS_I64):
        for payload, n in small:
            payload = np.asarray(payload, dtype=np.uint8)
            if len(payload) > 4096 and op != abi.OP_VARINT_ZZ_DELTA:
                continue
            blob = np.concatenate([np.zeros(5, np.uint8), payload, np.zeros(64, np.uint8)])
            got, st, cons = decoder.decode_stream(blob, op, byte_offset=5, byte_length=len(payload), num_values=n)
            want, wst, wcons = oracle.decode_stream(blob, op, byte_offset=5, byte_length=len(payload), num_values=n)
            assert (st, cons) == (wst, wcons), "op %s, %d bytes, n=%d: status/consumed %s vs oracle %s" % (
                abi.OP_NAMES[op], len(payload), n, (st, cons), (wst, wcons))
            flagged += wst in (abi.ERR_VARINT_OVERLONG, abi.ERR_TRUNCATED)
    assert flagged >= 45


@pytest.mark.parametrize("post", ["OP_VARINT_ZZ_DELTA_XY", "OP_VARINT_ZZ_DELTA", "OP_VARINT_DELTA_MORTON", "OP_VARINT_ZZ", "OP_VARINT_U32"])
def test_large_varint_streams_lookback_kernel(covt, oracle, gen, decoder, post):
    """Streams >= 256 KiB take the multi-CTA decoupled-look-back kernel; several streams per launch, any alignment."""
    abi = covt.abi
    op = getattr(abi, post)
    cases = []
    for k, target in enumerate([1 << 18, (1 << 20) + 13, 3 * (1 << 20) + 4095]):
        if op == abi.OP_VARINT_DELTA_MORTON:
            rng = np.random.default_rng(k)
            n = target // 2
            enc = gen.encode_varints(np.cumsum(rng.integers(0, 40000, n)).astype(np.int64) & 0x3FFFFFF, delta=False)
            vals = n
        else:
            enc, vals = gen.varint_stream(target & ~1 if op == abi.OP_VARINT_ZZ_DELTA_XY else target, seed=0xC0717 + k)
        cases.append((op, enc, vals, 14, True))
    _batch_check(covt, oracle, decoder, cases)


def test_config3_shape_roundtrip(covt, gen, decoder):
    """Config 3 at 64 MiB: decode -> re-encode the coordinates -> identical bytes (size-independent round trip)."""
    abi = covt.abi
    enc, n = gen.varint_stream(64 << 20, seed=0xC0717)
    got, st, cons = decoder.decode_stream(enc, abi.OP_VARINT_ZZ_DELTA_XY, num_values=n, byte_length=len(enc))
    assert st == 0 and len(got) == n
    zz = gen.encode_zigzag_delta_coordinates(got).astype(np.int64) & 0xFFFFFFFF
    assert np.array_equal(gen.encode_varints(zz), enc)


def test_config4_index_buffers_of_fixture_polygons(covt, oracle, gen, decoder, fixtures):
    """BASELINE config 4: the polygon layers of the zoom 5-8 fixtures with a synthetic INDEX_BUFFER per layer — fan triangulation
    (v0, v0+i, v0+i+1) of every ring, FAST_PFOR_DELTA_ZIG_ZAG, stream type 12 (an EXTENSION: the reference defines IndexBuffer in
    README prose only, parity unpinned beyond the oracle) — decoded by the FastPFOR routine through the descriptor dispatch."""
    abi = covt.abi
    tiles = [(n, b) for n, b in fixtures if n.startswith(("omt/5_", "omt/6_", "omt/7_", "omt/8_"))]
    assert len(tiles) >= 8
    blob, offs = util.concat_tiles([b for _, b in tiles])
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_ID_DVZZ_IS_RLE | abi.FLAG_MORTON_NO_SHIFT)
    payload = bytearray()
    wanted = []
    for L in ref.layers:
        if L["status"] != 0 or L["n_rings"] == 0:
            continue
        types = ref.buffer(abi.BUF_S_GEOMETRY_TYPES)[int(L["out"][abi.BUF_S_GEOMETRY_TYPES]):][:int(L["num_features"])]
        if not np.isin(types, (abi.GT_POLYGON, abi.GT_MULTIPOLYGON)).all():
            continue
        ring = ref.buffer(abi.BUF_A_RING_OFFSETS)[int(L["out"][abi.BUF_A_RING_OFFSETS]):][:int(L["n_rings"]) + 1].astype(np.int64)
        idx = []
        for r in range(len(ring) - 1):
            v0, n = ring[r], ring[r + 1] - ring[r]
            if n >= 3:
                i = np.arange(1, n - 1)
                idx.append(np.stack([np.full_like(i, v0), v0 + i, v0 + i + 1], axis=1).ravel())
        if not idx:
            continue
        idx = np.concatenate(idx).astype(np.int32)
        enc = gen.encode_fastpfor(idx, zigzag=True, delta=True)
        wanted.append((len(payload), len(enc), idx))
        payload += bytes(enc)
    assert len(wanted) >= 8 and sum(len(w[2]) for w in wanted) > 100000
    payload = np.frombuffer(bytes(payload) + bytes(64), dtype=np.uint8)
    descs = (abi.StreamDesc * len(wanted))()
    for i, (off, ln, idx) in enumerate(wanted):
        descs[i] = abi.StreamDesc(byte_offset=off, byte_length=ln, num_values=len(idx), stream_type=abi.ST_INDEX_BUFFER,
                                  encoding=abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, column_type=abi.CT_PLAIN)
    res = decoder.decode_streams(payload, descs, abi.FLAG_DEFAULT)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    for i, (off, ln, idx) in enumerate(wanted):
        d = descs[i]
        assert d.status == 0 and d.out_count == len(idx) and d.bytes_consumed == ln
        got = arena[d.out_offset:d.out_offset + 4 * d.out_count].view(np.int32)
        assert np.array_equal(got, idx), i
        want, wst, _ = oracle.decode_stream(payload, 0, byte_offset=off, byte_length=ln, num_values=len(idx),
                                            stream_type=abi.ST_INDEX_BUFFER, encoding=abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG)
        assert wst == 0 and np.array_equal(want, idx)
    res.free()


def _fuzz_cases(abi, gen, seed):
    """Valid streams of every op (small: lane decoders; medium: warp decoders), each mutated four ways."""
    rng = np.random.default_rng(seed)
    cases = []

    def base_streams(op, n):
        if op == abi.OP_BYTE_RLE:
            v = np.repeat(rng.integers(0, 6, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n].astype(np.uint8)
            return gen.encode_byte_rle(v), len(v), 0
        if op in (abi.OP_RLE_U32, abi.OP_RLE_U64, abi.OP_RLE_S64):
            v = np.repeat(rng.integers(0, 1 << 20, n // 2 + 1), rng.integers(1, 5, n // 2 + 1))[:n].astype(np.int64)
            if op == abi.OP_RLE_U64:
                v = v << 30
            if op == abi.OP_RLE_S64:
                v = v - (1 << 19)
            return gen.encode_rle(v, signed=(op == abi.OP_RLE_S64)), len(v), 0
        if op in (abi.OP_VARINT_U32, abi.OP_VARINT_ZZ, abi.OP_VARINT_ZZ_DELTA, abi.OP_VARINT_ZZ_DELTA_XY):
            n2 = n + (n & 1)
            v = rng.integers(-3000, 3000, n2).astype(np.int64)
            if op == abi.OP_VARINT_U32:
                return gen.encode_varints(np.abs(v) * 37), n2, 0
            if op == abi.OP_VARINT_ZZ:
                return gen.encode_varints(v, zigzag=True), n2, 0
            if op == abi.OP_VARINT_ZZ_DELTA:
                return gen.encode_varints(np.cumsum(v), zigzag=True, delta=True), n2, 0
            zz = gen.encode_zigzag_delta_coordinates(np.cumsum(v.reshape(-1, 2), axis=0).astype(np.int32).ravel()).astype(np.int64) & 0xFFFFFFFF
            return gen.encode_varints(zz), n2, 0
        if op == abi.OP_VARINT_DELTA_MORTON:
            return gen.encode_varints(rng.integers(0, 40000, n).astype(np.int64)), n, 13
        if op in (abi.OP_VARINT_U64, abi.OP_VARINT_ZZ_DELTA_64):
            v = np.cumsum(rng.integers(-(1 << 40), 1 << 40, n).astype(np.int64))
            if op == abi.OP_VARINT_U64:
                return gen.encode_varints(np.abs(v)), n, 0
            return gen.encode_varints(v, zigzag=True, delta=True), n, 0
        if op in (abi.OP_PFOR_ZZ_DELTA, abi.OP_PFOR_ZZ_DELTA_XY):
            n2 = n + (n & 1)
            v = np.cumsum(rng.integers(-300, 300, n2)).astype(np.int32)
            v[rng.integers(0, n2, max(1, n2 // 40))] += 1 << 20  # exceptions
            if op == abi.OP_PFOR_ZZ_DELTA:
                return gen.encode_fastpfor(v, zigzag=True, delta=True), n2, 0
            return gen.encode_fastpfor(gen.encode_zigzag_delta_coordinates(v), zigzag=False, delta=False), n2, 0
        if op == abi.OP_PFOR_DELTA_MORTON:
            d = rng.integers(0, 3000, n).astype(np.int32)
            return gen.encode_fastpfor(d, zigzag=False, delta=False), n, 13
        raise AssertionError(op)

    ops = [abi.OP_BYTE_RLE, abi.OP_RLE_U32, abi.OP_RLE_U64, abi.OP_RLE_S64, abi.OP_VARINT_U32, abi.OP_VARINT_ZZ, abi.OP_VARINT_ZZ_DELTA,
           abi.OP_VARINT_ZZ_DELTA_XY, abi.OP_VARINT_DELTA_MORTON, abi.OP_VARINT_U64, abi.OP_VARINT_ZZ_DELTA_64, abi.OP_PFOR_ZZ_DELTA,
           abi.OP_PFOR_ZZ_DELTA_XY, abi.OP_PFOR_DELTA_MORTON]
    for op in ops:
        sizes = [int(x) for x in rng.integers(1, 300, 8)] + [int(x) for x in rng.integers(3000, 20000, 3)]
        for n in sizes:
            enc, nv, nbits = base_streams(op, n)
            enc = np.asarray(enc, dtype=np.uint8)
            exact = op in (abi.OP_PFOR_ZZ_DELTA, abi.OP_PFOR_ZZ_DELTA_XY, abi.OP_PFOR_DELTA_MORTON) or bool(rng.integers(0, 2))
            cases.append((op, enc, nv, nbits, exact))  # the unmutated stream
            for kind in range(4):
                b = bytearray(enc.tobytes())
                if kind == 0:
                    b[int(rng.integers(0, len(b)))] ^= int(rng.integers(1, 256))
                elif kind == 1:
                    for _ in range(3):
                        b[int(rng.integers(0, len(b)))] ^= int(rng.integers(1, 256))
                elif kind == 2:
                    b = b[: int(rng.integers(0, len(b)))]
                elif len(b) > 1:
                    del b[int(rng.integers(0, len(b)))]
                if len(b) == 0:
                    continue
                cases.append((op, np.frombuffer(bytes(b), dtype=np.uint8), nv, nbits, exact))
    return cases


@pytest.mark.parametrize("seed", [11, 12])
def test_stream_mutation_fuzz_against_oracle(covt, oracle, gen, decoder, seed):
    """~750 streams per seed, every op, lane- and warp-decoded sizes, each also flipped / truncated / shortened: one
    covt_decode_streams call; status OK-ness and every accepted value equal the oracle's (DecodingUtils semantics)."""
    cases = _fuzz_cases(covt.abi, gen, seed)
    assert len(cases) > 700
    _batch_check(covt, oracle, decoder, cases)


def test_property_streams_of_the_fixtures(covt, oracle, decoder, fixtures):
    """SURVEY §8 f1, first GPU step: every listed stream of every property column of the 129 fixture tiles (present bitsets,
    INT_64 data, dictionary indices and lengths: ~28 000 streams) through covt_decode_streams in ONE call — the codecs are the
    ones of the geometry path (Byte-RLE, RLE signed / unsigned, zigzag varints). Values and consumed bytes equal the oracle's
    (which tests/test_oracle_properties.py pins on the MVT property values). FLOAT data and dictionary bytes are plain copies
    and need no kernel."""
    from oracle import properties as P
    abi = covt.abi
    blob = bytearray()
    wanted = []  # (op, offset, byte_length, num_values)
    for name, data in fixtures:
        base = len(blob)
        blob += data
        for L in P.walk_gen2b(bytes(data)):
            F = L["num_features"]
            for c in L["columns"]:
                if c["data_type"] == P.DT2_GEOMETRY or (c["name"] == "id" and c is L["columns"][0]):
                    continue
                for s in c["streams"]:
                    nm, enc, nv = s["name"], s["encoding"], s["num_values"]
                    if s["byte_length"] == 0:
                        continue
                    if nm.startswith("present"):
                        op, n = abi.OP_BYTE_RLE, (F + 7) // 8
                    elif c["data_type"] == P.DT2_BOOLEAN:
                        op, n = abi.OP_BYTE_RLE, (nv + 7) // 8
                    elif c["data_type"] in (P.DT2_FLOAT, P.DT2_DOUBLE) or nm == "dictionary":
                        continue
                    elif c["data_type"] in (P.DT2_INT_64, P.DT2_UINT_64):
                        signed = c["data_type"] == P.DT2_INT_64
                        op = {abi.ENC_RLE: abi.OP_RLE_S64 if signed else abi.OP_RLE_U64, abi.ENC_VARINT_ZIG_ZAG: abi.OP_VARINT_ZZ,
                              abi.ENC_VARINT_DELTA_ZIG_ZAG: abi.OP_VARINT_ZZ_DELTA, abi.ENC_VARINT: abi.OP_VARINT_U32}.get(enc)
                        if op is None:
                            continue
                        n = nv
                    else:  # dictionary indices (`data` / localized sub-keys) and `length`: unsigned RLE
                        op, n = abi.OP_RLE_U64, nv
                    wanted.append((op, base + s["offset"], s["byte_length"], n))
    assert len(wanted) > 25000
    blob = np.frombuffer(bytes(blob) + bytes(64), dtype=np.uint8)
    descs = (abi.StreamDesc * len(wanted))()
    for i, (op, off, bl, n) in enumerate(wanted):
        descs[i] = abi.StreamDesc(byte_offset=off, byte_length=bl, num_values=n, op=op)
    res = decoder.decode_streams(blob, descs, abi.FLAG_DEFAULT)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    seen = {}
    for i, (op, off, bl, n) in enumerate(wanted):
        d = descs[i]
        want, wst, wcons = oracle.decode_stream(blob, op, byte_offset=off, byte_length=bl, num_values=n)
        assert wst == 0 and wcons == bl, (i, abi.OP_NAMES[op], wst, wcons, bl)
        assert d.status == 0 and d.bytes_consumed == bl and d.out_count == len(want), (i, abi.OP_NAMES[op], d.status, d.bytes_consumed, bl)
        dt = np.dtype(abi.op_dtype(op))
        got = arena[d.out_offset:d.out_offset + d.out_count * dt.itemsize].view(dt)
        assert np.array_equal(got, want), (i, abi.OP_NAMES[op])
        seen[abi.OP_NAMES[op]] = seen.get(abi.OP_NAMES[op], 0) + 1
    assert seen.get("byte_rle", 0) > 10000 and seen.get("rle_u64", 0) > 10000 and seen.get("rle_s64", 0) > 500
    res.free()


def test_config2_rle_topology_streams_in_isolation(covt, oracle, decoder, fixtures):
    """the RLE branch on its own — the RLE-encoded topology streams of the 91 OMT fixtures (geometry_offsets,
    part_offsets, ring_offsets: the 1 123 streams of SURVEY §4.4) through covt_decode_streams in one call, by descriptor
    (stream type + encoding + column type -> op, the dispatch table of CovtParser.decodeGeometryColumn)."""
    from tools.gen import rewrite
    abi = covt.abi
    st_of = {"geometry_offsets": abi.ST_GEOMETRY_OFFSETS, "part_offsets": abi.ST_PART_OFFSETS, "ring_offsets": abi.ST_RING_OFFSETS}
    blob = bytearray()
    wanted = []
    for name, data in fixtures:
        if not name.startswith("omt/"):
            continue
        base = len(blob)
        blob += data
        for L in rewrite.walk(data)[1]:
            for c in L["columns"]:
                if c["data_type"] != rewrite.DT2_GEOMETRY:
                    continue
                for s in c["streams"]:
                    if s["name"] in st_of and s["encoding"] == abi.ENC_RLE:
                        wanted.append((st_of[s["name"]], c["column_type"], base + s["offset"], s["byte_length"], s["num_values"]))
    assert len(wanted) >= 1000
    blob = np.frombuffer(bytes(blob) + bytes(64), dtype=np.uint8)
    descs = (abi.StreamDesc * len(wanted))()
    for i, (stype, ct, off, bl, nv) in enumerate(wanted):
        descs[i] = abi.StreamDesc(byte_offset=off, byte_length=bl, num_values=nv, stream_type=stype, encoding=abi.ENC_RLE, column_type=ct)
    res = decoder.decode_streams(blob, descs, abi.FLAG_DEFAULT)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    total = 0
    for i, (stype, ct, off, bl, nv) in enumerate(wanted):
        d = descs[i]
        want, wst, wcons = oracle.decode_stream(blob, abi.OP_RLE_U32, byte_offset=off, byte_length=bl, num_values=nv)
        assert wst == 0 and wcons == bl
        assert d.status == 0 and d.bytes_consumed == bl and d.out_count == nv, (i, d.status, d.bytes_consumed, bl)
        assert np.array_equal(arena[d.out_offset:d.out_offset + 4 * nv].view(np.int32), want), i
        total += nv
    assert len(wanted) == 1123 and total > 80000  # the census of SURVEY §4.4
    res.free()
