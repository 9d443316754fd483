"""CPU: the oracle against the reference's own known-answer vectors (decodingUtils.spec.ts) and the encoders' inverses."""
import numpy as np

import vectors


def test_known_answer_vectors(oracle):
    abi = oracle.abi
    for v in vectors.VECTORS:
        got, st, cons = oracle.decode_stream(np.asarray(v["bytes"], dtype=np.uint8), getattr(abi, v["op"]),
                                             byte_offset=v.get("offset", 0), num_values=v["n"])
        assert st == 0, v["name"]
        assert list(got) == v["expect"], v["name"]
        assert cons == v["consumed"], v["name"]


def test_bitset_vector():
    b = vectors.BITSET_VECTOR
    isset = lambda i: (b["bytes"][i // 8] >> (i % 8)) & 1  # noqa: E731  BitSet layout (EncodingUtils.java:213-230)
    assert all(isset(i) for i in b["set"]) and not any(isset(i) for i in b["clear"])


def test_java_varint_four_byte_cap(oracle):
    """DecodingUtils.java:157-186: the 4th byte ends the value whatever its continuation bit says."""
    abi = oracle.abi
    got, st, cons = oracle.decode_stream(np.array([0xFF, 0xFF, 0xFF, 0xFF, 0x01], np.uint8), abi.OP_VARINT_U32, num_values=2)
    assert st == abi.ERR_VARINT_OVERLONG and list(got) == [0x0FFFFFFF, 1] and cons == 5


def test_morton_matches_encoder(oracle, gen):
    rng = np.random.default_rng(0)
    for nb in (13, 14):
        ext = 2 << (nb - 2)
        for _ in range(2000):
            x, y = int(rng.integers(-ext // 2, ext + ext // 2 - 1)), int(rng.integers(-ext // 2, ext + ext // 2 - 1))
            assert oracle.decode_morton(gen.encode_morton(x, y, nb), nb) == (x, y)


def test_roundtrips_all_codecs(oracle, gen):
    abi = oracle.abi
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 31, 255, 256, 257, 1000, 65536 + 300):
        v = np.cumsum(rng.integers(-1000, 1000, n)).astype(np.int64)
        for op, enc in ((abi.OP_VARINT_ZZ_DELTA, gen.encode_varints(v, True, True)),
                        (abi.OP_RLE_S64, gen.encode_rle(v, signed=True)),
                        (abi.OP_PFOR_ZZ_DELTA, gen.encode_fastpfor(v.astype(np.int32), True, True)),
                        (abi.OP_VARINT_ZZ_DELTA_64, gen.encode_varints(v << 30, True, True))):
            got, st, cons = oracle.decode_stream(enc, op, num_values=n)
            want = v << 30 if op == abi.OP_VARINT_ZZ_DELTA_64 else v
            assert st == 0 and cons == len(enc) and np.array_equal(got.astype(np.int64), want), (abi.OP_NAMES[op], n)
        b = (rng.integers(0, 3, n) * rng.integers(0, 2, n)).astype(np.uint8)
        got, st, cons = oracle.decode_stream(gen.encode_byte_rle(b), abi.OP_BYTE_RLE, num_values=n)
        assert st == 0 and np.array_equal(got, b)
