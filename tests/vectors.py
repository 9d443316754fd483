"""Known-answer vectors of the reference (parser/js/test/unit/decoder/decodingUtils.spec.ts:10-113), restated as data.
The 7-byte varint exceeds the Java decoder's 4-byte cap and therefore pins the 64-bit id path only."""


def _zz(i):
    return ((i >> 31) ^ (i << 1)) & 0xFF


VECTORS = [
    # decodeVarint :11-52
    {"name": "varint 1 byte", "op": "OP_VARINT_U32", "bytes": [10], "n": 1, "expect": [10], "consumed": 1},
    {"name": "varint 4 bytes", "op": "OP_VARINT_U32", "bytes": [0x80, 0x80, 0x80, 4], "n": 1, "expect": [8388608], "consumed": 4},
    {"name": "varint 4 bytes at offset 2", "op": "OP_VARINT_U32", "bytes": [0x80, 0x80, 0x80, 0x80, 0x80, 4], "offset": 2, "n": 1,
     "expect": [8388608], "consumed": 4},
    {"name": "varint 7 bytes at offset 2 (64-bit path)", "op": "OP_VARINT_U64", "bytes": [0x80] * 8 + [4], "offset": 2, "n": 1,
     "expect": [17592186044416], "consumed": 7},
    # decodeZigZagVarint :57-66
    {"name": "zigzag varint", "op": "OP_VARINT_ZZ", "bytes": [155, 4], "n": 1, "expect": [-270], "consumed": 2},
    # decodeRle: runs :70-78
    {"name": "rle runs", "op": "OP_RLE_U32", "bytes": [2, 1, 1, 2, 1, 1], "n": 10, "expect": [1, 2, 3, 4, 5, 1, 2, 3, 4, 5], "consumed": 6},
    # decodeRle: literals and runs in combination, signed :80-103 (expected consumed = 12). The spec zigzag-encodes the
    # delta byte of the second run (-1 -> 0x01) although the wire format stores it as a plain signed byte, so its own
    # decoder yields 50..99 ascending, not the 50..1 it asserts (the test carries a "TODO: check why failing?"); the
    # bytes and the consumed count are what pins the format here.
    {"name": "rle runs + literals signed", "op": "OP_RLE_S64",
     "bytes": [0x61, 0x00, 0x0E] + [0xFB] + [_zz(i) for i in (2, 3, 6, 7, 11)] + [0x2F] + [_zz(-1), _zz(0x32)],
     "n": 155, "expect": [7] * 100 + [2, 3, 6, 7, 11] + list(range(50, 100)), "consumed": 12},
    # the same run with the delta stored the ORC way (signed byte 0xFF = -1): 50 down to 1
    {"name": "rle descending run signed", "op": "OP_RLE_S64", "bytes": [0x2F, 0xFF, _zz(0x32)], "n": 50,
     "expect": list(range(50, 0, -1)), "consumed": 3},
]

# isBitSet([0, 2], 9) :107-112 — BitSet layout used by boolean/present streams (bit i of byte i/8, LSB first)
BITSET_VECTOR = {"bytes": [0, 2], "set": [9], "clear": [8]}
