"""ctypes mirror of include/covt_b200.h (enums, POD structs). Interface definitions only — no compute.

Enum values are the ordinals of the reference Java enums (J/converter/StreamEncoding.java:3-16,
StreamType.java:3-16, ColumnType.java:3-9, ColumnDataType.java:3-21; GeometryType CovtParser.java:20-27).
"""
import ctypes as C

import numpy as np

ABI_VERSION = 5

# covt_stream_encoding
ENC_PLAIN, ENC_VARINT, ENC_VARINT_ZIG_ZAG, ENC_VARINT_DELTA, ENC_VARINT_DELTA_ZIG_ZAG = 0, 1, 2, 3, 4
ENC_RLE, ENC_BOOLEAN_RLE, ENC_BYTE_RLE, ENC_FAST_PFOR_DELTA, ENC_FAST_PFOR_DELTA_ZIG_ZAG = 5, 6, 7, 8, 9
ENC_ABSENT = 0xFF
# covt_stream_type
(ST_PRESENT, ST_DATA, ST_LENGTH, ST_DICTIONARY, ST_GEOMETRY_TYPES, ST_GEOMETRY_OFFSETS, ST_PART_OFFSETS,
 ST_RING_OFFSETS, ST_VERTEX_OFFSETS, ST_VERTEX_BUFFER, ST_Z_VALUE, ST_M_VALUE, ST_INDEX_BUFFER) = range(13)
# covt_column_type
CT_PLAIN, CT_DICTIONARY, CT_LOCALIZED_DICTIONARY, CT_ICE, CT_ICE_MORTON_CODE = range(5)
# covt_column_data_type (gen-3)
DT_BOOLEAN, DT_INT_32, DT_UINT_32, DT_INT_64, DT_UINT_64, DT_FLOAT, DT_DOUBLE, DT_STRING, DT_GEOMETRY = range(9)
# covt_geometry_type
GT_POINT, GT_LINESTRING, GT_POLYGON, GT_MULTIPOINT, GT_MULTILINESTRING, GT_MULTIPOLYGON = range(6)
# covt_container
CONTAINER_GEN2B, CONTAINER_GEN3 = 0, 1
# covt_status
(OK, ERR_INVALID_ARG, ERR_CUDA, ERR_OOM, ERR_TRUNCATED, ERR_BAD_METADATA, ERR_UNSUPPORTED_ENCODING,
 ERR_UNSUPPORTED_GEOMETRY, ERR_VARINT_OVERLONG, ERR_COUNT_MISMATCH, ERR_TOPOLOGY) = range(11)
STATUS_NAMES = ["OK", "INVALID_ARG", "CUDA", "OOM", "TRUNCATED", "BAD_METADATA", "UNSUPPORTED_ENCODING",
                "UNSUPPORTED_GEOMETRY", "VARINT_OVERLONG", "COUNT_MISMATCH", "TOPOLOGY"]
# flags
FLAG_CLOSE_RINGS = 0x0001
FLAG_ID_DVZZ_IS_RLE = 0x0002
FLAG_MORTON_NO_SHIFT = 0x0004
FLAG_ID_WIDTH_32 = 0x0008
FLAG_ICE_VB_COUNT_IS_INTS = 0x0010
FLAG_SKIP_ASSEMBLY = 0x0020
FLAG_PROFILE_KERNELS = 0x0040
FLAG_DECODE_PROPERTIES = 0x0080
FLAG_DEFAULT = FLAG_CLOSE_RINGS
# slots
(SLOT_ID, SLOT_TYPES, SLOT_GEOM, SLOT_PART, SLOT_RING, SLOT_VOFF, SLOT_VBUF, SLOT_INDEX) = range(8)
NUM_SLOTS = 8
SLOT_NAMES = ["id", "geometry_types", "geometry_offsets", "part_offsets", "ring_offsets", "vertex_offsets",
              "vertex_buffer", "index_buffer"]
# buffers
(BUF_S_GEOMETRY_TYPES, BUF_S_IDS, BUF_S_GEOMETRY_OFFSETS, BUF_S_PART_OFFSETS, BUF_S_RING_OFFSETS,
 BUF_S_VERTEX_OFFSETS, BUF_S_VERTEX_BUFFER, BUF_S_INDEX_BUFFER, BUF_A_GEOM_OFFSETS, BUF_A_PART_OFFSETS,
 BUF_A_RING_OFFSETS, BUF_A_COORDS, BUF_STREAM_ARENA) = range(13)
NUM_BUFFERS = 13
BUF_NAMES = ["s_geometry_types", "s_ids", "s_geometry_offsets", "s_part_offsets", "s_ring_offsets",
             "s_vertex_offsets", "s_vertex_buffer", "s_index_buffer", "a_geom_offsets", "a_part_offsets",
             "a_ring_offsets", "a_coords", "stream_arena"]
BUF_DTYPES = [np.uint8, np.int64, np.int32, np.int32, np.int32, np.int32, np.int32, np.int32, np.int32,
              np.int32, np.int32, np.int32, np.uint8]
SLOT_BUF = [BUF_S_IDS, BUF_S_GEOMETRY_TYPES, BUF_S_GEOMETRY_OFFSETS, BUF_S_PART_OFFSETS, BUF_S_RING_OFFSETS,
            BUF_S_VERTEX_OFFSETS, BUF_S_VERTEX_BUFFER, BUF_S_INDEX_BUFFER]
# ops
(OP_NONE, OP_BYTE_RLE, OP_RLE_U32, OP_RLE_U64, OP_RLE_S64, OP_VARINT_U32, OP_VARINT_ZZ, OP_VARINT_ZZ_DELTA,
 OP_VARINT_ZZ_DELTA_XY, OP_VARINT_DELTA_MORTON, OP_VARINT_U64, OP_VARINT_ZZ_DELTA_64, OP_PFOR_ZZ_DELTA,
 OP_PFOR_ZZ_DELTA_XY, OP_PFOR_DELTA_MORTON, OP_VARINT_U32_AS_I64, OP_VARINT_ZZ_DELTA_AS_I64, OP_VARINT_ZZ_AS_I64) = range(18)
OP_NAMES = ["none", "byte_rle", "rle_u32", "rle_u64", "rle_s64", "varint_u32", "varint_zz", "varint_zz_delta",
            "varint_zz_delta_xy", "varint_delta_morton", "varint_u64", "varint_zz_delta_64", "pfor_zz_delta",
            "pfor_zz_delta_xy", "pfor_delta_morton", "varint_u32_as_i64", "varint_zz_delta_as_i64", "varint_zz_as_i64"]


def op_elem_size(op):
    if op == OP_BYTE_RLE:
        return 1
    if op in (OP_RLE_U64, OP_RLE_S64, OP_VARINT_U64, OP_VARINT_ZZ_DELTA_64, OP_VARINT_U32_AS_I64,
              OP_VARINT_ZZ_DELTA_AS_I64, OP_VARINT_ZZ_AS_I64):
        return 8
    return 4


def op_dtype(op):
    return {1: np.uint8, 8: np.int64, 4: np.int32}[op_elem_size(op)]


class StreamRef(C.Structure):
    _fields_ = [("byte_offset", C.c_uint64), ("byte_length", C.c_uint32), ("num_values", C.c_uint32),
                ("encoding", C.c_uint8), ("op", C.c_uint8), ("reserved", C.c_uint8 * 2), ("status", C.c_uint32)]


class Layer(C.Structure):
    _fields_ = [("tile", C.c_uint32), ("layer_index", C.c_uint32), ("extent", C.c_uint32),
                ("num_features", C.c_uint32), ("num_columns", C.c_uint32), ("status", C.c_uint32),
                ("geom_column_type", C.c_uint8), ("num_bits", C.c_uint8), ("has_id", C.c_uint8),
                ("reserved", C.c_uint8), ("name_length", C.c_uint32), ("name_offset", C.c_uint64),
                ("streams", StreamRef * NUM_SLOTS), ("out", C.c_uint64 * NUM_BUFFERS),
                ("n_parts", C.c_uint32), ("n_rings", C.c_uint32), ("n_vertices", C.c_uint32),
                ("n_coords", C.c_uint32), ("cap_parts", C.c_uint32), ("cap_rings", C.c_uint32),
                ("header_offset", C.c_uint64)]


class StreamDesc(C.Structure):
    _fields_ = [("byte_offset", C.c_uint64), ("byte_length", C.c_uint32), ("num_values", C.c_uint32),
                ("stream_type", C.c_uint8), ("encoding", C.c_uint8), ("column_type", C.c_uint8),
                ("column_data_type", C.c_uint8), ("num_bits", C.c_uint8), ("op", C.c_uint8),
                ("reserved", C.c_uint8 * 2), ("status", C.c_uint32), ("bytes_consumed", C.c_uint32),
                ("out_offset", C.c_uint64), ("out_count", C.c_uint64)]


class EncodeDesc(C.Structure):
    """covt_encode_desc: one EncodingUtils call (value_offset in BYTES into the values buffer)."""
    _fields_ = [("value_offset", C.c_uint64), ("num_values", C.c_uint32), ("op", C.c_uint8), ("num_bits", C.c_uint8),
                ("reserved", C.c_uint8 * 2), ("out_offset", C.c_uint64), ("byte_length", C.c_uint32), ("status", C.c_uint32)]


class HostSink(C.Structure):
    """covt_host_sink: page-locked host destinations (and capacities in elements) per result buffer; NULL = not wanted."""
    _fields_ = [("ptr", C.c_void_p * 13), ("capacity", C.c_uint64 * 13)]


class TileJson(C.Structure):
    _fields_ = [("n_vector_layers", C.c_uint32), ("n_fields", C.POINTER(C.c_uint32))]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("decode_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("payload_bytes", C.c_uint64), ("output_bytes", C.c_uint64),
                ("vertices", C.c_uint64), ("segments", C.c_uint32), ("capacity_retries", C.c_uint32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("launches", C.c_uint32),
                ("algorithmic_bytes", C.c_uint64)]


# property columns (COVT_FLAG_DECODE_PROPERTIES)
PV_NONE, PV_I64, PV_F32, PV_F64, PV_BOOL, PV_DICT_INDEX = range(6)
(PBUF_VALIDITY, PBUF_I64, PBUF_F32, PBUF_F64, PBUF_BOOL, PBUF_DICT_INDEX, PBUF_DICT_OFFSETS) = range(7)
NUM_PROP_BUFFERS = 7
PBUF_DTYPES = [np.uint8, np.int64, np.float32, np.float64, np.uint8, np.int32, np.int32]


class PropColumn(C.Structure):
    _fields_ = [("tile", C.c_uint32), ("layer", C.c_uint32), ("name_offset", C.c_uint64), ("sub_offset", C.c_uint64),
                ("name_length", C.c_uint32), ("sub_length", C.c_uint32), ("data_type", C.c_uint8), ("column_type", C.c_uint8),
                ("value_kind", C.c_uint8), ("reserved", C.c_uint8), ("status", C.c_uint32), ("num_features", C.c_uint32),
                ("num_values", C.c_uint32), ("validity_offset", C.c_uint64), ("values_offset", C.c_uint64),
                ("dictionary", C.c_uint32), ("data_num_values", C.c_uint32)]


class PropDictionary(C.Structure):
    _fields_ = [("tile", C.c_uint32), ("layer", C.c_uint32), ("n_entries", C.c_uint32), ("status", C.c_uint32),
                ("offsets_offset", C.c_uint64), ("bytes_offset", C.c_uint64), ("n_bytes", C.c_uint64)]


PROP_COLUMN_DTYPE = np.dtype(PropColumn)
PROP_DICTIONARY_DTYPE = np.dtype(PropDictionary)
assert C.sizeof(PropColumn) == 72 and C.sizeof(PropDictionary) == 40, (C.sizeof(PropColumn), C.sizeof(PropDictionary))


def prop_column_values(blob, c, validity, values, dict_offsets, dictionaries):
    """List<Optional> view of one decoded property column (what CovtParser.decodePropertyColumn returns, CovtParser.java:276-377):
    the value or None per feature. c: a PropColumn record (or a row of PROP_COLUMN_DTYPE); validity / values: the column's buffers
    (values = the buffer of c's value_kind); dictionaries: the PropDictionary records. Host-side convenience for tests and callers."""
    F = int(c["num_features"])
    vo = int(c["validity_offset"])
    valid = np.unpackbits(validity[vo:vo + (F + 7) // 8], bitorder="little")[:F].astype(bool)
    o = int(c["values_offset"])
    kind = int(c["value_kind"])
    # one value slot per feature (Arrow layout): slot i belongs to feature i, slots of absent values hold 0
    if kind == PV_BOOL:
        slots = [bool(b) for b in np.unpackbits(values[o:o + (F + 7) // 8], bitorder="little")[:F]]
    elif kind == PV_I64:
        slots = [int(x) for x in values[o:o + F]]
    elif kind in (PV_F32, PV_F64):
        slots = [float(x) for x in values[o:o + F]]
    elif kind == PV_DICT_INDEX:
        d = dictionaries[int(c["dictionary"])]
        oo, ne, bo, nb = int(d["offsets_offset"]), int(d["n_entries"]), int(d["bytes_offset"]), int(d["n_bytes"])
        off = dict_offsets[oo:oo + ne + 1].astype(np.int64)
        raw = bytes(blob[bo:bo + nb])
        words = [raw[int(off[i]):int(off[i + 1])].decode("utf-8") for i in range(ne)]
        slots = [words[i] if v else None for i, v in zip(values[o:o + F], valid)]
    else:
        slots = [None] * F
    return [s if v else None for s, v in zip(slots, valid)]


LAYER_DTYPE = np.dtype(Layer)
STREAM_DESC_DTYPE = np.dtype(StreamDesc)
assert C.sizeof(StreamRef) == 24 and C.sizeof(StreamDesc) == 48, (C.sizeof(StreamRef), C.sizeof(StreamDesc))
assert C.sizeof(EncodeDesc) == 32, C.sizeof(EncodeDesc)
