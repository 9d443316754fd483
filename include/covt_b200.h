/*
 * covt_b200.h — C ABI of the B200-native COVT (Cloud Optimized Vector Tiles) decoder.
 *
 * Drop-in boundary for the tile-decode path of the reference Java decoder
 * (springmeyer/cov-tiles). Every entry point below replaces one reference interface;
 * citations are <file>:<line> under
 *   J/ = evaluation/java/src/main/java/com/covt/
 *
 *   covt_decode_batch / covt_batch_decode   <- CovtParser.decodeCovt(byte[], TileJson)      J/decoder/CovtParser.java:53
 *                                              (decodeLayerMetadata :574, decodeGeometryColumn :392,
 *                                               convertGeometryColumn :135, decodedIds :552) — batched over tiles
 *   covt_decode_streams                      <- the static stream codecs of DecodingUtils   J/decoder/DecodingUtils.java
 *                                              decodeVarint :35, decodeZigZagVarint :46, decodeZigZagDeltaVarint :55,
 *                                              decodeZigZagDeltaVarintCoordinates :95, decodeRle :257, decodeByteRle :275/:290,
 *                                              decodeFastPfor128ZigZagDelta :316, decodeFastPfor128DeltaCoordinates :349,
 *                                              decodeDeltaVarintMortonCodes :394, decodeFastPfor128DeltaMortonCodes :411
 *   covt_decode_batch_to_host                <- the same call for a host-side consumer (CovtParser.java:87-102 builds its objects on the
 *                                              host): result buffers delivered to page-locked memory while the batch still uploads
 *   covt_encode_streams                      <- the static stream encoders of EncodingUtils  J/converter/EncodingUtils.java
 *                                              encodeVarints :39, encodeRle :123, encodeByteRle :136, encodeFastPfor128 :149,
 *                                              encodeZigZagDeltaCoordinates :190; GeometryUtils.encodeMorton J/converter/GeometryUtils.java:23
 *   covt_result_prop_*                       <- CovtParser.decodePropertyColumn :276-390 (columnar, Arrow layout)
 *   covt_create_multi / covt_decode_batch_multi  the batch scheduler: one call, the GPUs of one box (no reference counterpart)
 *   covt_stream_desc / covt_stream_ref       <- StreamMetadata(streamEncoding,numValues,byteLength)  J/converter/StreamMetadata.java:3
 *   enum values                              <- ordinals of StreamEncoding.java:3-16, StreamType.java:3-16, ColumnType.java:3-9,
 *                                              ColumnDataType.java:3-21, GeometryType (CovtParser.java:20-27): the ordinals ARE the wire values
 *   covt_layer                               <- LayerMetadata + GeometryColumn record (CovtParser.java:29-36) flattened
 *
 * FFM-friendly: plain C, no callbacks, no structs by value, all pointers + sizes. The library
 * never throws or aborts across this boundary: every function returns an int32 status
 * (COVT_OK == 0) and records a message retrievable with covt_last_error().
 *
 * There is NO CPU fallback: if the CUDA device or the sm_100a kernels are unavailable every entry
 * point fails with COVT_ERR_CUDA.
 */
#ifndef COVT_B200_H
#define COVT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COVT_ABI_VERSION 5

/* ---- wire enums (ordinals identical to the Java enums) ------------------------------------ */

/* J/converter/StreamEncoding.java:3-16 */
enum covt_stream_encoding {
    COVT_ENC_PLAIN = 0,
    COVT_ENC_VARINT = 1,
    COVT_ENC_VARINT_ZIG_ZAG = 2,
    COVT_ENC_VARINT_DELTA = 3,
    COVT_ENC_VARINT_DELTA_ZIG_ZAG = 4,
    COVT_ENC_RLE = 5,
    COVT_ENC_BOOLEAN_RLE = 6,
    COVT_ENC_BYTE_RLE = 7,
    COVT_ENC_FAST_PFOR_DELTA = 8,
    COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG = 9,
    COVT_ENC_ABSENT = 0xFF /* not a wire value: slot has no stream */
};

/* J/converter/StreamType.java:3-16 (+ one documented extension) */
enum covt_stream_type {
    COVT_ST_PRESENT = 0,
    COVT_ST_DATA = 1,
    COVT_ST_LENGTH = 2,
    COVT_ST_DICTIONARY = 3,
    COVT_ST_GEOMETRY_TYPES = 4,
    COVT_ST_GEOMETRY_OFFSETS = 5,
    COVT_ST_PART_OFFSETS = 6,
    COVT_ST_RING_OFFSETS = 7,
    COVT_ST_VERTEX_OFFSETS = 8,
    COVT_ST_VERTEX_BUFFER = 9,
    COVT_ST_Z_VALUE = 10,
    COVT_ST_M_VALUE = 11,
    COVT_ST_INDEX_BUFFER = 12 /* EXTENSION: exists only in README prose (README.md:84,114-121); parity unpinned */
};

/* J/converter/ColumnType.java:3-9 */
enum covt_column_type {
    COVT_CT_PLAIN = 0,
    COVT_CT_DICTIONARY = 1,
    COVT_CT_LOCALIZED_DICTIONARY = 2,
    COVT_CT_ICE = 3,
    COVT_CT_ICE_MORTON_CODE = 4
};

/* J/converter/ColumnDataType.java:3-21 (gen-3 wire values) */
enum covt_column_data_type {
    COVT_DT_BOOLEAN = 0, COVT_DT_INT_32 = 1, COVT_DT_UINT_32 = 2, COVT_DT_INT_64 = 3, COVT_DT_UINT_64 = 4,
    COVT_DT_FLOAT = 5, COVT_DT_DOUBLE = 6, COVT_DT_STRING = 7, COVT_DT_GEOMETRY = 8, COVT_DT_GEOMETRY_M = 9,
    COVT_DT_GEOMETRY_Z = 10, COVT_DT_GEOMETRY_ZM = 11, COVT_DT_BINARY = 12, COVT_DT_TIMESTAMP = 13,
    COVT_DT_DATE = 14, COVT_DT_LIST = 15, COVT_DT_STRUCT = 16
};

/* J/decoder/CovtParser.java:20-27 */
enum covt_geometry_type {
    COVT_GT_POINT = 0, COVT_GT_LINESTRING = 1, COVT_GT_POLYGON = 2,
    COVT_GT_MULTIPOINT = 3, /* rejected by encoder (CovtConverter.java:598-600) and decoder (CovtParser.java:268-270) */
    COVT_GT_MULTILINESTRING = 4, COVT_GT_MULTIPOLYGON = 5
};

/* Container generation. */
enum covt_container {
    COVT_CONTAINER_GEN2B = 0, /* committed fixtures: file header + string names + per-stream encoding byte */
    COVT_CONTAINER_GEN3 = 1   /* HEAD CovtParser.decodeLayerMetadata, CovtParser.java:574-652 */
};

/* ---- status codes ----------------------------------------------------------------------- */
enum covt_status {
    COVT_OK = 0,
    COVT_ERR_INVALID_ARG = 1,
    COVT_ERR_CUDA = 2,            /* no device / kernel image missing / CUDA runtime error */
    COVT_ERR_OOM = 3,
    COVT_ERR_TRUNCATED = 4,       /* metadata or stream runs past its tile (Java: ArrayIndexOutOfBounds) */
    COVT_ERR_BAD_METADATA = 5,    /* unknown enum ordinal, first column not id/geometry (CovtParser.java:67-69) */
    COVT_ERR_UNSUPPORTED_ENCODING = 6, /* CovtParser.java:425-427,442-444,459-461,474-476,492-494,507-509,571 */
    COVT_ERR_UNSUPPORTED_GEOMETRY = 7, /* MULTIPOINT / ordinal > 5 (CovtParser.java:268-270) */
    COVT_ERR_VARINT_OVERLONG = 8, /* a varint longer than the Java reader's cap (4 bytes int, 10 bytes long; values follow the
                                   * Java reader), or five consecutive non-final bytes in FastPFOR's VariableByte tail (the
                                   * reference decodes garbage there; values unspecified) */
    COVT_ERR_COUNT_MISMATCH = 9,  /* stream decodes to a different number of values than numValues */
    COVT_ERR_TOPOLOGY = 10        /* counts in topology streams overrun their streams or the vertex buffer */
};

/* ---- decode flags (SURVEY.md §A.6 quirk switches) ------------------------------------------ */
#define COVT_FLAG_CLOSE_RINGS            0x0001u /* append vertex 0 to every polygon ring (CovtParser.java:513-535) */
#define COVT_FLAG_ID_DVZZ_IS_RLE         0x0002u /* fixture quirk: id streams labelled VARINT_DELTA_ZIG_ZAG hold RLE bytes (CovtConverter.java:564-565) */
#define COVT_FLAG_MORTON_NO_SHIFT        0x0004u /* older converter wrote Morton codes without the extent/2 shift (omt zoom 8 fixtures) */
#define COVT_FLAG_ID_WIDTH_32            0x0008u /* emulate Java's int-varint ids (CovtParser.java:557-566); default decodes 64-bit */
#define COVT_FLAG_ICE_VB_COUNT_IS_INTS   0x0010u /* emulate HEAD decoder for ColumnType.ICE (CovtParser.java:499-505): numValues counts ints */
#define COVT_FLAG_SKIP_ASSEMBLY          0x0020u /* decode streams only */
#define COVT_FLAG_PROFILE_KERNELS        0x0040u /* record one CUDA event pair per kernel launch (serialises nothing, adds events) */
#define COVT_FLAG_DECODE_PROPERTIES      0x0080u /* also decode the property columns (CovtParser.decodePropertyColumn, CovtParser.java:276-390) */
#define COVT_FLAG_DEFAULT                (COVT_FLAG_CLOSE_RINGS)

/* ---- slots of a layer's stream table ------------------------------------------------------- */
enum covt_slot {
    COVT_SLOT_ID = 0,       /* id column, DATA stream */
    COVT_SLOT_TYPES = 1,    /* GEOMETRY_TYPES */
    COVT_SLOT_GEOM = 2,     /* GEOMETRY_OFFSETS (counts) */
    COVT_SLOT_PART = 3,     /* PART_OFFSETS (counts) */
    COVT_SLOT_RING = 4,     /* RING_OFFSETS (counts) */
    COVT_SLOT_VOFF = 5,     /* VERTEX_OFFSETS */
    COVT_SLOT_VBUF = 6,     /* VERTEX_BUFFER */
    COVT_SLOT_INDEX = 7,    /* INDEX_BUFFER (extension) */
    COVT_NUM_SLOTS = 8
};

/* Result buffers. S_* are the decoded streams exactly as the reference's GeometryColumn record holds
 * them (CovtParser.java:29-36) plus ids; A_* are the assembled GeoArrow-style buffers that replace
 * the JTS Geometry[] of convertGeometryColumn (CovtParser.java:135-274). */
enum covt_buffer {
    COVT_BUF_S_GEOMETRY_TYPES = 0,   /* u8  [F]            */
    COVT_BUF_S_IDS = 1,              /* i64 [F]            */
    COVT_BUF_S_GEOMETRY_OFFSETS = 2, /* i32 counts          */
    COVT_BUF_S_PART_OFFSETS = 3,     /* i32 counts          */
    COVT_BUF_S_RING_OFFSETS = 4,     /* i32 counts          */
    COVT_BUF_S_VERTEX_OFFSETS = 5,   /* i32                 */
    COVT_BUF_S_VERTEX_BUFFER = 6,    /* i32 x,y interleaved */
    COVT_BUF_S_INDEX_BUFFER = 7,     /* i32 (extension)     */
    COVT_BUF_A_GEOM_OFFSETS = 8,     /* i32 [F+1] per layer -> parts (layer-local) */
    COVT_BUF_A_PART_OFFSETS = 9,     /* i32 [P+1] per layer -> rings               */
    COVT_BUF_A_RING_OFFSETS = 10,    /* i32 [R+1] per layer -> vertices            */
    COVT_BUF_A_COORDS = 11,          /* i32 x,y interleaved, 2 per assembled vertex */
    COVT_BUF_STREAM_ARENA = 12,      /* bytes: output of covt_decode_streams        */
    COVT_NUM_BUFFERS = 13
};

/* ---- descriptors (plain-old-data, identical layout on host and device) ---------------------- */

/* One stream of a layer; mirrors StreamMetadata + where its payload lives in the batch blob. */
typedef struct covt_stream_ref {
    uint64_t byte_offset;  /* absolute offset of the payload in the batch blob */
    uint32_t byte_length;  /* StreamMetadata.byteLength */
    uint32_t num_values;   /* StreamMetadata.numValues (unit per SURVEY §8a dispatch table) */
    uint8_t  encoding;     /* covt_stream_encoding ordinal, COVT_ENC_ABSENT if the slot is empty */
    uint8_t  op;           /* resolved decode routine (covt_op) */
    uint8_t  reserved[2];
    uint32_t status;       /* covt_status of this stream after decode */
} covt_stream_ref;

/* One layer of one tile. */
typedef struct covt_layer {
    uint32_t tile;              /* tile index inside the batch */
    uint32_t layer_index;       /* layer index inside its tile */
    uint32_t extent;
    uint32_t num_features;
    uint32_t num_columns;
    uint32_t status;            /* covt_status: first error met in this layer */
    uint8_t  geom_column_type;  /* covt_column_type of the geometry column */
    uint8_t  num_bits;          /* 32 - nlz(extent), CovtParser.java:77 */
    uint8_t  has_id;
    uint8_t  reserved;
    uint32_t name_length;       /* layer name bytes (gen-2b / non-optimised gen-3) or 0 */
    uint64_t name_offset;       /* absolute blob offset of the UTF-8 layer name; for optimised gen-3: the TileJSON layerId */
    covt_stream_ref streams[COVT_NUM_SLOTS];
    uint64_t out[COVT_NUM_BUFFERS]; /* element offset of this layer's slice in each result buffer */
    uint32_t n_parts;           /* P: assembled parts (written by the assembler) */
    uint32_t n_rings;           /* R: assembled rings */
    uint32_t n_vertices;        /* V: assembled vertices before ring closing (the Mvertices/s unit) */
    uint32_t n_coords;          /* V': vertices written to A_COORDS (V + closed rings) */
    uint32_t cap_parts;         /* allocation of the A_PART_OFFSETS slice minus 1 */
    uint32_t cap_rings;         /* allocation of the A_RING_OFFSETS slice minus 1 */
    uint64_t header_offset;     /* absolute blob offset of the first byte of this layer's metadata */
} covt_layer;

/* Stream-level request: one DecodingUtils call. */
typedef struct covt_stream_desc {
    uint64_t byte_offset;     /* offset of the payload in the blob ("pos" of the Java signature) */
    uint32_t byte_length;     /* bytes available to the stream (Java byteLength; for varint/RLE an upper bound) */
    uint32_t num_values;      /* numValues / numVertices of the Java signature */
    uint8_t  stream_type;     /* covt_stream_type */
    uint8_t  encoding;        /* covt_stream_encoding */
    uint8_t  column_type;     /* covt_column_type */
    uint8_t  column_data_type;/* covt_column_data_type */
    uint8_t  num_bits;        /* Morton bits, CovtParser.java:77 */
    uint8_t  op;              /* 0 = resolve from the four fields above, else force a covt_op */
    uint8_t  reserved[2];
    uint32_t status;          /* out */
    uint32_t bytes_consumed;  /* out: how far "pos" advanced */
    uint64_t out_offset;      /* out: byte offset of the decoded values in COVT_BUF_STREAM_ARENA */
    uint64_t out_count;       /* out: decoded elements (ints; 2 per vertex for Morton ops; bytes for Byte-RLE) */
} covt_stream_desc;

/* Stream-level ENCODE request: one EncodingUtils call. The stream is named by the covt_op that DECODES it; the
 * values are what that op produces (u8 for Byte-RLE, i32 / i64 as the covt_op table says, x,y int pairs for the Morton ops) and
 * the bytes are what the reference's encoder writes for them:
 *   COVT_OP_BYTE_RLE                     EncodingUtils.encodeByteRle :136-147 (orc RunLengthByteWriter)
 *   COVT_OP_RLE_U32 / _U64 / _S64        EncodingUtils.encodeRle :123-134 (orc RunLengthIntegerWriter; signed = zigzag LEB128)
 *   COVT_OP_VARINT_U32 / _ZZ / _ZZ_DELTA EncodingUtils.encodeVarints(values, zigZag, delta) :39-55 (64-bit arithmetic on the widened ints)
 *   COVT_OP_VARINT_ZZ_DELTA_XY           encodeZigZagDeltaCoordinates :190-211, then encodeVarints(.., false, false)
 *   COVT_OP_VARINT_U64 / _ZZ_DELTA_64    encodeVarints on longs (id columns)
 *   COVT_OP_PFOR_ZZ_DELTA / _XY          encodeFastPfor128 :149-188 (Composition(FastPFOR, VariableByte), big-endian words)
 *   COVT_OP_VARINT_DELTA_MORTON / COVT_OP_PFOR_DELTA_MORTON   GeometryUtils.encodeMorton :23-32 of every vertex, deltas without
 *                                        zigzag (CovtConverter.java:939-948); num_values = vertices
 * (the _AS_I64 ops are decode-side width emulations: encode with the 32-bit op). */
typedef struct covt_encode_desc {
    uint64_t value_offset;    /* in: BYTE offset of the stream's first value in `values` (aligned to the value size) */
    uint32_t num_values;      /* in: values (Morton ops: vertices, two ints each) */
    uint8_t  op;              /* in: covt_op */
    uint8_t  num_bits;        /* in: Morton bits */
    uint8_t  reserved[2];
    uint64_t out_offset;      /* out: byte offset of the encoded stream in COVT_BUF_STREAM_ARENA (16-byte aligned) */
    uint32_t byte_length;     /* out: its byteLength */
    uint32_t status;          /* out */
} covt_encode_desc;

/* Decode routines = rows a1..a10 of SURVEY §8a. */
enum covt_op {
    COVT_OP_NONE = 0,
    COVT_OP_BYTE_RLE = 1,              /* decodeByteRle                        -> u8  */
    COVT_OP_RLE_U32 = 2,               /* decodeRle(signed=false) narrowed     -> i32 */
    COVT_OP_RLE_U64 = 3,               /* decodeRle(signed=false)              -> i64 */
    COVT_OP_RLE_S64 = 4,               /* decodeRle(signed=true)               -> i64 */
    COVT_OP_VARINT_U32 = 5,            /* decodeVarint                         -> i32 */
    COVT_OP_VARINT_ZZ = 6,             /* decodeZigZagVarint                   -> i32 */
    COVT_OP_VARINT_ZZ_DELTA = 7,       /* decodeZigZagDeltaVarint              -> i32 */
    COVT_OP_VARINT_ZZ_DELTA_XY = 8,    /* decodeZigZagDeltaVarintCoordinates   -> i32 */
    COVT_OP_VARINT_DELTA_MORTON = 9,   /* decodeDeltaVarintMortonCodes         -> i32 x2 */
    COVT_OP_VARINT_U64 = 10,           /* 64-bit LEB128 (ids, ID_WIDTH 64)     -> i64 */
    COVT_OP_VARINT_ZZ_DELTA_64 = 11,   /* 64-bit zigzag delta (ids)            -> i64 */
    COVT_OP_PFOR_ZZ_DELTA = 12,        /* decodeFastPfor128ZigZagDelta         -> i32 */
    COVT_OP_PFOR_ZZ_DELTA_XY = 13,     /* decodeFastPfor128DeltaCoordinates    -> i32 */
    COVT_OP_PFOR_DELTA_MORTON = 14,    /* decodeFastPfor128DeltaMortonCodes    -> i32 x2 */
    COVT_OP_VARINT_U32_AS_I64 = 15,    /* decodeVarint widened to long (ids, COVT_FLAG_ID_WIDTH_32) */
    COVT_OP_VARINT_ZZ_DELTA_AS_I64 = 16,/* decodeZigZagDeltaVarint widened to long (ids, COVT_FLAG_ID_WIDTH_32; INT_64 property data) */
    COVT_OP_VARINT_ZZ_AS_I64 = 17,     /* decodeZigZagVarint widened to long (INT_64 property data, CovtParser.java:303-306) */
    COVT_NUM_OPS = 18
};

/* ---- property columns (COVT_FLAG_DECODE_PROPERTIES) ------------------------------------------------
 * Replaces CovtParser.decodePropertyColumn (CovtParser.java:276-390) + getStringDictionary (:379-390) with a columnar result
 * in the Apache Arrow layout instead of List<Optional>: per column a VALIDITY bitmap (bit i = feature i has a value; byte i >> 3,
 * bit i & 7 = java.util.BitSet order = Arrow's) and ONE VALUE SLOT PER FEATURE (the slots of features without a value hold 0) —
 * the null expansion of CovtParser.java:317-326,331-340,354-364 done on the device. Strings are dictionary indices; a dictionary
 * is an int32 offsets array into the tile's own UTF-8 bytes, which stay in the input blob (Arrow utf8 layout). Localized
 * dictionary columns (gen-2b fixtures) are flattened: one column per sub-key, all pointing at the shared dictionary. */
enum covt_prop_value_kind {
    COVT_PV_NONE = 0, COVT_PV_I64 = 1, COVT_PV_F32 = 2, COVT_PV_F64 = 3,
    COVT_PV_BOOL = 4,       /* dense bits, BitSet order */
    COVT_PV_DICT_INDEX = 5  /* i32 indices into dictionaries[dictionary] */
};
enum covt_prop_buffer {
    COVT_PBUF_VALIDITY = 0,     /* u8: ceil(num_features / 8) bytes per column */
    COVT_PBUF_I64 = 1,          /* i64 */
    COVT_PBUF_F32 = 2,          /* f32 (DecodingUtils.decodeFloatsLE :446-453) */
    COVT_PBUF_F64 = 3,          /* f64 */
    COVT_PBUF_BOOL = 4,         /* u8: ceil(num_features / 8) bytes per BOOLEAN column (bit i = value of feature i) */
    COVT_PBUF_DICT_INDEX = 5,   /* i32 */
    COVT_PBUF_DICT_OFFSETS = 6, /* i32: n_entries + 1 byte offsets per dictionary, relative to its bytes_offset */
    COVT_NUM_PROP_BUFFERS = 7
};
typedef struct covt_prop_column {
    uint32_t tile, layer;        /* layer = index within the tile */
    uint64_t name_offset;        /* column name in the blob; optimised gen-3: index into the TileJSON fields of the layer */
    uint64_t sub_offset;         /* localized sub-key name in the blob (sub_length 0: not a localized sub-column) */
    uint32_t name_length, sub_length;
    uint8_t  data_type;          /* covt_column_data_type (HEAD ordinals) */
    uint8_t  column_type;        /* covt_column_type */
    uint8_t  value_kind;         /* covt_prop_value_kind */
    uint8_t  reserved;
    uint32_t status;             /* covt_status of the column */
    uint32_t num_features;
    uint32_t num_values;         /* set bits of the validity bitmap = features that have a value */
    uint64_t validity_offset;    /* bytes into COVT_PBUF_VALIDITY */
    uint64_t values_offset;      /* elements into the buffer of value_kind: num_features slots (COVT_PV_BOOL: BYTES into COVT_PBUF_BOOL) */
    uint32_t dictionary;         /* COVT_PV_DICT_INDEX: index into the dictionaries (0 for a column without a dictionary of its own) */
    uint32_t data_num_values;    /* numValues the data stream declares (== num_values for a column with status COVT_OK) */
} covt_prop_column;
typedef struct covt_prop_dictionary {
    uint32_t tile, layer, n_entries;
    uint32_t status;             /* covt_status of the length stream / offsets */
    uint64_t offsets_offset;     /* elements into COVT_PBUF_DICT_OFFSETS: n_entries + 1 offsets relative to bytes_offset */
    uint64_t bytes_offset;       /* the UTF-8 bytes of all entries, back to back, in the blob */
    uint64_t n_bytes;
} covt_prop_dictionary;

/* Optional TileJSON side-car for optimised gen-3 metadata (CovtParser.java:583-590): only the number
 * of fields per vector layer matters to the decode path (column ids >= 2 index the fields). */
typedef struct covt_tilejson {
    uint32_t n_vector_layers;
    const uint32_t* n_fields; /* [n_vector_layers] */
} covt_tilejson;

typedef struct covt_timing {
    float h2d_ms;        /* blob + tile offsets host->device */
    float decode_ms;     /* first kernel start -> last kernel end (device-resident decode) */
    float d2h_ms;        /* status + layer table device->host (only what the call itself copied) */
    uint32_t kernel_launches;
    uint64_t payload_bytes;   /* sum of byteLength of every decoded stream */
    uint64_t output_bytes;    /* bytes of all result buffers written */
    uint64_t vertices;        /* sum of covt_layer.n_vertices */
    uint32_t segments;        /* upload/decode segments the call was pipelined over (1 = not pipelined) */
    uint32_t capacity_retries;/* 1 if an extrapolated result capacity was too small and the batch was decoded again with exact sizes */
} covt_timing;

/* per-kernel times, available when COVT_FLAG_PROFILE_KERNELS was set */
typedef struct covt_kernel_time {
    char     name[48];
    float    ms;         /* summed over launches */
    uint32_t launches;
    uint64_t algorithmic_bytes; /* bytes the kernel must read + write (DESIGN.md "algorithmic bytes") */
} covt_kernel_time;

typedef struct covt_ctx covt_ctx;
typedef struct covt_batch covt_batch;
typedef struct covt_result covt_result;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int32_t covt_abi_version(void);
/* One context per GPU; several contexts (on the same or on different GPUs) may coexist in one process. device = CUDA ordinal.
 * THREADING: a context, its batches and its results must be used by ONE thread at a time (no internal locking: the pinned scratch,
 * the device block cache and the error string are per context). Different contexts may be used from different threads at the same
 * time — that is how covt_decode_batch_multi drives several GPUs — so a multi-threaded caller (e.g. a JVM thread pool) creates
 * one context per thread. */
int32_t covt_create(int32_t device, covt_ctx** out);
/* Free the context's batches and results first: they borrow its streams and device blocks (a result freed after its context is
 * undefined behaviour, like free() after the allocator is gone). */
void    covt_destroy(covt_ctx* ctx);
/* Copies the last error message of this context (or of a failed covt_create when ctx == NULL). */
int32_t covt_last_error(covt_ctx* ctx, char* buf, size_t buf_len);
/* The context parks the large device blocks of finished decodes (input blob, result arena, layer table) for the next call;
 * covt_trim returns them to the driver. */
int32_t covt_trim(covt_ctx* ctx);

/* ---- batch path: replaces CovtParser.decodeCovt (CovtParser.java:53), batched over tiles ---- */
/* blob holds n_tiles tiles back to back; tile i occupies [tile_offsets[i], tile_offsets[i+1]). Host memory (pinned or
 * covt_host_register-ed for full PCIe rate). Large batches go up in at most 8 segments of >= 64 MiB on a copy stream while earlier segments
 * are being decoded; the result is one set of buffers whatever the segmentation. */
int32_t covt_decode_batch(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles,
                          uint32_t container, const covt_tilejson* tilejson, uint32_t flags, covt_result** out);
/* The same with the results delivered to HOST memory: every buffer the sink names (ptr != NULL; page-locked memory, capacity in
 * elements) receives the decoded buffer, segment by segment on its own stream while later segments are still being uploaded and
 * decoded (PCIe is full duplex) — what a List<Layer> caller with a host-side consumer needs (CovtParser.java:87-102 materialises
 * JTS objects on the host). The result still owns the device-resident buffers, layer table and statuses; covt_timing.d2h_ms =
 * first to last device->host copy. A sink buffer that is too small fails the call with COVT_ERR_INVALID_ARG (the message names the
 * sizes needed). */
typedef struct covt_host_sink {
    void*    ptr[COVT_NUM_BUFFERS];
    uint64_t capacity[COVT_NUM_BUFFERS];
} covt_host_sink;
int32_t covt_decode_batch_to_host(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles,
                                  uint32_t container, const covt_tilejson* tilejson, uint32_t flags, const covt_host_sink* sink,
                                  covt_result** out);
/* The same in two steps so that host->device transfer is timed apart from device-resident decode. */
int32_t covt_batch_upload(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles,
                          covt_batch** out);
int32_t covt_batch_decode(covt_ctx* ctx, covt_batch* batch, uint32_t container, const covt_tilejson* tilejson,
                          uint32_t flags, covt_result** out);
void    covt_batch_free(covt_batch* batch);

/* ---- stream path: replaces the static codecs of DecodingUtils ------------------------------- */
int32_t covt_decode_streams(covt_ctx* ctx, const uint8_t* blob, uint64_t blob_len, covt_stream_desc* descs,
                            uint32_t n_streams, uint32_t flags, covt_result** out);
int32_t covt_batch_decode_streams(covt_ctx* ctx, covt_batch* batch, covt_stream_desc* descs, uint32_t n_streams,
                                  uint32_t flags, covt_result** out);
/* Encodes n streams of host `values` on the GPU; the encoded bytes are COVT_BUF_STREAM_ARENA of the result (device-resident, read
 * with covt_result_read). flags: COVT_FLAG_MORTON_NO_SHIFT. covt_timing of the result: h2d_ms = upload of the values, decode_ms =
 * device time of the encode, payload_bytes = bytes written, output_bytes = value bytes read. Replaces the static encoders of
 * EncodingUtils (J/converter/EncodingUtils.java:39-230). */
int32_t covt_encode_streams(covt_ctx* ctx, const void* values, uint64_t values_bytes, covt_encode_desc* descs, uint32_t n_streams,
                            uint32_t flags, covt_result** out);
/* Dispatch table of CovtParser.decodeGeometryColumn / decodedIds: which routine decodes a stream. */
int32_t covt_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags);

/* ---- results (owned by the library until covt_result_free) ---------------------------------- */
uint32_t covt_result_num_tiles(const covt_result* res);
uint32_t covt_result_num_layers(const covt_result* res);
/* Host copies (pinned, fetched lazily on first call). */
int32_t covt_result_layers(covt_result* res, const covt_layer** layers);
int32_t covt_result_tile_status(covt_result* res, const uint32_t** status, const uint32_t** first_layer /* [n_tiles+1] */);
/* Device pointer + element count + element size of one result buffer. */
int32_t covt_result_buffer(const covt_result* res, uint32_t which, const void** device_ptr, uint64_t* count,
                           uint32_t* elem_size);
/* Device->host copy of count elements starting at elem_offset of buffer `which`. */
int32_t covt_result_read(covt_result* res, uint32_t which, uint64_t elem_offset, uint64_t count, void* host_dst);
/* Property columns (COVT_FLAG_DECODE_PROPERTIES; zero columns otherwise). Host copies of the records (pinned, fetched lazily),
 * device pointer / read-back of the value buffers (covt_prop_buffer). */
int32_t covt_result_prop_columns(covt_result* res, const covt_prop_column** columns, uint32_t* n_columns);
int32_t covt_result_prop_dictionaries(covt_result* res, const covt_prop_dictionary** dictionaries, uint32_t* n_dictionaries);
int32_t covt_result_prop_buffer(const covt_result* res, uint32_t which, const void** device_ptr, uint64_t* count, uint32_t* elem_size);
int32_t covt_result_prop_read(covt_result* res, uint32_t which, uint64_t elem_offset, uint64_t count, void* host_dst);
int32_t covt_result_timing(const covt_result* res, covt_timing* out);
/* Fills up to cap entries, returns the number of distinct kernels in *n. */
int32_t covt_result_kernel_times(const covt_result* res, covt_kernel_time* out, uint32_t cap, uint32_t* n);
void    covt_result_free(covt_result* res);

/* ---- host memory helpers ----------------------------------------------------------------------- */
/* Page-locks a caller-owned host range (e.g. a Java MemorySegment) so that uploads run at full PCIe rate. */
int32_t covt_host_register(covt_ctx* ctx, void* ptr, size_t bytes);
int32_t covt_host_unregister(covt_ctx* ctx, void* ptr);

/* ---- batch scheduler helper (host only) ------------------------------------------------------ */
/* Splits tiles [0,n_tiles) into n_parts contiguous ranges balanced by payload bytes
 * (prefix sum over tile_offsets). starts has n_parts+1 entries. No collective: tiles share nothing. */
int32_t covt_partition_tiles(const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t n_parts, uint32_t* starts);

/* ---- multi-GPU batch scheduler: one call, N GPUs of one box ------------------------------------------ */
/* Replaces the same single entry point (CovtParser.decodeCovt, CovtParser.java:53) for a caller that owns several GPUs from
 * one process (a single JVM): the batch is cut into contiguous tile ranges balanced by payload bytes (covt_partition_tiles),
 * range g is uploaded and decoded by GPU g on the scheduler's own host thread with that GPU's own context, streams and result
 * arena. No collective and no peer traffic: tiles share nothing. Results stay resident on the GPU that decoded them; the
 * result handle exposes one covt_result per GPU. A covt_multi handle serves one call at a time. */
typedef struct covt_multi covt_multi;
typedef struct covt_multi_result covt_multi_result;
/* device_count == 0: every visible GPU. device_ids == NULL: ordinals 0 .. device_count-1. */
int32_t covt_create_multi(uint32_t device_count, const int32_t* device_ids, covt_multi** out);
void    covt_destroy_multi(covt_multi* m);
int32_t covt_multi_last_error(covt_multi* m, char* buf, size_t buf_len);
uint32_t covt_multi_device_count(const covt_multi* m);
/* The single-GPU context of part `part` (owned by the scheduler), e.g. for covt_host_register / covt_trim. */
int32_t covt_multi_context(covt_multi* m, uint32_t part, covt_ctx** ctx, int32_t* device);
int32_t covt_decode_batch_multi(covt_multi* m, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles,
                                uint32_t container, const covt_tilejson* tilejson, uint32_t flags, covt_multi_result** out);
uint32_t covt_multi_result_parts(const covt_multi_result* res);
/* Part `part` = tiles [first_tile, first_tile + n_tiles) of the batch, decoded on `device`. Tile indices inside the part's
 * covt_result (covt_layer.tile, the status arrays) are relative to first_tile. The covt_result stays owned by the multi result. */
int32_t covt_multi_result_part(const covt_multi_result* res, uint32_t part, covt_result** part_result, uint32_t* first_tile,
                               uint32_t* n_tiles, int32_t* device);
/* Times: maximum over the GPUs (they run side by side); counters: sums. */
int32_t covt_multi_result_timing(const covt_multi_result* res, covt_timing* out);
/* Device->host copy of buffer `which` of every part at once (one copy thread per GPU): part p -> host_dst[p]. */
int32_t covt_multi_result_read(covt_multi_result* res, uint32_t which, void* const* host_dst);
void    covt_multi_result_free(covt_multi_result* res);

#ifdef __cplusplus
}
#endif
#endif /* COVT_B200_H */
