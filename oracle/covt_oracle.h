/*
 * covt_oracle.h — CPU ORACLE for the COVT tile-decode path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement of the reference Java decoder (springmeyer/cov-tiles):
 *   J/decoder/DecodingUtils.java, J/decoder/CovtParser.java, J/converter/GeometryUtils.java
 * plus the two third-party codecs the reference calls but does not vendor:
 *   org.apache.orc:orc-core:1.8.1        (RunLengthIntegerReader / RunLengthByteReader, RLE v1)
 *   me.lemire.integercompression:JavaFastPFOR:0.1.12  (Composition(FastPFOR, VariableByte))
 * restated from their published algorithms.
 *
 * PARITY PINNING: the JVM is absent in this image, so the Java reference cannot be executed.
 * The oracle is pinned by (a) the reference's own known-answer vectors
 * (parser/js/test/unit/decoder/decodingUtils.spec.ts:10-113), (b) all 129 gen-2b fixture tiles in
 * test/fixtures: EOF-exact container walk, every stream consumes exactly its declared byteLength,
 * assembled geometry equals the partner .mvt/.pbf, re-encoding is byte-identical
 * (tests/test_oracle_fixtures.py, tests/golden/). INDEX_BUFFER (stream type 12) is an extension
 * with no reference implementation: parity unpinned for that stream type only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (cov-tiles_b200/) never links, imports or calls it.
 */
#ifndef COVT_ORACLE_H
#define COVT_ORACLE_H

#include "../include/covt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- DecodingUtils.java restated (each advances *pos like the Java IntWrapper) ---------------- */
/* status: COVT_OK or COVT_ERR_TRUNCATED. *overlong (nullable) is set when a value hit the
 * Java reader's 4-byte cap with the continuation bit still set (DecodingUtils.java:182-185). */
int32_t covt_oracle_decode_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong);
int32_t covt_oracle_decode_zigzag_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong);
int32_t covt_oracle_decode_zigzag_delta_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong);
int32_t covt_oracle_decode_zigzag_delta_varint_coordinates(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong);
int32_t covt_oracle_decode_delta_varint_morton_codes(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n_vertices,
                                                     uint32_t num_bits, int no_shift, int32_t* out /*2n*/, int* overlong);
/* 64-bit id variants (ID_WIDTH 64): full LEB128 as written by EncodingUtils.putVarInt (:105-114) */
int32_t covt_oracle_decode_varint64(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int64_t* out, int* overlong);
int32_t covt_oracle_decode_zigzag_delta_varint64(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int64_t* out, int* overlong);
/* orc RunLengthIntegerReader / RunLengthByteReader (DecodingUtils.java:257-306) */
int32_t covt_oracle_decode_rle(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int is_signed, int64_t* out);
int32_t covt_oracle_decode_byte_rle(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint8_t* out);
/* Composition(FastPFOR, VariableByte).uncompress over big-endian words (DecodingUtils.java:317-333).
 * out has n entries, zero-filled first like the Java `new int[numValues]`. */
int32_t covt_oracle_fastpfor_uncompress(const uint8_t* buf, uint32_t byte_length, uint32_t n, int32_t* out);
int32_t covt_oracle_decode_fastpfor_zigzag_delta(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint32_t byte_length, int32_t* out);
int32_t covt_oracle_decode_fastpfor_delta_coordinates(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint32_t byte_length, int32_t* out);
int32_t covt_oracle_decode_fastpfor_delta_morton_codes(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n_vertices,
                                                       uint32_t byte_length, uint32_t num_bits, int no_shift, int32_t* out /*2n*/);
/* GeometryUtils.decodeMorton (GeometryUtils.java:34-47) */
void covt_oracle_decode_morton(int32_t code, uint32_t num_bits, int no_shift, int32_t* x, int32_t* y);

/* ---- dispatch (CovtParser.decodeGeometryColumn :392-511, decodedIds :552-572) ----------------- */
int32_t covt_oracle_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags);
uint32_t covt_oracle_op_elem_size(uint32_t op);
uint64_t covt_oracle_op_out_count(uint32_t op, uint32_t num_values);
/* One DecodingUtils call described by d; writes d->status, d->bytes_consumed, d->out_count. */
int32_t covt_oracle_decode_stream(const uint8_t* blob, uint64_t blob_len, covt_stream_desc* d, uint32_t flags,
                                  void* out, uint64_t out_cap_bytes);

/* ---- container walkers (gen-3 = CovtParser.decodeLayerMetadata :574-652) --------- */
/* Parses the tile in blob[begin,end) into layers[0..cap). *end_pos = cursor after the last layer
 * (must equal `end` for a well-formed tile). Returns the tile status. */
int32_t covt_oracle_parse_tile(const uint8_t* blob, uint64_t begin, uint64_t end, uint32_t container,
                               const covt_tilejson* tj, uint32_t flags, uint32_t tile_index,
                               covt_layer* layers, uint32_t cap, uint32_t* n_layers, uint64_t* end_pos);

/* ---- whole batch, same result layout as libcovt_b200 (DESIGN.md "result layout") -------------- */
typedef struct covt_oracle_result {
    uint32_t n_tiles, n_layers;
    uint32_t* tile_status;   /* [n_tiles] */
    uint32_t* first_layer;   /* [n_tiles+1] */
    covt_layer* layers;      /* [n_layers] */
    void* buffers[COVT_NUM_BUFFERS];
    uint64_t counts[COVT_NUM_BUFFERS]; /* elements */
    uint64_t payload_bytes, vertices;
} covt_oracle_result;

int32_t covt_oracle_decode_batch(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                 const covt_tilejson* tj, uint32_t flags, uint32_t n_threads, covt_oracle_result** out);
/* Decode without keeping the outputs: the timed CPU-baseline loop (reuses per-thread scratch). Returns
 * payload bytes and assembled vertices; checksum guards against dead-code elimination. */
int32_t covt_oracle_decode_batch_timed(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                       const covt_tilejson* tj, uint32_t flags, uint32_t n_threads,
                                       uint64_t* payload_bytes, uint64_t* vertices, uint64_t* checksum);
void covt_oracle_result_free(covt_oracle_result* r);

/* ---- property columns -----------------------------------------------------------------------------------------
 * CovtParser.decodePropertyColumn (J/decoder/CovtParser.java:276-390) restated with the columnar, Arrow-like result of
 * include/covt_b200.h (covt_prop_column, covt_prop_dictionary, the seven value buffers): per column a validity bitmap (bit i =
 * feature i has a value, java.util.BitSet order) and the DENSE values of the features that have one. Localized dictionary columns
 * are flattened: one column per sub-key, all pointing at the shared dictionary. Strings are dictionary indices; a dictionary is an
 * offsets array into the tile's own UTF-8 bytes (left in the input blob). gen-2b and gen-3 containers; test infrastructure like the
 * rest. */
typedef struct covt_oracle_props {
    uint32_t n_tiles, n_columns, n_dictionaries, reserved;
    uint32_t* tile_status;       /* [n_tiles] status of the container walk of each tile (the same walk as the geometry path) */
    covt_prop_column* columns;
    covt_prop_dictionary* dictionaries;
    uint8_t* validity;  uint64_t validity_bytes;
    int64_t* i64;       uint64_t n_i64;
    float*   f32;       uint64_t n_f32;
    double*  f64;       uint64_t n_f64;
    uint8_t* bools;     uint64_t bool_bytes;
    int32_t* dict_index; uint64_t n_dict_index;
    int32_t* dict_offsets; uint64_t n_dict_offsets;
    uint64_t payload_bytes;      /* byteLength of every property stream handed to a codec (+ dictionary bytes): the metric's unit */
} covt_oracle_props;
int32_t covt_oracle_decode_properties(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                      const covt_tilejson* tj, uint32_t flags, covt_oracle_props** out);
void covt_oracle_props_free(covt_oracle_props* p);
uint32_t covt_oracle_buffer_elem_size(uint32_t which);

#ifdef __cplusplus
}
#endif
#endif
