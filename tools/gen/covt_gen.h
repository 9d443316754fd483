/*
 * covt_gen.h — COVT encoders + synthetic tile generator (input synthesis for bench.py and the tests).
 *
 * Restates the ENCODER side of the reference so that synthetic inputs are what the reference converter
 * would write (validated by byte-identical re-encoding of the fixture streams):
 *   J/converter/EncodingUtils.java:39-230   varint / zigzag / delta, encodeRle, encodeByteRle, encodeFastPfor128
 *   J/converter/GeometryUtils.java:23-32    encodeMorton
 *   J/converter/CovtConverter.java:571-986  stream selection ("encode both ways, keep the shorter")
 *   orc-core 1.8.1 RunLengthIntegerWriter / RunLengthByteWriter, JavaFastPFOR 0.1.12 Composition(FastPFOR,VariableByte)
 * Not part of the decode product path and not the oracle: it produces inputs, it never decodes.
 */
#ifndef COVT_GEN_H
#define COVT_GEN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- stream encoders: return bytes written, or (size_t)-1 if cap is too small ------------------ */
size_t covt_enc_varints(const int64_t* v, size_t n, int zigzag, int delta, uint8_t* out, size_t cap);
size_t covt_enc_rle(const int64_t* v, size_t n, int is_signed, uint8_t* out, size_t cap);
size_t covt_enc_byte_rle(const uint8_t* v, size_t n, uint8_t* out, size_t cap);
size_t covt_enc_fastpfor(const int32_t* v, size_t n, int zigzag, int delta, uint8_t* out, size_t cap);
/* EncodingUtils.encodeZigZagDeltaCoordinates (:190-211): x,y interleaved in, zigzag deltas out */
void covt_enc_zigzag_delta_coordinates(const int32_t* xy, size_t n_ints, int32_t* out);
int32_t covt_enc_morton(int32_t x, int32_t y, uint32_t num_bits);

/* ---- layer model (what a converter sees after reading an MVT layer) ----------------------------- */
typedef struct covt_gen_layer {
    const char* name;
    uint32_t extent;
    uint32_t n_features;
    const uint8_t* types;       /* [n_features] GeometryType ordinals */
    const int32_t* geom_counts; uint32_t n_geom;   /* geometry_offsets stream (counts) */
    const int32_t* part_counts; uint32_t n_part;
    const int32_t* ring_counts; uint32_t n_ring;   /* without closing vertex */
    const int32_t* xy;          uint32_t n_vertices; /* vertices in feature order, x,y interleaved, no closing vertices */
    const int64_t* ids;         /* [n_features] or NULL */
    const int32_t* index_buffer; uint32_t n_index;  /* extension stream or NULL */
} covt_gen_layer;

#define COVT_GEN_ALLOW_PFOR_TOPOLOGY  0x01u
#define COVT_GEN_ALLOW_PFOR_VERTEX    0x02u
#define COVT_GEN_ICE_MORTON           0x04u  /* build a sorted Morton vertex dictionary + vertex_offsets */
#define COVT_GEN_ID_DELTA_VARINT      0x08u  /* ids as genuine 64-bit VARINT_DELTA_ZIG_ZAG instead of the shorter of RLE/VARINT */
#define COVT_GEN_FORCE_VARINT_VERTEX  0x10u
#define COVT_GEN_FORCE_RLE_TOPOLOGY   0x20u

typedef struct covt_gen_buf { uint8_t* data; size_t len, cap; } covt_gen_buf;
void covt_gen_buf_free(covt_gen_buf* b);

/* Appends one encoded layer (metadata + payload) to `tile`. container: 0 = gen-2b, 1 = gen-3 (non-optimised
 * metadata: string layer name, geometry column id 1; id columns cannot be expressed at HEAD and are written
 * as column id 0 — SURVEY §A.6 HEAD_ID_COLUMN). Returns 0 on success. */
int32_t covt_gen_append_layer(covt_gen_buf* tile, const covt_gen_layer* layer, uint32_t container, uint32_t options);
/* gen-2b file header: varint version (=1), varint numLayers */
int32_t covt_gen_begin_tile(covt_gen_buf* tile, uint32_t container, uint32_t num_layers);

/* ---- synthetic workloads ---------------------------------------------------------- */
/* Config 3: one PLAIN VERTEX_BUFFER / VARINT_DELTA_ZIG_ZAG stream of exactly target_bytes bytes; varint lengths
 * 1/2/3/4 B with P = .531/.426/.040/.003, xorshift64* seeded with `seed`, even value count. Returns #ints. */
uint64_t covt_gen_varint_stream(uint8_t* out, uint64_t target_bytes, uint64_t seed);

typedef struct covt_gen_params {
    uint32_t layers_per_tile;      /* 2 */
    double   mean_features;        /* geometric, mean 48 */
    double   p_point, p_line, p_polygon, p_multiline, p_multipolygon; /* .12 .70 .15 .015 .015 */
    double   mean_line_extra;      /* vertices/linestring = 2 + geometric(mean 6) */
    double   mean_ring_extra;      /* vertices/ring = 3 + geometric(mean 5) */
    double   p_second_ring;        /* .10 */
    uint32_t extent;               /* 4096 */
    uint32_t container;            /* 0 gen-2b, 1 gen-3 */
    uint32_t with_ids;             /* 1 */
    uint32_t with_index_buffer;    /* 0; 1 = fan-triangulation INDEX_BUFFER on polygon layers (config 4) */
    uint32_t max_step;             /* coordinate random-walk step bound (default 48) */
} covt_gen_params;
void covt_gen_default_params(covt_gen_params* p);

typedef struct covt_gen_truth { /* what was encoded, for round-trip checks */
    uint64_t features, vertices, parts, rings, polygon_rings;
    int64_t  sum_x, sum_y;      /* over assembled vertices WITHOUT closing vertices */
    int64_t  sum_x_closed, sum_y_closed; /* including one closing vertex per non-empty polygon ring */
} covt_gen_truth;

/* Config 5: tiles [first_tile, first_tile + n_tiles), seed = tile index. Allocates *blob (free with covt_gen_free)
 * and fills tile_offsets[n_tiles+1] (caller-provided). truth is summed over the tiles (nullable). */
int32_t covt_gen_tiles(uint64_t first_tile, uint32_t n_tiles, const covt_gen_params* p, uint32_t n_threads,
                       uint8_t** blob, uint64_t* blob_len, uint64_t* tile_offsets, covt_gen_truth* truth);
void covt_gen_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
