// covt_walk.cuh — device-side container walk (K0): one thread per tile, gen-2b and gen-3.
//
// Replaces CovtParser.decodeLayerMetadata (J/decoder/CovtParser.java:574-652, gen-3) and the gen-2b grammar of the committed
// fixtures. The layer being parsed lives in SHARED memory (struct-of-arrays over the block's threads: the stream
// slot is a run-time index, and a thread-private covt_layer in local memory cost 592 bytes of stack per thread, which thrashed L1
// at 12 resident warps per SM); nothing of the walk touches local memory.
//
// Payload placement follows the reference: the decoder consumes the columns IN METADATA ORDER (CovtParser.java:64-85, a
// LinkedHashMap), the geometry streams inside their column in the fixed order types, geometry_offsets, part_offsets, ring_offsets,
// vertex_offsets, vertex_buffer (:405-510). The common orders (id, geometry, properties... / geometry, id, properties...) are
// placed straight from the first metadata pass; any other order, gen-3 layers with property columns (their PRESENT streams are not
// listed, CovtConverter.java:434-436, so their length comes from walking the Byte-RLE bytes) and property decoding take a second
// pass over the column metadata. There is no limit on the number of property columns.
#pragma once
#include "covt_device.cuh"

namespace covt {

#ifndef K0_BLOCK_THREADS
#define K0_BLOCK_THREADS 64
#endif
constexpr int K0_BLOCK = K0_BLOCK_THREADS;

struct Cursor { const uint8_t* b; uint64_t p, end; bool err; };

// DecodingUtils.decodeVarint (DecodingUtils.java:157-186): at most 4 bytes
__device__ __forceinline__ uint32_t c_varint(Cursor& c)
{
    uint32_t v = 0;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        if (c.p >= c.end) { c.err = true; return 0; }
        const uint32_t b = __ldg(c.b + c.p);
        c.p++;
        v |= (b & 0x7fu) << (7 * i);
        if (!(b & 0x80u)) break;
    }
    return v;
}
__device__ __forceinline__ uint32_t c_byte(Cursor& c)
{
    if (c.p >= c.end) { c.err = true; return 0; }
    const uint32_t b = __ldg(c.b + c.p);
    c.p++;
    return b;
}
// DecodingUtils.decodeString (:21-26)
__device__ __forceinline__ void c_string(Cursor& c, uint64_t& off, uint32_t& len)
{
    const uint32_t n = c_varint(c);
    if (c.err || c.p + n > c.end) { c.err = true; off = 0; len = 0; return; }
    off = c.p;
    len = n;
    c.p += n;
}
template <int N>
__device__ __forceinline__ bool name_is(const uint8_t* b, uint64_t off, uint32_t len, const char (&lit)[N])
{
    if (len != N - 1) return false;
#pragma unroll
    for (int i = 0; i < N - 1; i++)
        if (__ldg(b + off + i) != (uint8_t)lit[i]) return false;
    return true;
}
template <int N>
__device__ __forceinline__ bool name_starts(const uint8_t* b, uint64_t off, uint32_t len, const char (&lit)[N])
{
    if (len < N - 1) return false;
#pragma unroll
    for (int i = 0; i < N - 1; i++)
        if (__ldg(b + off + i) != (uint8_t)lit[i]) return false;
    return true;
}

// The result buffers are sized from numValues before anything is decoded, so a corrupt header must not be able to claim
// gigabytes: no codec of the path packs more than 128 values into a byte (FastPFOR at bit width 0: two container bytes per block
// of 256; Byte-RLE: 130 per 2 bytes; RLE: 130 per 3), so a stream that claims more than 256 values per byte (+ slack for the
// fixed headers) cannot decode — the reference would run off the end of the array (ArrayIndexOutOfBounds): the tile fails.
__host__ __device__ __forceinline__ bool plausible_count(uint32_t num_values, uint32_t byte_length)
{
    return (uint64_t)num_values <= 256ull * ((uint64_t)byte_length + 16ull);
}

// bytes of a Byte-RLE stream that decodes to n bytes: unlisted PRESENT streams of gen-3 property columns
// (CovtConverter.java:434-436, DecodingUtils.java:290-306)
__device__ __forceinline__ bool byte_rle_span(const uint8_t* b, uint64_t& p, uint64_t end, uint32_t n)
{
    uint32_t done = 0;
    while (done < n) {
        if (p >= end) return false;
        const uint32_t c = __ldg(b + p);
        p++;
        if (c < 0x80u) { done += c + 3u; p += 1; }
        else { done += 256u - c; p += 256u - c; }
        if (p > end) return false;
    }
    return true;
}

// The streams of the layer being parsed: numValues, byteLength, encoding of the 8 slots, word w of thread t at s[w * K0_BLOCK].
constexpr int LITE_WORDS = 3 * COVT_NUM_SLOTS;
struct Lite {
    uint32_t* s;  // &block_scratch[threadIdx.x]
    __device__ __forceinline__ uint32_t& nv(int slot) const { return s[slot * K0_BLOCK]; }
    __device__ __forceinline__ uint32_t& bl(int slot) const { return s[(COVT_NUM_SLOTS + slot) * K0_BLOCK]; }
    __device__ __forceinline__ uint32_t& enc(int slot) const { return s[(2 * COVT_NUM_SLOTS + slot) * K0_BLOCK]; }
    __device__ __forceinline__ bool has(int slot) const { return enc(slot) != COVT_ENC_ABSENT; }
    __device__ __forceinline__ void clear() const
    {
#pragma unroll
        for (int s_ = 0; s_ < COVT_NUM_SLOTS; s_++) { nv(s_) = 0; bl(s_) = 0; enc(s_) = COVT_ENC_ABSENT; }
    }
    __device__ __forceinline__ void set(int slot, uint32_t n, uint32_t b, uint32_t e) const { nv(slot) = n; bl(slot) = b; enc(slot) = e; }
};

struct LayerHead {
    uint64_t layer_start;   // blob offset of the first byte of the layer's metadata
    uint64_t name_offset;   // blob offset of the UTF-8 layer name; optimised gen-3: the TileJSON layerId
    uint32_t name_length, extent, num_features, num_columns;
    uint32_t geom_ct;       // covt_column_type of the geometry column
    uint64_t id_offset;     // payload position of the id column's DATA stream
    uint64_t geom_offset;   // payload position of the geometry column (its streams follow each other in slot order)
};

// What a walker does with property columns. want() == false: they are only hopped over.
//   set_layer(layer_index)                                                                          before a layer is walked
//   begin_column(name_offset, name_length, data_type /*covt_column_data_type, 0xFF = unknown*/, column_type, num_features)
//   stream(stream_type, sub_offset, sub_length, num_values, byte_length, encoding, payload_offset)   in payload order
//   end_column()
struct NoProps {
    __device__ __forceinline__ bool want() const { return false; }
    __device__ __forceinline__ void set_layer(uint32_t) {}
    __device__ __forceinline__ void begin_column(uint64_t, uint32_t, uint32_t, uint32_t, uint32_t) {}
    __device__ __forceinline__ void stream(uint32_t, uint64_t, uint32_t, uint32_t, uint32_t, uint32_t, uint64_t) {}
    __device__ __forceinline__ void end_column() {}
};

// gen-2 data type byte of the column header -> covt_column_data_type; 0xFF = unknown
__device__ __forceinline__ uint32_t dt_of_gen2(uint32_t g)
{
    switch (g) {
    case 0: return COVT_DT_STRING;
    case 1: return COVT_DT_FLOAT;
    case 2: return COVT_DT_DOUBLE;
    case 3: return COVT_DT_INT_64;
    case 4: return COVT_DT_UINT_64;
    case 5: return COVT_DT_BOOLEAN;
    case 6: return COVT_DT_GEOMETRY;
    default: return 0xFFu;
    }
}

__device__ __forceinline__ uint64_t geometry_bytes(const Lite& lite)
{
    uint64_t n = 0;
#pragma unroll
    for (int s = COVT_SLOT_TYPES; s < COVT_NUM_SLOTS; s++) n += lite.has(s) ? lite.bl(s) : 0u;
    return n;
}

// id / geometry placement state shared by the two grammars
struct ColumnOrder {
    bool have_id = false, have_geom = false, id_first = false, nonstandard = false;
    uint32_t n_props = 0;
    // returns false on a duplicate id / geometry column (the reference keeps columns in a map keyed by name: one of each)
    __device__ __forceinline__ bool see(bool is_id, bool is_geom)
    {
        if (is_id) {
            if (have_id) return false;
            have_id = true;
            id_first = !have_geom;
            if (n_props) nonstandard = true;
        } else if (is_geom) {
            if (have_geom) return false;
            have_geom = true;
            if (n_props) nonstandard = true;
        } else n_props++;
        return true;
    }
};

__device__ __forceinline__ int gen2b_geometry_slot(const uint8_t* blob, uint64_t soff, uint32_t slen)
{
    if (name_is(blob, soff, slen, "geometry_types")) return COVT_SLOT_TYPES;
    if (name_is(blob, soff, slen, "geometry_offsets")) return COVT_SLOT_GEOM;
    if (name_is(blob, soff, slen, "part_offsets")) return COVT_SLOT_PART;
    if (name_is(blob, soff, slen, "ring_offsets")) return COVT_SLOT_RING;
    if (name_is(blob, soff, slen, "vertex_offsets")) return COVT_SLOT_VOFF;
    if (name_is(blob, soff, slen, "vertex_buffer")) return COVT_SLOT_VBUF;
    if (name_is(blob, soff, slen, "index_buffer")) return COVT_SLOT_INDEX;
    return -1;
}

// ---- gen-2b: string name, varint extent, numFeatures, numColumns | per column: string name, byte dataType,
// byte columnType, varint numStreams | per stream: string name, varint numValues, varint byteLength, byte StreamEncoding.
// Parses ONE layer at c.p and places its payload; on COVT_OK c.p is the first byte after the layer.
template <class Props>
__device__ uint32_t walk_layer_gen2b(Cursor& c, const Lite& lite, LayerHead& H, Props& props)
{
    const uint8_t* blob = c.b;
    lite.clear();
    H.layer_start = c.p;
    c_string(c, H.name_offset, H.name_length);
    H.extent = c_varint(c);
    H.num_features = c_varint(c);
    H.num_columns = c_varint(c);
    if (c.err) return COVT_ERR_TRUNCATED;
    const uint64_t cols_start = c.p;
    uint64_t prop_listed = 0;
    ColumnOrder ord;
    H.geom_ct = 0;
    for (uint32_t ci = 0; ci < H.num_columns; ci++) {
        uint64_t noff; uint32_t nlen;
        c_string(c, noff, nlen);
        (void)c_byte(c);  // gen-2 data type
        const uint32_t column_type = c_byte(c);
        const uint32_t num_streams = c_varint(c);
        if (c.err) return COVT_ERR_TRUNCATED;
        const bool is_id = name_is(blob, noff, nlen, "id");
        const bool is_geom = name_is(blob, noff, nlen, "geometry");
        if (ci == 0 && !is_id && !is_geom) return COVT_ERR_BAD_METADATA;  // CovtParser.java:67-69
        if (!ord.see(is_id, is_geom)) return COVT_ERR_BAD_METADATA;
        if (is_geom) {
            if (column_type > COVT_CT_ICE_MORTON_CODE) return COVT_ERR_BAD_METADATA;
            H.geom_ct = column_type;
        }
        for (uint32_t si = 0; si < num_streams; si++) {
            uint64_t soff; uint32_t slen;
            c_string(c, soff, slen);
            const uint32_t nv = c_varint(c);
            const uint32_t bl = c_varint(c);
            const uint32_t enc = c_byte(c);
            if (c.err) return COVT_ERR_TRUNCATED;
            int slot = -1;
            if (is_id) { if (name_is(blob, soff, slen, "data")) slot = COVT_SLOT_ID; }
            else if (is_geom) slot = gen2b_geometry_slot(blob, soff, slen);
            if (slot >= 0) {
                if (enc > COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_ERR_BAD_METADATA;
                if (!plausible_count(nv, bl)) return COVT_ERR_TRUNCATED;
                lite.set(slot, nv, bl, enc);
            } else if (is_id || is_geom) {
                return COVT_ERR_BAD_METADATA;
            } else {
                prop_listed += bl;
            }
        }
    }
    if (!ord.have_geom || !lite.has(COVT_SLOT_TYPES) || !lite.has(COVT_SLOT_VBUF)) return COVT_ERR_BAD_METADATA;
    const uint64_t meta_end = c.p;
    const uint64_t geom_bytes = geometry_bytes(lite);
    const uint64_t id_bytes = lite.has(COVT_SLOT_ID) ? lite.bl(COVT_SLOT_ID) : 0u;
    uint64_t p = meta_end;
    H.id_offset = H.geom_offset = p;
    if (!ord.nonstandard && !(ord.n_props && props.want())) {
        if (ord.id_first) { H.id_offset = p; p += id_bytes; H.geom_offset = p; p += geom_bytes; }
        else { H.geom_offset = p; p += geom_bytes; H.id_offset = p; p += id_bytes; }
        p += prop_listed;
    } else {
        // second pass: payloads in column-metadata order (the bytes parsed fine a moment ago)
        Cursor m = {blob, cols_start, meta_end, false};
        for (uint32_t ci = 0; ci < H.num_columns; ci++) {
            uint64_t noff; uint32_t nlen;
            c_string(m, noff, nlen);
            const uint32_t dt2 = c_byte(m);
            const uint32_t column_type = c_byte(m);
            const uint32_t num_streams = c_varint(m);
            const bool is_id = name_is(blob, noff, nlen, "id");
            const bool is_geom = name_is(blob, noff, nlen, "geometry");
            if (is_id) { H.id_offset = p; p += id_bytes; }
            else if (is_geom) { H.geom_offset = p; p += geom_bytes; }
            else props.begin_column(noff, nlen, dt_of_gen2(dt2), column_type, H.num_features);
            for (uint32_t si = 0; si < num_streams; si++) {
                uint64_t soff; uint32_t slen;
                c_string(m, soff, slen);
                const uint32_t nv = c_varint(m);
                const uint32_t bl = c_varint(m);
                const uint32_t enc = c_byte(m);
                if (is_id || is_geom) continue;
                if (props.want()) {
                    // stream names of property columns: present, data, length, dictionary; localized dictionaries list pairs
                    // (present_<s>, <s>) and share one length + dictionary (oracle/properties.py)
                    uint32_t st = COVT_ST_DATA;
                    uint64_t sub_off = 0; uint32_t sub_len = 0;
                    if (name_is(blob, soff, slen, "present")) st = COVT_ST_PRESENT;
                    else if (name_is(blob, soff, slen, "data")) st = COVT_ST_DATA;
                    else if (name_is(blob, soff, slen, "length")) st = COVT_ST_LENGTH;
                    else if (name_is(blob, soff, slen, "dictionary")) st = COVT_ST_DICTIONARY;
                    else if (name_starts(blob, soff, slen, "present_")) { st = COVT_ST_PRESENT; sub_off = soff + 8; sub_len = slen - 8; }
                    else { st = COVT_ST_DATA; sub_off = soff; sub_len = slen; }
                    props.stream(st, sub_off, sub_len, nv, bl, enc, p);
                }
                p += bl;
            }
            if (!is_id && !is_geom) props.end_column();
            if (p > c.end) return COVT_ERR_TRUNCATED;
        }
    }
    if (p > c.end) return COVT_ERR_TRUNCATED;
    c.p = p;
    return COVT_OK;
}

// one column header of the gen-3 grammar (CovtParser.java:604-624); false = BAD_METADATA
__device__ __forceinline__ bool gen3_column_header(Cursor& c, bool optimized, uint32_t ci, uint32_t n_fields, bool& is_id, bool& is_geom,
                                                   uint64_t& name_off, uint32_t& name_len, uint32_t& data_type, uint32_t& column_type)
{
    is_id = is_geom = false;
    name_off = 0;
    name_len = 0;
    if (optimized || ci == 0) {  // :604-614
        const uint32_t column_id = c_varint(c);
        if (column_id > 1) {
            if (!optimized || column_id - 2 >= n_fields) return false;  // fields == null -> NPE / IndexOutOfBounds
            name_off = column_id - 2;  // index into the TileJSON fields of the layer
        } else if (column_id == 0) is_id = true;
        else is_geom = true;
    } else {
        c_string(c, name_off, name_len);  // :616
        if (!c.err) { is_id = name_is(c.b, name_off, name_len, "id"); is_geom = name_is(c.b, name_off, name_len, "geometry"); }
    }
    const uint32_t column_desc = c_byte(c);  // :619-624
    data_type = (column_desc >> 3) & 0xFu;
    column_type = column_desc & 0x7u;
    return true;
}
// last stream of a column (:639-647). (INDEX_BUFFER, when present, precedes VERTEX_BUFFER in the metadata so that the
// reference terminator still ends the column.)
__device__ __forceinline__ bool gen3_last_stream(uint32_t data_type, uint32_t column_type, uint32_t stream_type)
{
    if (data_type == COVT_DT_GEOMETRY && stream_type == COVT_ST_VERTEX_BUFFER) return true;
    if (stream_type == COVT_ST_DATA && column_type == COVT_CT_PLAIN) return true;
    return stream_type == COVT_ST_DICTIONARY;
}

// ---- gen-3 = CovtParser.decodeLayerMetadata (CovtParser.java:574-652); no tile header, the caller loops until EOF (:56)
template <class Props>
__device__ uint32_t walk_layer_gen3(Cursor& c, const uint32_t* tj_fields, uint32_t tj_layers, const Lite& lite, LayerHead& H, Props& props)
{
    const uint8_t* blob = c.b;
    lite.clear();
    H.layer_start = c.p;
    const uint32_t header = c_byte(c);
    const bool optimized = header & 1u;  // :575-578
    uint32_t n_fields = 0;
    if (optimized) {
        const uint32_t layer_id = c_varint(c);  // :584-589
        if (c.err) return COVT_ERR_TRUNCATED;
        if (!tj_fields || layer_id >= tj_layers) return COVT_ERR_BAD_METADATA;
        n_fields = tj_fields[layer_id];
        H.name_offset = layer_id;
        H.name_length = 0;
    } else {
        c_string(c, H.name_offset, H.name_length);  // :592
    }
    H.extent = c_varint(c);  // :595-598
    H.num_features = c_varint(c);
    H.num_columns = c_varint(c);
    if (c.err) return COVT_ERR_TRUNCATED;
    const uint64_t cols_start = c.p;
    ColumnOrder ord;
    H.geom_ct = 0;
    for (uint32_t ci = 0; ci < H.num_columns; ci++) {
        bool is_id, is_geom;
        uint64_t noff; uint32_t nlen, data_type, column_type;
        if (!gen3_column_header(c, optimized, ci, n_fields, is_id, is_geom, noff, nlen, data_type, column_type)) return COVT_ERR_BAD_METADATA;
        if (c.err) return COVT_ERR_TRUNCATED;
        if (column_type > COVT_CT_ICE_MORTON_CODE) return COVT_ERR_BAD_METADATA;  // ColumnType.values()[..] throws
        if (ci == 0 && !is_id && !is_geom) return COVT_ERR_BAD_METADATA;           // :67-69
        if (!ord.see(is_id, is_geom)) return COVT_ERR_BAD_METADATA;
        if (is_geom) H.geom_ct = column_type;
        for (;;) {  // :628-648
            const uint32_t stream_desc = c_byte(c);
            const uint32_t stream_type = stream_desc >> 4;
            const uint32_t enc = stream_desc & 0xFu;
            const uint32_t nv = c_varint(c);
            const uint32_t bl = c_varint(c);
            if (c.err) return COVT_ERR_TRUNCATED;
            if (stream_type > COVT_ST_INDEX_BUFFER || enc > COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_ERR_BAD_METADATA;
            int slot = -1;
            if (is_id && stream_type == COVT_ST_DATA) slot = COVT_SLOT_ID;
            else if (is_geom && stream_type >= COVT_ST_GEOMETRY_TYPES && stream_type <= COVT_ST_VERTEX_BUFFER)
                slot = COVT_SLOT_TYPES + (int)(stream_type - COVT_ST_GEOMETRY_TYPES);
            else if (is_geom && stream_type == COVT_ST_INDEX_BUFFER) slot = COVT_SLOT_INDEX;
            if (slot >= 0) {
                if (!plausible_count(nv, bl)) return COVT_ERR_TRUNCATED;
                lite.set(slot, nv, bl, enc);
            } else if (is_id || is_geom) return COVT_ERR_BAD_METADATA;
            if (gen3_last_stream(data_type, column_type, stream_type)) break;
        }
    }
    if (!ord.have_geom || !lite.has(COVT_SLOT_TYPES) || !lite.has(COVT_SLOT_VBUF)) return COVT_ERR_BAD_METADATA;
    const uint64_t meta_end = c.p;
    const uint64_t geom_bytes = geometry_bytes(lite);
    const uint64_t id_bytes = lite.has(COVT_SLOT_ID) ? lite.bl(COVT_SLOT_ID) : 0u;
    uint64_t p = meta_end;
    H.id_offset = H.geom_offset = p;
    if (!ord.nonstandard && ord.n_props == 0) {
        if (ord.id_first) { H.id_offset = p; p += id_bytes; H.geom_offset = p; p += geom_bytes; }
        else { H.geom_offset = p; p += geom_bytes; H.id_offset = p; p += id_bytes; }
    } else {
        // second pass in column order. BOOLEAN = the listed data stream only (CovtParser.java:280-290); every other property
        // type = an unlisted Byte-RLE present stream of ceil(numFeatures / 8) bytes (:295) followed by its listed streams
        Cursor m = {blob, cols_start, meta_end, false};
        for (uint32_t ci = 0; ci < H.num_columns; ci++) {
            bool is_id, is_geom;
            uint64_t noff; uint32_t nlen, data_type, column_type;
            (void)gen3_column_header(m, optimized, ci, n_fields, is_id, is_geom, noff, nlen, data_type, column_type);
            const bool prop = !is_id && !is_geom;
            if (is_id) { H.id_offset = p; p += id_bytes; }
            else if (is_geom) { H.geom_offset = p; p += geom_bytes; }
            else {
                props.begin_column(noff, nlen, data_type, column_type, H.num_features);
                if (data_type != COVT_DT_BOOLEAN) {
                    uint64_t q = p;
                    if (!byte_rle_span(blob, q, c.end, (H.num_features + 7u) / 8u)) return COVT_ERR_TRUNCATED;
                    props.stream(COVT_ST_PRESENT, 0, 0, H.num_features, (uint32_t)(q - p), COVT_ENC_BOOLEAN_RLE, p);
                    p = q;
                }
            }
            for (;;) {
                const uint32_t stream_desc = c_byte(m);
                const uint32_t stream_type = stream_desc >> 4;
                const uint32_t nv = c_varint(m);
                const uint32_t bl = c_varint(m);
                if (prop) {
                    // BOOLEAN data = one bit per FEATURE whatever numValues says (CovtParser.java:280-283 reads ceil(numFeatures / 8) bytes)
                    props.stream(stream_type, 0, 0, (data_type == COVT_DT_BOOLEAN && stream_type == COVT_ST_DATA) ? H.num_features : nv, bl,
                                 stream_desc & 0xFu, p);
                    p += bl;
                }
                if (m.err || gen3_last_stream(data_type, column_type, stream_type)) break;
            }
            if (prop) props.end_column();
            if (p > c.end) return COVT_ERR_TRUNCATED;
        }
    }
    if (p > c.end) return COVT_ERR_TRUNCATED;
    c.p = p;
    return COVT_OK;
}

// One layer at c.p (either grammar); on COVT_OK c.p is the first byte after the layer's payload.
template <class Props>
__device__ __forceinline__ uint32_t walk_layer(Cursor& c, uint32_t container, const uint32_t* tj_fields, uint32_t tj_layers, const Lite& lite,
                                               LayerHead& H, Props& props)
{
    if (container == COVT_CONTAINER_GEN2B) return walk_layer_gen2b(c, lite, H, props);
    return walk_layer_gen3(c, tj_fields, tj_layers, lite, H, props);
}

// Walks one tile; on_layer(layer_index, lite, H) is called for every complete layer. Returns the tile status.
template <class Props, class OnLayer>
__device__ __forceinline__ uint32_t walk_tile(const uint8_t* blob, uint64_t begin, uint64_t end, uint32_t container, const uint32_t* tj_fields,
                                              uint32_t tj_layers, const Lite& lite, Props& props, OnLayer&& on_layer)
{
    Cursor c = {blob, begin, end, false};
    LayerHead H;
    if (container == COVT_CONTAINER_GEN2B) {
        (void)c_varint(c);  // version
        const uint32_t num_layers = c_varint(c);
        if (c.err) return COVT_ERR_TRUNCATED;
        for (uint32_t li = 0; li < num_layers; li++) {
            props.set_layer(li);
            const uint32_t st = walk_layer_gen2b(c, lite, H, props);
            if (st != COVT_OK) return st;
            on_layer(li, H);
        }
        return c.p == end ? COVT_OK : COVT_ERR_TRUNCATED;
    }
    uint32_t li = 0;
    while (c.p < end) {
        props.set_layer(li);
        const uint32_t st = walk_layer_gen3(c, tj_fields, tj_layers, lite, H, props);
        if (st != COVT_OK) return st;
        on_layer(li, H);
        li++;
    }
    return COVT_OK;
}

}  // namespace covt
