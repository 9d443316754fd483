/*
 * covt_gen.c — COVT encoders + synthetic tile generator (see covt_gen.h). Input synthesis only.
 *
 * Encoder restatements:
 *   varint/zigzag/delta   J/converter/EncodingUtils.java:39-114
 *   ORC RLE v1 writers    orc-core 1.8.1 RunLengthIntegerWriter / RunLengthByteWriter via EncodingUtils.java:123-147
 *   FastPFOR+VB           JavaFastPFOR 0.1.12 Composition(FastPFOR, VariableByte) via EncodingUtils.java:149-188
 *   stream selection      J/converter/CovtConverter.java:571-986
 */
#include "covt_gen.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * growable byte buffer
 * ---------------------------------------------------------------------------------------------- */
static int buf_reserve(covt_gen_buf* b, size_t extra)
{
    if (b->len + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 256;
    while (nc < b->len + extra) nc *= 2;
    uint8_t* nd = (uint8_t*)realloc(b->data, nc);
    if (!nd) return -1;
    b->data = nd;
    b->cap = nc;
    return 0;
}
static void buf_put(covt_gen_buf* b, const void* p, size_t n)
{
    if (buf_reserve(b, n)) return;
    memcpy(b->data + b->len, p, n);
    b->len += n;
}
static void buf_byte(covt_gen_buf* b, uint8_t v) { buf_put(b, &v, 1); }
/* EncodingUtils.putVarInt, EncodingUtils.java:105-114 (64-bit, logical shift) */
static void buf_varint(covt_gen_buf* b, uint64_t v)
{
    do {
        uint8_t bits = (uint8_t)(v & 0x7F);
        v >>= 7;
        buf_byte(b, (uint8_t)(bits | (v ? 0x80 : 0)));
    } while (v);
}
/* EncodingUtils.encodeString, :116-120 */
static void buf_string(covt_gen_buf* b, const char* s)
{
    size_t n = strlen(s);
    buf_varint(b, n);
    buf_put(b, s, n);
}
void covt_gen_buf_free(covt_gen_buf* b) { free(b->data); b->data = NULL; b->len = b->cap = 0; }
void covt_gen_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * varint / zigzag / delta                                       EncodingUtils.java:39-93
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t zz64(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
static inline uint32_t zz32(int32_t v) { return (uint32_t)(v >> 31) ^ ((uint32_t)v << 1); }

size_t covt_enc_varints(const int64_t* v, size_t n, int zigzag, int delta, uint8_t* out, size_t cap)
{
    size_t o = 0;
    int64_t prev = 0;
    for (size_t i = 0; i < n; i++) {
        int64_t x = v[i];
        if (delta) { int64_t d = (int64_t)((uint64_t)x - (uint64_t)prev); prev = x; x = d; }
        uint64_t u = zigzag ? zz64(x) : (uint64_t)x;
        do {
            if (o >= cap) return (size_t)-1;
            uint8_t bits = (uint8_t)(u & 0x7F);
            u >>= 7;
            out[o++] = (uint8_t)(bits | (u ? 0x80 : 0));
        } while (u);
    }
    return o;
}

void covt_enc_zigzag_delta_coordinates(const int32_t* xy, size_t n_ints, int32_t* out)
{
    int32_t px = 0, py = 0;
    for (size_t j = 0; j < n_ints; j++) {
        if ((j & 1) == 0) { out[j] = (int32_t)zz32((int32_t)((uint32_t)xy[j] - (uint32_t)px)); px = xy[j]; }
        else { out[j] = (int32_t)zz32((int32_t)((uint32_t)xy[j] - (uint32_t)py)); py = xy[j]; }
    }
}

/* GeometryUtils.encodeMorton, GeometryUtils.java:23-32 */
int32_t covt_enc_morton(int32_t x, int32_t y, uint32_t num_bits)
{
    int32_t tile_extent = (int32_t)((uint32_t)2 << ((num_bits - 2) & 31));
    x += tile_extent / 2;
    y += tile_extent / 2;
    uint32_t code = 0;
    for (uint32_t i = 0; i < num_bits; i++)
        code |= (((uint32_t)x & (1u << i)) << i) | (((uint32_t)y & (1u << i)) << (i + 1));
    return (int32_t)code;
}

/* ------------------------------------------------------------------------------------------------
 * ORC RLE v1 writers
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint8_t* out; size_t o, cap; int overflow; } sink_t;
static inline void sk_byte(sink_t* s, uint8_t v) { if (s->o < s->cap) s->out[s->o] = v; else s->overflow = 1; s->o++; }
static void sk_vulong(sink_t* s, uint64_t v)
{
    for (;;) {
        if ((v & ~(uint64_t)0x7f) == 0) { sk_byte(s, (uint8_t)v); return; }
        sk_byte(s, (uint8_t)(0x80 | (v & 0x7f)));
        v >>= 7;
    }
}

typedef struct { int64_t literals[128]; int num, repeat, tail; int64_t delta; int is_signed; sink_t* s; } rlew_t;
static void rlew_flush(rlew_t* w)
{
    if (w->num == 0) return;
    if (w->repeat) {
        sk_byte(w->s, (uint8_t)(w->num - 3));
        sk_byte(w->s, (uint8_t)(int8_t)w->delta);
        sk_vulong(w->s, w->is_signed ? zz64(w->literals[0]) : (uint64_t)w->literals[0]);
    } else {
        sk_byte(w->s, (uint8_t)(-w->num));
        for (int i = 0; i < w->num; i++) sk_vulong(w->s, w->is_signed ? zz64(w->literals[i]) : (uint64_t)w->literals[i]);
    }
    w->repeat = 0;
    w->num = 0;
    w->tail = 0;
}
static void rlew_write(rlew_t* w, int64_t v)
{
    if (w->num == 0) {
        w->literals[w->num++] = v;
        w->tail = 1;
    } else if (w->repeat) {
        if (v == (int64_t)((uint64_t)w->literals[0] + (uint64_t)w->delta * (uint64_t)w->num)) {
            w->num++;
            if (w->num == 130) rlew_flush(w);
        } else {
            rlew_flush(w);
            w->literals[w->num++] = v;
            w->tail = 1;
        }
    } else {
        if (w->tail == 1 || v != (int64_t)((uint64_t)w->literals[w->num - 1] + (uint64_t)w->delta)) {
            w->delta = (int64_t)((uint64_t)v - (uint64_t)w->literals[w->num - 1]);
            w->tail = (w->delta < -128 || w->delta > 127) ? 1 : 2;
        } else
            w->tail++;
        if (w->tail == 3) {
            if (w->num + 1 == 3) {
                w->repeat = 1;
                w->num++;
            } else {
                w->num -= 2;
                int64_t base = w->literals[w->num];
                rlew_flush(w);
                w->literals[0] = base;
                w->repeat = 1;
                w->num = 3;
            }
        } else {
            w->literals[w->num++] = v;
            if (w->num == 128) rlew_flush(w);
        }
    }
}

size_t covt_enc_rle(const int64_t* v, size_t n, int is_signed, uint8_t* out, size_t cap)
{
    sink_t s = {out, 0, cap, 0};
    rlew_t w;
    memset(&w, 0, sizeof(w));
    w.is_signed = is_signed;
    w.s = &s;
    for (size_t i = 0; i < n; i++) rlew_write(&w, v[i]);
    rlew_flush(&w);
    return s.overflow ? (size_t)-1 : s.o;
}

size_t covt_enc_byte_rle(const uint8_t* v, size_t n, uint8_t* out, size_t cap)
{
    sink_t s = {out, 0, cap, 0};
    uint8_t lit[128];
    int num = 0, repeat = 0, tail = 0;
#define BFLUSH()                                                                    \
    do {                                                                            \
        if (num) {                                                                  \
            if (repeat) { sk_byte(&s, (uint8_t)(num - 3)); sk_byte(&s, lit[0]); }   \
            else { sk_byte(&s, (uint8_t)(-num)); for (int i_ = 0; i_ < num; i_++) sk_byte(&s, lit[i_]); } \
            repeat = 0; num = 0; tail = 0;                                          \
        }                                                                           \
    } while (0)
    for (size_t i = 0; i < n; i++) {
        uint8_t x = v[i];
        if (num == 0) { lit[num++] = x; tail = 1; }
        else if (repeat) {
            if (x == lit[0]) { num++; if (num == 130) BFLUSH(); }
            else { BFLUSH(); lit[num++] = x; tail = 1; }
        } else {
            if (x == lit[num - 1]) tail++; else tail = 1;
            if (tail == 3) {
                if (num + 1 == 3) { repeat = 1; num++; }
                else { num -= 2; BFLUSH(); lit[0] = x; repeat = 1; num = 3; }
            } else { lit[num++] = x; if (num == 128) BFLUSH(); }
        }
    }
    BFLUSH();
#undef BFLUSH
    return s.overflow ? (size_t)-1 : s.o;
}

/* ------------------------------------------------------------------------------------------------
 * Composition(FastPFOR, VariableByte).compress, words serialised big-endian
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t* w; size_t n, cap; } words_t;
static void w_push(words_t* W, uint32_t v)
{
    if (W->n == W->cap) { W->cap = W->cap ? W->cap * 2 : 1024; W->w = (uint32_t*)realloc(W->w, W->cap * sizeof(uint32_t)); }
    W->w[W->n++] = v;
}
static inline uint32_t bits32(uint32_t v) { return v ? 32u - (uint32_t)__builtin_clz(v) : 0u; }

/* BitPacking.fastpack: 32 values -> `bit` words */
static void fastpack32(const uint32_t* in, words_t* W, uint32_t bit)
{
    if (bit == 0) return;
    uint64_t acc = 0;
    uint32_t fill = 0;
    uint32_t mask = bit == 32 ? 0xFFFFFFFFu : ((1u << bit) - 1u);
    for (int i = 0; i < 32; i++) {
        acc |= (uint64_t)(in[i] & mask) << fill;
        fill += bit;
        if (fill >= 32) { w_push(W, (uint32_t)acc); acc >>= 32; fill -= 32; }
    }
}

typedef struct { uint32_t* v; size_t n, cap; } u32vec_t;
static void v_push(u32vec_t* V, uint32_t x)
{
    if (V->n == V->cap) { V->cap = V->cap ? V->cap * 2 : 64; V->v = (uint32_t*)realloc(V->v, V->cap * sizeof(uint32_t)); }
    V->v[V->n++] = x;
}

static void fastpfor_encode_page(const uint32_t* in, uint32_t thissize, words_t* W)
{
    size_t headerpos = W->n;
    w_push(W, 0);
    u32vec_t exc[33];
    memset(exc, 0, sizeof(exc));
    covt_gen_buf bc = {0, 0, 0};
    for (uint32_t blk = 0; blk + 256 <= thissize; blk += 256) {
        const uint32_t* d = in + blk;
        /* getBestBFromData */
        int freqs[33];
        memset(freqs, 0, sizeof(freqs));
        for (int k = 0; k < 256; k++) freqs[bits32(d[k])]++;
        int bestb = 32;
        while (freqs[bestb] == 0) bestb--;
        int maxb = bestb;
        int bestcost = bestb * 256;
        int cexcept = 0, bestc = 0;
        for (int b = bestb - 1; b >= 0; b--) {
            cexcept += freqs[b + 1];
            if (cexcept == 256) break;
            int thiscost = cexcept * 8 + cexcept * (maxb - b) + b * 256 + 8;
            if (maxb - b == 1) thiscost -= cexcept;
            if (thiscost < bestcost) { bestcost = thiscost; bestb = b; bestc = cexcept; }
        }
        buf_byte(&bc, (uint8_t)bestb);
        buf_byte(&bc, (uint8_t)bestc);
        if (bestc > 0) {
            buf_byte(&bc, (uint8_t)maxb);
            int index = maxb - bestb;
            for (int k = 0; k < 256; k++) {
                if ((bestb == 32 ? 0 : (d[k] >> bestb)) != 0) {
                    buf_byte(&bc, (uint8_t)k);
                    v_push(&exc[index], d[k] >> bestb);
                }
            }
        }
        for (int k = 0; k < 256; k += 32) fastpack32(d + k, W, (uint32_t)bestb);
    }
    W->w[headerpos] = (uint32_t)(W->n - headerpos);
    uint32_t bytesize = (uint32_t)bc.len;
    while (bc.len & 3) buf_byte(&bc, 0);
    w_push(W, bytesize);
    for (size_t i = 0; i < bc.len; i += 4)
        w_push(W, (uint32_t)bc.data[i] | ((uint32_t)bc.data[i + 1] << 8) | ((uint32_t)bc.data[i + 2] << 16) | ((uint32_t)bc.data[i + 3] << 24));
    uint32_t bitmap = 0;
    for (int k = 2; k <= 32; k++) if (exc[k].n) bitmap |= 1u << (k - 1);
    w_push(W, bitmap);
    for (int k = 2; k <= 32; k++) {
        if (!exc[k].n) continue;
        w_push(W, (uint32_t)exc[k].n);
        size_t size = exc[k].n;
        size_t start = W->n;
        uint32_t grp[32];
        for (size_t j = 0; j < size; j += 32) {
            for (int q = 0; q < 32; q++) grp[q] = j + q < size ? exc[k].v[j + q] : 0;
            fastpack32(grp, W, (uint32_t)k);
        }
        W->n = start + (size * k + 31) / 32; /* tmpoutpos -= overflow * k / 32 */
    }
    for (int k = 0; k <= 32; k++) free(exc[k].v);
    free(bc.data);
}

size_t covt_enc_fastpfor(const int32_t* v, size_t n, int zigzag, int delta, uint8_t* out, size_t cap)
{
    uint32_t* enc = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    int32_t prev = 0;
    for (size_t i = 0; i < n; i++) {
        int32_t x = v[i];
        if (delta) { int32_t d = (int32_t)((uint32_t)x - (uint32_t)prev); prev = x; x = d; }
        enc[i] = zigzag ? zz32(x) : (uint32_t)x;
    }
    words_t W = {0, 0, 0};
    size_t n256 = n / 256 * 256;
    if (n256 > 0) {
        w_push(&W, (uint32_t)n256);
        for (size_t p = 0; p < n256; p += 65536) {
            uint32_t thissize = (uint32_t)(n256 - p < 65536 ? n256 - p : 65536);
            fastpfor_encode_page(enc + p, thissize, &W);
        }
    } else if (n > 0)
        w_push(&W, 0); /* Composition writes a literal 0 when FastPFOR emitted nothing (inlength == 0 returns earlier) */
    /* VariableByte tail: 7 bits per byte LSB-first, MSB set on the LAST byte, zero padded to a word, bytes LE in words */
    if (n > n256) {
        covt_gen_buf vb = {0, 0, 0};
        for (size_t i = n256; i < n; i++) {
            uint32_t val = enc[i];
            while (val >= 128) { buf_byte(&vb, (uint8_t)(val & 127)); val >>= 7; }
            buf_byte(&vb, (uint8_t)(val | 128));
        }
        while (vb.len & 3) buf_byte(&vb, 0);
        for (size_t i = 0; i < vb.len; i += 4)
            w_push(&W, (uint32_t)vb.data[i] | ((uint32_t)vb.data[i + 1] << 8) | ((uint32_t)vb.data[i + 2] << 16) | ((uint32_t)vb.data[i + 3] << 24));
        free(vb.data);
    }
    free(enc);
    size_t bytes = W.n * 4;
    if (bytes > cap) { free(W.w); return (size_t)-1; }
    for (size_t i = 0; i < W.n; i++) { /* EncodingUtils.java:174-185: big-endian */
        uint32_t x = W.w[i];
        out[4 * i] = (uint8_t)(x >> 24);
        out[4 * i + 1] = (uint8_t)(x >> 16);
        out[4 * i + 2] = (uint8_t)(x >> 8);
        out[4 * i + 3] = (uint8_t)x;
    }
    free(W.w);
    return bytes;
}

/* ------------------------------------------------------------------------------------------------
 * layer writer (CovtConverter stream selection)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint8_t* data; size_t len; uint8_t encoding; uint32_t num_values; int present; } enc_stream_t;

static uint8_t* scratch_alloc(size_t n) { return (uint8_t*)malloc(n + 64); }

/* CovtConverter.addOffsets, :899-920 */
static void encode_offsets(const int32_t* v, uint32_t n, uint32_t options, enc_stream_t* s)
{
    s->present = n > 0;
    if (!n) return;
    int64_t* l = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    for (uint32_t i = 0; i < n; i++) l[i] = v[i];
    size_t cap = (size_t)n * 11 + 64;
    uint8_t* rle = scratch_alloc(cap);
    size_t rl = covt_enc_rle(l, n, 0, rle, cap);
    free(l);
    s->num_values = n;
    if (!(options & COVT_GEN_ALLOW_PFOR_TOPOLOGY) || (options & COVT_GEN_FORCE_RLE_TOPOLOGY)) {
        s->data = rle; s->len = rl; s->encoding = 5;
        return;
    }
    size_t pcap = (size_t)n * 5 + 4096;
    uint8_t* pf = scratch_alloc(pcap);
    size_t pl = covt_enc_fastpfor(v, n, 1, 1, pf, pcap);
    if (pl <= rl) { s->data = pf; s->len = pl; s->encoding = 9; free(rle); }
    else { s->data = rle; s->len = rl; s->encoding = 5; free(pf); }
}

static int cmp_u32(const void* a, const void* b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}

static void meta_stream_gen2b(covt_gen_buf* m, const char* name, const enc_stream_t* s)
{
    buf_string(m, name);
    buf_varint(m, s->num_values);
    buf_varint(m, s->len);
    buf_byte(m, s->encoding);
}
/* CovtConverter.addOptimizedStreamMetadata, :478-483 */
static void meta_stream_gen3(covt_gen_buf* m, uint32_t stream_type, const enc_stream_t* s)
{
    buf_byte(m, (uint8_t)(stream_type << 4 | s->encoding));
    buf_varint(m, s->num_values);
    buf_varint(m, s->len);
}

int32_t covt_gen_begin_tile(covt_gen_buf* tile, uint32_t container, uint32_t num_layers)
{
    if (container == 0) { buf_varint(tile, 1); buf_varint(tile, num_layers); }
    return 0;
}

#define COVT_GEN_OPTIMIZED_METADATA 0x40u

int32_t covt_gen_append_layer(covt_gen_buf* tile, const covt_gen_layer* L, uint32_t container, uint32_t options)
{
    enc_stream_t id = {0}, types = {0}, geom = {0}, part = {0}, ring = {0}, voff = {0}, vbuf = {0}, idx = {0};
    uint32_t num_bits = 32 - (L->extent ? (uint32_t)__builtin_clz(L->extent) : 32);
    uint8_t column_type = 0;

    /* geometry_types: always Byte-RLE (CovtConverter.convertTopologyStreams, :876-879) */
    types.present = 1;
    types.num_values = L->n_features;
    types.data = scratch_alloc((size_t)L->n_features * 2 + 16);
    types.len = covt_enc_byte_rle(L->types, L->n_features, types.data, (size_t)L->n_features * 2 + 16);
    types.encoding = 7;
    encode_offsets(L->geom_counts, L->n_geom, options, &geom);
    encode_offsets(L->part_counts, L->n_part, options, &part);
    encode_offsets(L->ring_counts, L->n_ring, options, &ring);

    const uint32_t nv = L->n_vertices;
    const int allow_pfor_v = (options & COVT_GEN_ALLOW_PFOR_VERTEX) && !(options & COVT_GEN_FORCE_VARINT_VERTEX);
    if (options & COVT_GEN_ICE_MORTON) {
        /* CovtConverter.convertIceCodedGeometryColumn :671-769 + encodeVertexBuffer :771-856 (Morton branch only) */
        column_type = 4;
        uint32_t* codes = (uint32_t*)malloc(((size_t)nv + 1) * sizeof(uint32_t));
        for (uint32_t i = 0; i < nv; i++) codes[i] = (uint32_t)covt_enc_morton(L->xy[2 * i], L->xy[2 * i + 1], num_bits);
        uint32_t* dict = (uint32_t*)malloc(((size_t)nv + 1) * sizeof(uint32_t));
        memcpy(dict, codes, (size_t)nv * sizeof(uint32_t));
        qsort(dict, nv, sizeof(uint32_t), cmp_u32);
        uint32_t nd = 0;
        for (uint32_t i = 0; i < nv; i++) if (i == 0 || dict[i] != dict[i - 1]) dict[nd++] = dict[i];
        int32_t* offs = (int32_t*)malloc(((size_t)nv + 1) * sizeof(int32_t));
        for (uint32_t i = 0; i < nv; i++) {
            uint32_t lo = 0, hi = nd;
            while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (dict[mid] < codes[i]) lo = mid + 1; else hi = mid; }
            offs[i] = (int32_t)lo;
        }
        /* vertex_offsets: varint(zz,delta) if strictly shorter than FastPFOR(zz,delta), :813-820 */
        {
            int64_t* l = (int64_t*)malloc(((size_t)nv + 1) * sizeof(int64_t));
            for (uint32_t i = 0; i < nv; i++) l[i] = offs[i];
            size_t cap = (size_t)nv * 10 + 64;
            uint8_t* vi = scratch_alloc(cap);
            size_t vl = covt_enc_varints(l, nv, 1, 1, vi, cap);
            free(l);
            voff.present = 1;
            voff.num_values = nv;
            if (allow_pfor_v) {
                size_t pcap = (size_t)nv * 5 + 4096;
                uint8_t* pf = scratch_alloc(pcap);
                size_t pl = covt_enc_fastpfor(offs, nv, 1, 1, pf, pcap);
                if (vl < pl) { voff.data = vi; voff.len = vl; voff.encoding = 4; free(pf); }
                else { voff.data = pf; voff.len = pl; voff.encoding = 9; free(vi); }
            } else { voff.data = vi; voff.len = vl; voff.encoding = 4; }
        }
        /* vertex_buffer: sorted Morton codes, delta WITHOUT zigzag (:939-948); numValues = #vertices (:853-854) */
        {
            int64_t* l = (int64_t*)malloc(((size_t)nd + 1) * sizeof(int64_t));
            for (uint32_t i = 0; i < nd; i++) l[i] = (int64_t)dict[i];
            size_t cap = (size_t)nd * 10 + 64;
            uint8_t* vi = scratch_alloc(cap);
            size_t vl = covt_enc_varints(l, nd, 0, 1, vi, cap);
            free(l);
            vbuf.present = 1;
            vbuf.num_values = nd;
            if (allow_pfor_v) {
                size_t pcap = (size_t)nd * 5 + 4096;
                uint8_t* pf = scratch_alloc(pcap);
                size_t pl = covt_enc_fastpfor((const int32_t*)dict, nd, 0, 1, pf, pcap);
                if (vl < pl) { vbuf.data = vi; vbuf.len = vl; vbuf.encoding = 4; free(pf); }
                else { vbuf.data = pf; vbuf.len = pl; vbuf.encoding = 9; free(vi); }
            } else { vbuf.data = vi; vbuf.len = vl; vbuf.encoding = 4; }
        }
        free(codes);
        free(dict);
        free(offs);
    } else {
        /* CovtConverter.convertUnorderedGeometryColumn :641-668: numValues = #ints; FastPFOR if <= varint */
        column_type = 0;
        size_t ni = (size_t)nv * 2;
        int32_t* zz = (int32_t*)malloc((ni + 1) * sizeof(int32_t));
        covt_enc_zigzag_delta_coordinates(L->xy, ni, zz);
        int64_t* l = (int64_t*)malloc((ni + 1) * sizeof(int64_t));
        for (size_t i = 0; i < ni; i++) l[i] = (int64_t)(uint32_t)zz[i];
        size_t cap = ni * 5 + 64;
        uint8_t* vi = scratch_alloc(cap);
        size_t vl = covt_enc_varints(l, ni, 0, 0, vi, cap);
        free(l);
        vbuf.present = 1;
        vbuf.num_values = (uint32_t)ni;
        if (allow_pfor_v) {
            size_t pcap = ni * 5 + 4096;
            uint8_t* pf = scratch_alloc(pcap);
            size_t pl = covt_enc_fastpfor(zz, ni, 0, 0, pf, pcap);
            if (pl <= vl) { vbuf.data = pf; vbuf.len = pl; vbuf.encoding = 9; free(vi); }
            else { vbuf.data = vi; vbuf.len = vl; vbuf.encoding = 4; free(pf); }
        } else { vbuf.data = vi; vbuf.len = vl; vbuf.encoding = 4; }
        free(zz);
    }

    if (L->ids) {
        /* CovtConverter.convertIdColumn :546-569 without its mislabel bug (:564-565) */
        uint32_t n = L->n_features;
        size_t cap = (size_t)n * 11 + 64;
        id.present = 1;
        id.num_values = n;
        uint8_t* dv = scratch_alloc(cap);
        size_t dl = covt_enc_varints(L->ids, n, 1, 1, dv, cap);
        if (options & COVT_GEN_ID_DELTA_VARINT) { id.data = dv; id.len = dl; id.encoding = 4; }
        else {
            uint8_t* rl = scratch_alloc(cap);
            size_t rlen = covt_enc_rle(L->ids, n, 0, rl, cap);
            uint8_t* vi = scratch_alloc(cap);
            size_t vl = covt_enc_varints(L->ids, n, 0, 0, vi, cap);
            if (rlen < vl && rlen < dl) { id.data = rl; id.len = rlen; id.encoding = 5; free(vi); free(dv); }
            else if (dl < vl) { id.data = dv; id.len = dl; id.encoding = 4; free(vi); free(rl); }
            else { id.data = vi; id.len = vl; id.encoding = 1; free(rl); free(dv); }
        }
    }
    if (L->index_buffer && L->n_index) {
        /* extension: UInt32 stream, FAST_PFOR_DELTA_ZIG_ZAG like the topology streams */
        size_t pcap = (size_t)L->n_index * 5 + 4096;
        idx.present = 1;
        idx.num_values = L->n_index;
        idx.data = scratch_alloc(pcap);
        idx.len = covt_enc_fastpfor(L->index_buffer, L->n_index, 1, 1, idx.data, pcap);
        idx.encoding = 9;
    }

    /* ---- metadata ---- */
    covt_gen_buf m = {0, 0, 0};
    uint32_t num_columns = 1 + (id.present ? 1 : 0);
    uint32_t n_geom_streams = 2 + geom.present + part.present + ring.present + voff.present + idx.present;
    if (container == 0) {
        buf_string(&m, L->name);
        buf_varint(&m, L->extent);
        buf_varint(&m, L->n_features);
        buf_varint(&m, num_columns);
        if (id.present) {
            buf_string(&m, "id");
            buf_byte(&m, 4); /* gen-2 data type UINT_64 */
            buf_byte(&m, 0);
            buf_varint(&m, 1);
            meta_stream_gen2b(&m, "data", &id);
        }
        buf_string(&m, "geometry");
        buf_byte(&m, 6); /* gen-2 data type GEOMETRY */
        buf_byte(&m, column_type);
        buf_varint(&m, n_geom_streams);
        if (voff.present) { meta_stream_gen2b(&m, "vertex_offsets", &voff); meta_stream_gen2b(&m, "vertex_buffer", &vbuf); }
        meta_stream_gen2b(&m, "geometry_types", &types);
        if (geom.present) meta_stream_gen2b(&m, "geometry_offsets", &geom);
        if (part.present) meta_stream_gen2b(&m, "part_offsets", &part);
        if (ring.present) meta_stream_gen2b(&m, "ring_offsets", &ring);
        if (!voff.present) meta_stream_gen2b(&m, "vertex_buffer", &vbuf);
        if (idx.present) meta_stream_gen2b(&m, "index_buffer", &idx);
    } else {
        /* CovtConverter.convertLayerMetadata :383-426 / convertOptimizedLayerMetadata :300-353 */
        int optimized = (options & COVT_GEN_OPTIMIZED_METADATA) != 0;
        buf_byte(&m, (uint8_t)(1 << 1 | (optimized ? 1 : 0)));
        if (optimized) buf_varint(&m, (uint64_t)strtoul(L->name, NULL, 10)); /* layerId */
        else buf_string(&m, L->name);
        buf_varint(&m, L->extent);
        buf_varint(&m, L->n_features);
        buf_varint(&m, num_columns);
        int col = 0;
        if (id.present) {
            buf_varint(&m, 0); /* column id 0 = "id" */
            buf_byte(&m, (uint8_t)(4 << 3 | 0)); /* UINT_64, PLAIN */
            meta_stream_gen3(&m, 1, &id);
            col++;
        }
        if (optimized || col == 0) buf_varint(&m, 1); /* column id 1 = "geometry" */
        else buf_string(&m, "geometry");
        buf_byte(&m, (uint8_t)(8 << 3 | column_type)); /* GEOMETRY */
        meta_stream_gen3(&m, 4, &types);
        if (geom.present) meta_stream_gen3(&m, 5, &geom);
        if (part.present) meta_stream_gen3(&m, 6, &part);
        if (ring.present) meta_stream_gen3(&m, 7, &ring);
        if (voff.present) meta_stream_gen3(&m, 8, &voff);
        if (idx.present) meta_stream_gen3(&m, 12, &idx);
        meta_stream_gen3(&m, 9, &vbuf);
    }
    buf_put(tile, m.data, m.len);
    free(m.data);
    /* ---- payload: [id] | types, geometry_offsets, part_offsets, ring_offsets, vertex_offsets, vertex_buffer [, index] ---- */
    enc_stream_t* order[8] = {&id, &types, &geom, &part, &ring, &voff, &vbuf, &idx};
    for (int i = 0; i < 8; i++) {
        if (order[i]->present) buf_put(tile, order[i]->data, order[i]->len);
        free(order[i]->data);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * PRNG + distributions
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t* r) /* xorshift64* */
{
    uint64_t x = r->s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1DULL;
}
static inline void rng_seed(rng_t* r, uint64_t seed)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL; /* splitmix64 scramble so that seed 0 works */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    r->s = (z ^ (z >> 31)) | 1;
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32); }
static uint32_t rng_geometric(rng_t* r, double mean)
{
    if (mean <= 0) return 0;
    double p = 1.0 / (1.0 + mean);
    double u = rng_unit(r);
    if (u <= 0) u = 1e-300;
    double k = floor(log(u) / log(1.0 - p));
    return k > 1e6 ? 1000000u : (uint32_t)k;
}

/* Config 3: exactly target_bytes of x/y zigzag-delta varints, even value count */
static inline uint64_t emit_varint_of_length(rng_t* r, uint8_t* out, uint64_t o, uint32_t L)
{
    uint32_t lo = L == 1 ? 0 : 1u << (7 * (L - 1));
    uint32_t hi = 1u << (7 * L);
    uint32_t v = lo + rng_below(r, hi - lo);
    for (uint32_t i = 0; i < L; i++) {
        uint8_t b = (uint8_t)(v & 0x7f);
        v >>= 7;
        out[o++] = (uint8_t)(b | (i + 1 < L ? 0x80 : 0));
    }
    return o;
}
static inline uint32_t draw_length(rng_t* r)
{
    double u = rng_unit(r);
    return u < 0.531 ? 1 : u < 0.957 ? 2 : u < 0.997 ? 3 : 4;
}
uint64_t covt_gen_varint_stream(uint8_t* out, uint64_t target_bytes, uint64_t seed)
{
    rng_t r;
    rng_seed(&r, seed);
    uint64_t o = 0, n = 0;
    if (target_bytes < 2) return 0;
    while (target_bytes - o > 16) { /* a pair uses at most 8 bytes, so at least 8 are left for the closing pairs */
        o = emit_varint_of_length(&r, out, o, draw_length(&r));
        o = emit_varint_of_length(&r, out, o, draw_length(&r));
        n += 2;
    }
    uint64_t left = target_bytes - o;
    while (left > 8) { /* 1-byte pairs until one last pair can absorb the remainder */
        o = emit_varint_of_length(&r, out, o, 1);
        o = emit_varint_of_length(&r, out, o, 1);
        n += 2;
        left -= 2;
    }
    uint32_t a = (uint32_t)(left - 1 < 4 ? left - 1 : 4), b = (uint32_t)left - a;
    o = emit_varint_of_length(&r, out, o, a);
    o = emit_varint_of_length(&r, out, o, b);
    n += 2;
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Config 5: synthetic mixed-geometry tiles
 * ---------------------------------------------------------------------------------------------- */
void covt_gen_default_params(covt_gen_params* p)
{
    memset(p, 0, sizeof(*p));
    p->layers_per_tile = 2;
    p->mean_features = 48;
    p->p_point = 0.12; p->p_line = 0.70; p->p_polygon = 0.15; p->p_multiline = 0.015; p->p_multipolygon = 0.015;
    p->mean_line_extra = 6;
    p->mean_ring_extra = 5;
    p->p_second_ring = 0.10;
    p->extent = 4096;
    p->container = 0;
    p->with_ids = 1;
    p->with_index_buffer = 0;
    p->max_step = 48;
}

typedef struct { int32_t* v; size_t n, cap; } i32vec_t;
static void iv_push(i32vec_t* V, int32_t x)
{
    if (V->n == V->cap) { V->cap = V->cap ? V->cap * 2 : 256; V->v = (int32_t*)realloc(V->v, V->cap * sizeof(int32_t)); }
    V->v[V->n++] = x;
}
typedef struct {
    i32vec_t geom, part, ring, xy, index;
    uint8_t* types; size_t types_cap;
    int64_t* ids; size_t ids_cap;
} layer_scratch_t;

static void walk(rng_t* r, const covt_gen_params* p, uint32_t n, i32vec_t* xy, covt_gen_truth* t, int closed)
{
    int32_t lo = -(int32_t)(p->extent / 8), hi = (int32_t)(p->extent + p->extent / 8) - 1;
    int32_t x = (int32_t)rng_below(r, p->extent), y = (int32_t)rng_below(r, p->extent);
    int32_t x0 = x, y0 = y;
    for (uint32_t i = 0; i < n; i++) {
        if (i) {
            x += (int32_t)rng_below(r, 2 * p->max_step + 1) - (int32_t)p->max_step;
            y += (int32_t)rng_below(r, 2 * p->max_step + 1) - (int32_t)p->max_step;
            if (x < lo) x = lo; if (x > hi) x = hi;
            if (y < lo) y = lo; if (y > hi) y = hi;
        }
        iv_push(xy, x);
        iv_push(xy, y);
        t->sum_x += x; t->sum_y += y;
        t->sum_x_closed += x; t->sum_y_closed += y;
    }
    t->vertices += n;
    if (closed && n > 0) { t->sum_x_closed += x0; t->sum_y_closed += y0; }
}

static void gen_polygon(rng_t* r, const covt_gen_params* p, layer_scratch_t* S, covt_gen_truth* t)
{
    uint32_t nr = rng_unit(r) < p->p_second_ring ? 2 : 1;
    iv_push(&S->part, (int32_t)nr);
    for (uint32_t k = 0; k < nr; k++) {
        uint32_t n = 3 + rng_geometric(r, p->mean_ring_extra);
        iv_push(&S->ring, (int32_t)n);
        if (p->with_index_buffer) { /* fan triangulation (0,i,i+1) offset by the ring's first vertex */
            int32_t base = (int32_t)(S->xy.n / 2);
            for (uint32_t i = 1; i + 1 < n; i++) { iv_push(&S->index, base); iv_push(&S->index, base + (int32_t)i); iv_push(&S->index, base + (int32_t)i + 1); }
        }
        walk(r, p, n, &S->xy, t, 1);
        t->rings++;
        t->polygon_rings++;
    }
    t->parts++;
}

static void gen_tile(uint64_t tile_index, const covt_gen_params* p, layer_scratch_t* S, covt_gen_buf* out, covt_gen_truth* t)
{
    rng_t r;
    rng_seed(&r, tile_index);
    covt_gen_begin_tile(out, p->container == 0 ? 0 : 1, p->layers_per_tile);
    for (uint32_t li = 0; li < p->layers_per_tile; li++) {
        S->geom.n = S->part.n = S->ring.n = S->xy.n = S->index.n = 0;
        uint32_t F = 1 + rng_geometric(&r, p->mean_features - 1);
        if (F > S->types_cap) { S->types_cap = F * 2; S->types = (uint8_t*)realloc(S->types, S->types_cap); }
        if (F > S->ids_cap) { S->ids_cap = F * 2; S->ids = (int64_t*)realloc(S->ids, S->ids_cap * sizeof(int64_t)); }
        /* ids: even tiles -> run-friendly (RLE wins); odd tiles -> large irregular 64-bit ids (genuine delta varint) */
        int64_t idv = (tile_index & 1) ? (int64_t)(rng_next(&r) >> 20) : (int64_t)(1 + rng_below(&r, 1000));
        for (uint32_t f = 0; f < F; f++) {
            double u = rng_unit(&r);
            uint8_t type;
            if (u < p->p_point) type = 0;
            else if (u < p->p_point + p->p_line) type = 1;
            else if (u < p->p_point + p->p_line + p->p_polygon) type = 2;
            else if (u < p->p_point + p->p_line + p->p_polygon + p->p_multiline) type = 4;
            else type = 5;
            S->types[f] = type;
            if (tile_index & 1) idv += (int64_t)rng_below(&r, 1u << 20) - (1 << 18);
            else idv += 1;
            S->ids[f] = idv;
            switch (type) {
            case 0: walk(&r, p, 1, &S->xy, t, 0); t->parts++; t->rings++; break;
            case 1: {
                uint32_t n = 2 + rng_geometric(&r, p->mean_line_extra);
                iv_push(&S->part, (int32_t)n);
                walk(&r, p, n, &S->xy, t, 0);
                t->parts++; t->rings++;
                break;
            }
            case 2: gen_polygon(&r, p, S, t); break;
            case 4: {
                uint32_t nl = 2 + rng_geometric(&r, 1.0);
                iv_push(&S->geom, (int32_t)nl);
                for (uint32_t k = 0; k < nl; k++) {
                    uint32_t n = 2 + rng_geometric(&r, p->mean_line_extra);
                    iv_push(&S->part, (int32_t)n);
                    walk(&r, p, n, &S->xy, t, 0);
                    t->parts++; t->rings++;
                }
                break;
            }
            default: {
                uint32_t np = 2 + rng_geometric(&r, 1.0);
                iv_push(&S->geom, (int32_t)np);
                for (uint32_t k = 0; k < np; k++) gen_polygon(&r, p, S, t);
                break;
            }
            }
        }
        t->features += F;
        char name[32];
        int optimized = p->container == 2;
        if (optimized) { name[0] = (char)('0' + (li % 10)); name[1] = 0; }
        else { memcpy(name, "layer_", 6); name[6] = (char)('a' + (li % 26)); name[7] = 0; }
        covt_gen_layer L;
        memset(&L, 0, sizeof(L));
        L.name = name;
        L.extent = p->extent;
        L.n_features = F;
        L.types = S->types;
        L.geom_counts = S->geom.v; L.n_geom = (uint32_t)S->geom.n;
        L.part_counts = S->part.v; L.n_part = (uint32_t)S->part.n;
        L.ring_counts = S->ring.v; L.n_ring = (uint32_t)S->ring.n;
        L.xy = S->xy.v; L.n_vertices = (uint32_t)(S->xy.n / 2);
        L.ids = p->with_ids ? S->ids : NULL;
        L.index_buffer = S->index.n ? S->index.v : NULL;
        L.n_index = (uint32_t)S->index.n;
        uint32_t options = COVT_GEN_ALLOW_PFOR_TOPOLOGY | COVT_GEN_ALLOW_PFOR_VERTEX;
        if (li & 1) options |= COVT_GEN_ICE_MORTON;           /* ICE_MORTON on odd layers, PLAIN on even */
        if (tile_index & 1) options |= COVT_GEN_ID_DELTA_VARINT; /* genuine 64-bit delta-varint ids on odd tiles */
        if (optimized) options |= COVT_GEN_OPTIMIZED_METADATA;
        covt_gen_append_layer(out, &L, p->container == 0 ? 0 : 1, options);
    }
}

typedef struct {
    uint64_t first_tile; uint32_t n_tiles; const covt_gen_params* p;
    atomic_uint next; uint32_t chunk;
    covt_gen_buf* chunk_bufs; uint64_t* tile_sizes; /* [n_tiles] */
    covt_gen_truth* truths; /* per thread */
} gen_shared_t;
typedef struct { gen_shared_t* sh; uint32_t thread; } gen_thread_t;

static void* gen_worker(void* arg)
{
    gen_thread_t* T = (gen_thread_t*)arg;
    gen_shared_t* sh = T->sh;
    layer_scratch_t S;
    memset(&S, 0, sizeof(S));
    covt_gen_truth* t = &sh->truths[T->thread];
    for (;;) {
        uint32_t c = atomic_fetch_add(&sh->next, 1u);
        uint64_t b = (uint64_t)c * sh->chunk;
        if (b >= sh->n_tiles) break;
        uint64_t e = b + sh->chunk < sh->n_tiles ? b + sh->chunk : sh->n_tiles;
        covt_gen_buf* out = &sh->chunk_bufs[c];
        for (uint64_t i = b; i < e; i++) {
            size_t before = out->len;
            gen_tile(sh->first_tile + i, sh->p, &S, out, t);
            sh->tile_sizes[i] = out->len - before;
        }
    }
    free(S.geom.v); free(S.part.v); free(S.ring.v); free(S.xy.v); free(S.index.v); free(S.types); free(S.ids);
    return NULL;
}

int32_t covt_gen_tiles(uint64_t first_tile, uint32_t n_tiles, const covt_gen_params* p, uint32_t n_threads,
                       uint8_t** blob, uint64_t* blob_len, uint64_t* tile_offsets, covt_gen_truth* truth)
{
    if (n_threads == 0) { long n = sysconf(_SC_NPROCESSORS_ONLN); n_threads = n < 1 ? 1u : (uint32_t)n; }
    if (n_threads > 256) n_threads = 256;
    gen_shared_t sh;
    memset(&sh, 0, sizeof(sh));
    sh.first_tile = first_tile;
    sh.n_tiles = n_tiles;
    sh.p = p;
    sh.chunk = 256;
    atomic_init(&sh.next, 0);
    uint32_t n_chunks = (n_tiles + sh.chunk - 1) / sh.chunk;
    sh.chunk_bufs = (covt_gen_buf*)calloc((size_t)n_chunks + 1, sizeof(covt_gen_buf));
    sh.tile_sizes = (uint64_t*)calloc((size_t)n_tiles + 1, sizeof(uint64_t));
    sh.truths = (covt_gen_truth*)calloc(n_threads, sizeof(covt_gen_truth));
    pthread_t th[256];
    gen_thread_t ta[256];
    uint32_t started = 0;
    for (uint32_t i = 1; i < n_threads; i++) {
        ta[i].sh = &sh; ta[i].thread = i;
        if (pthread_create(&th[i], NULL, gen_worker, &ta[i]) != 0) break;
        started = i;
    }
    ta[0].sh = &sh; ta[0].thread = 0;
    gen_worker(&ta[0]);
    for (uint32_t i = 1; i <= started; i++) pthread_join(th[i], NULL);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_tiles; i++) { tile_offsets[i] = total; total += sh.tile_sizes[i]; }
    tile_offsets[n_tiles] = total;
    uint8_t* out = (uint8_t*)malloc(total + 64);
    if (!out) return -1;
    uint64_t o = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        memcpy(out + o, sh.chunk_bufs[c].data, sh.chunk_bufs[c].len);
        o += sh.chunk_bufs[c].len;
        free(sh.chunk_bufs[c].data);
    }
    memset(out + total, 0, 64);
    if (truth) {
        memset(truth, 0, sizeof(*truth));
        for (uint32_t i = 0; i < n_threads; i++) {
            truth->features += sh.truths[i].features; truth->vertices += sh.truths[i].vertices;
            truth->parts += sh.truths[i].parts; truth->rings += sh.truths[i].rings;
            truth->polygon_rings += sh.truths[i].polygon_rings;
            truth->sum_x += sh.truths[i].sum_x; truth->sum_y += sh.truths[i].sum_y;
            truth->sum_x_closed += sh.truths[i].sum_x_closed; truth->sum_y_closed += sh.truths[i].sum_y_closed;
        }
    }
    free(sh.chunk_bufs);
    free(sh.tile_sizes);
    free(sh.truths);
    *blob = out;
    *blob_len = total;
    return 0;
}
