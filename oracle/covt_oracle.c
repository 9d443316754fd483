/*
 * covt_oracle.c — CPU ORACLE (test infrastructure; see covt_oracle.h for the parity-pinning statement).
 *
 * Scalar, allocation-light C restatement of the reference Java decode path. Citations:
 *   J/  = /root/reference/evaluation/java/src/main/java/com/covt/
 *   JS/ = /root/reference/parser/js/
 * Third-party algorithms restated from SURVEY.md §A.4 (ORC RLE v1, orc-core 1.8.1) and §A.5
 * (JavaFastPFOR 0.1.12: FastPFOR block 256 / page 65536 composed with VariableByte).
 */
#include "covt_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for (no OpenMP dependency) ------------------------------------------ */
typedef void (*pf_body_t)(void* arg, int64_t begin, int64_t end, uint32_t thread);
typedef struct { pf_body_t body; void* arg; int64_t n, chunk; atomic_llong next; } pf_shared_t;
typedef struct { pf_shared_t* sh; uint32_t thread; } pf_thread_t;
static void* pf_worker(void* p)
{
    pf_thread_t* t = (pf_thread_t*)p;
    for (;;) {
        int64_t b = atomic_fetch_add(&t->sh->next, t->sh->chunk);
        if (b >= t->sh->n) break;
        int64_t e = b + t->sh->chunk < t->sh->n ? b + t->sh->chunk : t->sh->n;
        t->sh->body(t->sh->arg, b, e, t->thread);
    }
    return NULL;
}
static uint32_t pf_default_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1u : (uint32_t)n;
}
static void parallel_for(int64_t n, int64_t chunk, uint32_t n_threads, pf_body_t body, void* arg)
{
    if (n_threads == 0) n_threads = pf_default_threads();
    if (n_threads > 256) n_threads = 256;
    pf_shared_t sh;
    sh.body = body; sh.arg = arg; sh.n = n; sh.chunk = chunk < 1 ? 1 : chunk;
    atomic_init(&sh.next, 0);
    pthread_t th[256];
    pf_thread_t ta[256];
    uint32_t started = 0;
    for (uint32_t i = 1; i < n_threads; i++) {
        ta[i].sh = &sh; ta[i].thread = i;
        if (pthread_create(&th[i], NULL, pf_worker, &ta[i]) != 0) break;
        started = i;
    }
    ta[0].sh = &sh; ta[0].thread = 0;
    pf_worker(&ta[0]);
    for (uint32_t i = 1; i <= started; i++) pthread_join(th[i], NULL);
}

/* Java semantics helpers: int shifts mask the distance to 5 bits, long shifts to 6 bits. */
static inline int32_t jshl32(int32_t v, int s) { return (int32_t)((uint32_t)v << (s & 31)); }
static inline int32_t jadd32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int64_t jadd64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }

/* ---------------------------------------------------------------------------------------------
 * Varint / zigzag / delta                                   J/decoder/DecodingUtils.java:35-254
 * ------------------------------------------------------------------------------------------- */

/* private decodeVarint(byte[],int,int[],int), DecodingUtils.java:157-186: at most 4 bytes, the 4th
 * byte's continuation bit is ignored. Returns 0 on success, -1 when the buffer ends first (Java:
 * ArrayIndexOutOfBoundsException). */
static inline int varint_java(const uint8_t* buf, uint64_t len, uint64_t* pos, int32_t* value, int* overlong)
{
    uint64_t p = *pos;
    int32_t v = 0;
    for (int i = 0; i < 4; i++) {
        if (p >= len) return -1;
        uint8_t b = buf[p++];
        v |= (int32_t)(b & 0x7f) << (7 * i);
        if ((b & 0x80) == 0) break;
        if (i == 3 && overlong) *overlong = 1;
    }
    *value = v;
    *pos = p;
    return 0;
}

/* decodeZigZag, DecodingUtils.java:252-254 */
static inline int32_t zigzag32(int32_t e) { return (int32_t)(((uint32_t)e >> 1) ^ (uint32_t)(-(e & 1))); }
static inline int64_t zigzag64(uint64_t e) { return (int64_t)((e >> 1) ^ (uint64_t)(-(int64_t)(e & 1))); }

/* decodeVarint(byte[], IntWrapper, int), DecodingUtils.java:35-44 */
int32_t covt_oracle_decode_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong)
{
    for (uint32_t i = 0; i < n; i++)
        if (varint_java(buf, len, pos, &out[i], overlong)) return COVT_ERR_TRUNCATED;
    return COVT_OK;
}

/* decodeZigZagVarint(byte[], IntWrapper, int), DecodingUtils.java:46-53 */
int32_t covt_oracle_decode_zigzag_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong)
{
    for (uint32_t i = 0; i < n; i++) {
        int32_t v;
        if (varint_java(buf, len, pos, &v, overlong)) return COVT_ERR_TRUNCATED;
        out[i] = zigzag32(v);
    }
    return COVT_OK;
}

/* decodeZigZagDeltaVarint, DecodingUtils.java:55-66 */
int32_t covt_oracle_decode_zigzag_delta_varint(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong)
{
    int32_t prev = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t v;
        if (varint_java(buf, len, pos, &v, overlong)) return COVT_ERR_TRUNCATED;
        prev = jadd32(prev, zigzag32(v));
        out[i] = prev;
    }
    return COVT_OK;
}

/* decodeZigZagDeltaVarintCoordinates, DecodingUtils.java:95-112 (the loop steps by 2 and always reads
 * a pair, so an odd numValues would write out[n]; the oracle reports that as a count mismatch). */
int32_t covt_oracle_decode_zigzag_delta_varint_coordinates(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int32_t* out, int* overlong)
{
    if (n & 1) return COVT_ERR_COUNT_MISMATCH;
    int32_t px = 0, py = 0;
    for (uint32_t i = 0; i < n; i += 2) {
        int32_t dx, dy;
        if (varint_java(buf, len, pos, &dx, overlong)) return COVT_ERR_TRUNCATED;
        if (varint_java(buf, len, pos, &dy, overlong)) return COVT_ERR_TRUNCATED;
        px = jadd32(px, zigzag32(dx));
        py = jadd32(py, zigzag32(dy));
        out[i] = px;
        out[i + 1] = py;
    }
    return COVT_OK;
}

/* GeometryUtils.decodeMortonCode, GeometryUtils.java:41-47: coordinate |= (code & (1L << 2i)) >> i with
 * `code` sign-extended to long and the compound assignment narrowing back to int. */
static inline int32_t morton_compact_java(int32_t code, uint32_t num_bits)
{
    int32_t c = 0;
    int64_t lc = (int64_t)code;
    for (uint32_t i = 0; i < num_bits; i++) {
        int64_t bit = lc & (int64_t)((uint64_t)1 << ((2 * i) & 63));
        c = (int32_t)((int64_t)c | (bit >> (i & 63)));
    }
    return c;
}

/* GeometryUtils.decodeMorton, GeometryUtils.java:34-39. no_shift reproduces the older converter that
 * wrote Morton codes without the extent/2 shift (SURVEY §0-8d, §A.6 MORTON_SHIFT). */
void covt_oracle_decode_morton(int32_t code, uint32_t num_bits, int no_shift, int32_t* x, int32_t* y)
{
    int32_t cx = morton_compact_java(code, num_bits);
    int32_t cy = morton_compact_java(code >> 1, num_bits);
    if (no_shift) {
        /* fixture-era converter (omt zoom 8): raw two's-complement coordinates were interleaved, so the
         * num_bits-bit value is sign-extended instead of un-shifted (verified against the partner .mvt). */
        if (num_bits >= 1 && num_bits < 32) {
            int sh = 32 - (int)num_bits;
            cx = (int32_t)((uint32_t)cx << sh) >> sh;
            cy = (int32_t)((uint32_t)cy << sh) >> sh;
        }
        *x = cx;
        *y = cy;
        return;
    }
    int32_t tile_extent = jshl32(2, (int)num_bits - 2);
    int32_t half = tile_extent / 2;
    *x = (int32_t)((uint32_t)cx - (uint32_t)half);
    *y = (int32_t)((uint32_t)cy - (uint32_t)half);
}

/* decodeDeltaVarintMortonCodes, DecodingUtils.java:394-409: UNSIGNED varint deltas (no zigzag). */
int32_t covt_oracle_decode_delta_varint_morton_codes(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n_vertices,
                                                     uint32_t num_bits, int no_shift, int32_t* out, int* overlong)
{
    int32_t prev = 0;
    for (uint32_t i = 0; i < n_vertices; i++) {
        int32_t d;
        if (varint_java(buf, len, pos, &d, overlong)) return COVT_ERR_TRUNCATED;
        prev = jadd32(prev, d);
        covt_oracle_decode_morton(prev, num_bits, no_shift, &out[2 * i], &out[2 * i + 1]);
    }
    return COVT_OK;
}

/* Full 64-bit LEB128 as written by EncodingUtils.putVarInt (EncodingUtils.java:105-114) and read by
 * orc SerializationUtils.readVulong: up to 10 bytes; overlong when the 10th byte still continues. */
static inline int varint64(const uint8_t* buf, uint64_t len, uint64_t* pos, uint64_t* value, int* overlong)
{
    uint64_t p = *pos, v = 0;
    int shift = 0;
    for (int i = 0; i < 10; i++) {
        if (p >= len) return -1;
        uint8_t b = buf[p++];
        v |= (uint64_t)(b & 0x7f) << (shift & 63);
        shift += 7;
        if ((b & 0x80) == 0) break;
        if (i == 9 && overlong) *overlong = 1;
    }
    *value = v;
    *pos = p;
    return 0;
}

int32_t covt_oracle_decode_varint64(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int64_t* out, int* overlong)
{
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v;
        if (varint64(buf, len, pos, &v, overlong)) return COVT_ERR_TRUNCATED;
        out[i] = (int64_t)v;
    }
    return COVT_OK;
}

/* inverse of EncodingUtils.encodeVarints(ids, zigZag=true, delta=true) (EncodingUtils.java:39-55,65-83) */
int32_t covt_oracle_decode_zigzag_delta_varint64(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int64_t* out, int* overlong)
{
    int64_t prev = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v;
        if (varint64(buf, len, pos, &v, overlong)) return COVT_ERR_TRUNCATED;
        prev = jadd64(prev, zigzag64(v));
        out[i] = prev;
    }
    return COVT_OK;
}

/* ---------------------------------------------------------------------------------------------
 * ORC RLE v1 (orc-core 1.8.1 RunLengthIntegerReader / RunLengthByteReader)
 *   call sites: DecodingUtils.java:257-306; in-repo twin JS/src/decoder/decodingUtils.ts:230-401
 * ------------------------------------------------------------------------------------------- */

/* decodeRle, DecodingUtils.java:257-272. Control byte c < 0x80: run of c+3 values base + i*delta
 * (delta = signed byte, base = (zigzag) LEB128); else literal group of 256-c (zigzag) LEB128 values.
 * The orc reader materialises a whole literal group on its header, so a group that straddles
 * numValues is consumed entirely; *pos ends after the last header/group touched. (The Java code
 * re-encodes to find the size, :308-310; equal on every fixture stream, tests/test_oracle_fixtures.py.) */
int32_t covt_oracle_decode_rle(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, int is_signed, int64_t* out)
{
    uint64_t p = *pos;
    uint32_t done = 0;
    while (done < n) {
        if (p >= len) return COVT_ERR_TRUNCATED;
        uint8_t c = buf[p++];
        if (c < 0x80) {
            uint32_t run = (uint32_t)c + 3;
            if (p >= len) return COVT_ERR_TRUNCATED;
            int64_t delta = (int8_t)buf[p++];
            uint64_t raw;
            if (varint64(buf, len, &p, &raw, NULL)) return COVT_ERR_TRUNCATED;
            int64_t base = is_signed ? zigzag64(raw) : (int64_t)raw;
            for (uint32_t i = 0; i < run && done < n; i++)
                out[done++] = (int64_t)((uint64_t)base + (uint64_t)i * (uint64_t)delta);
        } else {
            uint32_t lit = 256u - c;
            for (uint32_t i = 0; i < lit; i++) {
                uint64_t raw;
                if (varint64(buf, len, &p, &raw, NULL)) return COVT_ERR_TRUNCATED;
                if (done < n) out[done++] = is_signed ? zigzag64(raw) : (int64_t)raw;
            }
        }
    }
    *pos = p;
    return COVT_OK;
}

/* decodeByteRle, DecodingUtils.java:275-306 */
int32_t covt_oracle_decode_byte_rle(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint8_t* out)
{
    uint64_t p = *pos;
    uint32_t done = 0;
    while (done < n) {
        if (p >= len) return COVT_ERR_TRUNCATED;
        uint8_t c = buf[p++];
        if (c < 0x80) {
            uint32_t run = (uint32_t)c + 3;
            if (p >= len) return COVT_ERR_TRUNCATED;
            uint8_t v = buf[p++];
            for (uint32_t i = 0; i < run && done < n; i++) out[done++] = v;
        } else {
            uint32_t lit = 256u - c;
            if (p + lit > len) return COVT_ERR_TRUNCATED;
            for (uint32_t i = 0; i < lit; i++) {
                uint8_t v = buf[p++];
                if (done < n) out[done++] = v;
            }
        }
    }
    *pos = p;
    return COVT_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Composition(FastPFOR, VariableByte)  (JavaFastPFOR 0.1.12; call sites DecodingUtils.java:317-333)
 * ------------------------------------------------------------------------------------------- */

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* BitPacking.fastunpack: 32 values of `bit` bits from `bit` words, value i at bits [i*bit,(i+1)*bit)
 * of the LSB-first concatenation. w = big-endian word array base (bytes). */
static void fastunpack32(const uint8_t* w, uint32_t word_pos, int32_t* out, uint32_t bit)
{
    if (bit == 0) { memset(out, 0, 32 * sizeof(int32_t)); return; }
    if (bit == 32) { for (int i = 0; i < 32; i++) out[i] = (int32_t)be32(w + 4 * (uint64_t)(word_pos + i)); return; }
    uint32_t mask = (1u << bit) - 1u;
    for (uint32_t i = 0; i < 32; i++) {
        uint32_t bo = i * bit, wi = bo >> 5, sh = bo & 31;
        uint64_t lo = be32(w + 4 * (uint64_t)(word_pos + wi));
        if (sh + bit > 32) lo |= (uint64_t)be32(w + 4 * (uint64_t)(word_pos + wi + 1)) << 32;
        out[i] = (int32_t)((uint32_t)(lo >> sh) & mask);
    }
}

/* FastPFOR.decodePage (block 256, page 65536) — SURVEY §A.5. Returns 0 or a covt_status. */
static int32_t fastpfor_decode_page(const uint8_t* w, uint32_t n_words, uint32_t* inpos, int32_t* out, uint32_t thissize)
{
    uint32_t initpos = *inpos;
    if (initpos >= n_words) return COVT_ERR_TRUNCATED;
    uint32_t wheremeta = be32(w + 4 * (uint64_t)initpos);
    uint64_t inexcept = (uint64_t)initpos + wheremeta;
    if (inexcept >= n_words) return COVT_ERR_TRUNCATED;
    uint32_t bytesize = be32(w + 4 * inexcept);
    inexcept++;
    uint64_t bc_words = ((uint64_t)bytesize + 3) / 4;
    if (inexcept + bc_words > n_words) return COVT_ERR_TRUNCATED;
    uint64_t bc_word0 = inexcept; /* byte container: bytes little-endian inside each big-endian-serialised word */
    inexcept += bc_words;
    if (inexcept >= n_words) return COVT_ERR_TRUNCATED;
    uint32_t bitmap = be32(w + 4 * inexcept);
    inexcept++;

    /* exception arrays, widths k = 2..32: `size` k-bit values, ceil(size*k/32) words each */
    int32_t* exc[33];
    uint32_t exc_size[33], exc_ptr[33];
    memset(exc, 0, sizeof(exc));
    memset(exc_size, 0, sizeof(exc_size));
    memset(exc_ptr, 0, sizeof(exc_ptr));
    int32_t rc = COVT_OK;
    for (uint32_t k = 2; k <= 32 && rc == COVT_OK; k++) {
        if (!(bitmap & (1u << (k - 1)))) continue;
        if (inexcept >= n_words) { rc = COVT_ERR_TRUNCATED; break; }
        uint32_t size = be32(w + 4 * inexcept);
        inexcept++;
        uint64_t need = ((uint64_t)size * k + 31) / 32;
        if (inexcept + need > n_words) { rc = COVT_ERR_TRUNCATED; break; }
        uint32_t rounded = (size + 31) & ~31u;
        exc[k] = (int32_t*)malloc(((size_t)rounded + 32) * sizeof(int32_t));
        exc_size[k] = size;
        /* unpack whole groups of 32; the last group may read (but not use) words past `need`, so pad */
        uint8_t tmp[4 * 33];
        for (uint32_t j = 0; j < size; j += 32) {
            uint64_t wp = inexcept + (uint64_t)(j / 32) * k;
            uint64_t avail = n_words - wp;
            if (avail >= k) fastunpack32(w, (uint32_t)wp, exc[k] + j, k);
            else {
                memset(tmp, 0, sizeof(tmp));
                memcpy(tmp, w + 4 * wp, (size_t)avail * 4);
                fastunpack32(tmp, 0, exc[k] + j, k);
            }
        }
        inexcept += need;
    }

    uint32_t tmpin = initpos + 1;
    uint32_t bcpos = 0;
#define BC_GET(dst)                                                                                   \
    do {                                                                                              \
        if (bcpos >= bytesize) { rc = COVT_ERR_TRUNCATED; goto done; }                                \
        uint32_t word_ = be32(w + 4 * (bc_word0 + (bcpos >> 2)));                                     \
        (dst) = (uint8_t)(word_ >> (8 * (bcpos & 3)));                                                \
        bcpos++;                                                                                      \
    } while (0)

    if (rc != COVT_OK) goto done;
    for (uint32_t run = 0, run_end = thissize / 256; run < run_end; run++) {
        int32_t* o = out + (size_t)run * 256;
        uint8_t b, cexcept;
        BC_GET(b);
        BC_GET(cexcept);
        if (b > 32) { rc = COVT_ERR_BAD_METADATA; goto done; }
        if ((uint64_t)tmpin + 8ull * b > (uint64_t)initpos + wheremeta) { rc = COVT_ERR_TRUNCATED; goto done; }
        for (uint32_t k = 0; k < 256; k += 32) {
            fastunpack32(w, tmpin, o + k, b);
            tmpin += b;
        }
        if (cexcept > 0) {
            uint8_t maxbits;
            BC_GET(maxbits);
            int index = (int)maxbits - (int)b;
            if (index == 1) {
                for (uint32_t k = 0; k < cexcept; k++) {
                    uint8_t ppos;
                    BC_GET(ppos);
                    o[ppos] |= jshl32(1, b);
                }
            } else {
                if (index < 2 || index > 32 || !exc[index]) { rc = COVT_ERR_BAD_METADATA; goto done; }
                for (uint32_t k = 0; k < cexcept; k++) {
                    uint8_t ppos;
                    BC_GET(ppos);
                    if (exc_ptr[index] >= exc_size[index]) { rc = COVT_ERR_TRUNCATED; goto done; }
                    int32_t ev = exc[index][exc_ptr[index]++];
                    o[ppos] |= jshl32(ev, b);
                }
            }
        }
    }
    *inpos = (uint32_t)inexcept;
done:
#undef BC_GET
    for (int k = 2; k <= 32; k++) free(exc[k]);
    return rc;
}

int32_t covt_oracle_fastpfor_uncompress(const uint8_t* buf, uint32_t byte_length, uint32_t n, int32_t* out)
{
    uint32_t n_words = byte_length / 4; /* (int)Math.ceil(byteLength / 4) with integer division, DecodingUtils.java:324 */
    memset(out, 0, (size_t)n * sizeof(int32_t));
    if (n_words == 0) return COVT_OK; /* Composition.uncompress: inlength == 0 -> return */
    uint32_t inpos = 0, outpos = 0;
    /* FastPFOR.uncompress */
    uint32_t mynvalue = be32(buf);
    inpos = 1;
    if (mynvalue > n || (mynvalue & 255u)) return COVT_ERR_COUNT_MISMATCH;
    while (outpos != mynvalue) {
        uint32_t thissize = mynvalue - outpos < 65536u ? mynvalue - outpos : 65536u;
        int32_t rc = fastpfor_decode_page(buf, n_words, &inpos, out + outpos, thissize);
        if (rc != COVT_OK) return rc;
        outpos += thissize;
    }
    /* VariableByte.uncompress over all remaining words: 7 bits per byte LSB-first, MSB SET marks the
     * last byte of a value, bytes little-endian inside each word, zero bytes are padding. */
    int32_t v = 0;
    int shift = 0;
    int run = 0, overlong = 0; /* consecutive non-final bytes */
    for (uint32_t p = inpos; p < n_words; p++) {
        uint32_t val = be32(buf + 4 * (uint64_t)p);
        for (int s = 0; s < 32; s += 8) {
            uint8_t c = (uint8_t)(val >> s);
            v = jadd32(v, jshl32((int32_t)(c & 127), shift));
            if (c & 128) {
                if (outpos >= n) return COVT_ERR_COUNT_MISMATCH; /* Java: ArrayIndexOutOfBounds */
                out[outpos++] = v;
                v = 0;
                shift = 0;
                run = 0;
            } else {
                shift += 7;
                if (++run >= 5) overlong = 1;
            }
        }
    }
    if (outpos != n) return COVT_ERR_COUNT_MISMATCH; /* Java would silently leave zeros; flagged instead */
    /* Library policy shared with the product: VariableByte.uncompress has no length cap - a 6th byte is added at shift
     * 35 mod 32 = 3 - so five consecutive non-final bytes make the reference decode garbage; flagged instead of imitated. */
    if (overlong) return COVT_ERR_VARINT_OVERLONG;
    return COVT_OK;
}

/* decodeFastPfor128ZigZagDelta, DecodingUtils.java:316-347 */
int32_t covt_oracle_decode_fastpfor_zigzag_delta(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint32_t byte_length, int32_t* out)
{
    if (*pos + byte_length > len) return COVT_ERR_TRUNCATED;
    int32_t rc = covt_oracle_fastpfor_uncompress(buf + *pos, byte_length, n, out);
    if (rc != COVT_OK) return rc;
    int32_t prev = 0;
    for (uint32_t i = 0; i < n; i++) {
        prev = jadd32(prev, zigzag32(out[i]));
        out[i] = prev;
    }
    *pos += byte_length;
    return COVT_OK;
}

/* decodeFastPfor128DeltaCoordinates, DecodingUtils.java:349-392 */
int32_t covt_oracle_decode_fastpfor_delta_coordinates(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n, uint32_t byte_length, int32_t* out)
{
    if (n & 1) return COVT_ERR_COUNT_MISMATCH;
    if (*pos + byte_length > len) return COVT_ERR_TRUNCATED;
    int32_t rc = covt_oracle_fastpfor_uncompress(buf + *pos, byte_length, n, out);
    if (rc != COVT_OK) return rc;
    int32_t px = 0, py = 0;
    for (uint32_t i = 0; i < n; i += 2) {
        px = jadd32(px, zigzag32(out[i]));
        py = jadd32(py, zigzag32(out[i + 1]));
        out[i] = px;
        out[i + 1] = py;
    }
    *pos += byte_length;
    return COVT_OK;
}

/* decodeFastPfor128DeltaMortonCodes, DecodingUtils.java:411-444 (no zigzag) */
int32_t covt_oracle_decode_fastpfor_delta_morton_codes(const uint8_t* buf, uint64_t len, uint64_t* pos, uint32_t n_vertices,
                                                       uint32_t byte_length, uint32_t num_bits, int no_shift, int32_t* out)
{
    if (*pos + byte_length > len) return COVT_ERR_TRUNCATED;
    /* decode into the upper half so that the in-place expansion to (x,y) never overtakes the source */
    int32_t* codes = out + n_vertices;
    int32_t rc = covt_oracle_fastpfor_uncompress(buf + *pos, byte_length, n_vertices, codes);
    if (rc != COVT_OK) return rc;
    int32_t prev = 0;
    for (uint32_t i = 0; i < n_vertices; i++) {
        prev = jadd32(prev, codes[i]);
        int32_t x, y;
        covt_oracle_decode_morton(prev, num_bits, no_shift, &x, &y);
        out[2 * i] = x;
        out[2 * i + 1] = y;
    }
    *pos += byte_length;
    return COVT_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Dispatch                                  CovtParser.decodeGeometryColumn :392-511, decodedIds :552-572
 * ------------------------------------------------------------------------------------------- */
int32_t covt_oracle_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags)
{
    switch (stream_type) {
    case COVT_ST_GEOMETRY_TYPES: /* always Byte-RLE whatever the label, CovtParser.java:405-406 */
        return COVT_OP_BYTE_RLE;
    case COVT_ST_GEOMETRY_OFFSETS:
    case COVT_ST_PART_OFFSETS:
    case COVT_ST_RING_OFFSETS: /* :412-462 */
        if (encoding == COVT_ENC_RLE) return COVT_OP_RLE_U32;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA;
        return COVT_OP_NONE;
    case COVT_ST_VERTEX_OFFSETS: /* :464-477 */
    case COVT_ST_INDEX_BUFFER:   /* extension: decoded like a topology stream */
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_ZZ_DELTA;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA;
        return COVT_OP_NONE;
    case COVT_ST_VERTEX_BUFFER: /* :479-510 */
        if (column_type == COVT_CT_ICE_MORTON_CODE) {
            if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_DELTA_MORTON;
            if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_DELTA_MORTON;
            return COVT_OP_NONE;
        }
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_ZZ_DELTA_XY;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA_XY;
        return COVT_OP_NONE;
    case COVT_ST_DATA: /* id column, :552-572 */
        if (encoding == COVT_ENC_RLE) return COVT_OP_RLE_U64;
        if (encoding == COVT_ENC_VARINT) return (flags & COVT_FLAG_ID_WIDTH_32) ? COVT_OP_VARINT_U32_AS_I64 : COVT_OP_VARINT_U64;
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) {
            if (flags & COVT_FLAG_ID_DVZZ_IS_RLE) return COVT_OP_RLE_U64;
            return (flags & COVT_FLAG_ID_WIDTH_32) ? COVT_OP_VARINT_ZZ_DELTA_AS_I64 : COVT_OP_VARINT_ZZ_DELTA_64;
        }
        return COVT_OP_NONE;
    default:
        return COVT_OP_NONE;
    }
}

uint32_t covt_oracle_op_elem_size(uint32_t op)
{
    switch (op) {
    case COVT_OP_BYTE_RLE: return 1;
    case COVT_OP_RLE_U64: case COVT_OP_RLE_S64: case COVT_OP_VARINT_U64: case COVT_OP_VARINT_ZZ_DELTA_64:
    case COVT_OP_VARINT_U32_AS_I64: case COVT_OP_VARINT_ZZ_DELTA_AS_I64: case COVT_OP_VARINT_ZZ_AS_I64: return 8;
    default: return 4;
    }
}

uint64_t covt_oracle_op_out_count(uint32_t op, uint32_t num_values)
{
    if (op == COVT_OP_VARINT_DELTA_MORTON || op == COVT_OP_PFOR_DELTA_MORTON) return 2ull * num_values;
    return num_values;
}

static int32_t decode_op(const uint8_t* blob, uint64_t limit, uint64_t* pos, uint32_t op, uint32_t n, uint32_t byte_length,
                         uint32_t num_bits, uint32_t flags, void* out)
{
    int overlong = 0;
    int32_t rc;
    int no_shift = (flags & COVT_FLAG_MORTON_NO_SHIFT) != 0;
    switch (op) {
    case COVT_OP_BYTE_RLE: rc = covt_oracle_decode_byte_rle(blob, limit, pos, n, (uint8_t*)out); break;
    case COVT_OP_RLE_U64: rc = covt_oracle_decode_rle(blob, limit, pos, n, 0, (int64_t*)out); break;
    case COVT_OP_RLE_S64: rc = covt_oracle_decode_rle(blob, limit, pos, n, 1, (int64_t*)out); break;
    case COVT_OP_RLE_U32: {
        /* Arrays.stream(decodeRle(...)).mapToInt(i -> (int)i), CovtParser.java:419-420 */
        int64_t* tmp = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
        rc = covt_oracle_decode_rle(blob, limit, pos, n, 0, tmp);
        if (rc == COVT_OK) for (uint32_t i = 0; i < n; i++) ((int32_t*)out)[i] = (int32_t)tmp[i];
        free(tmp);
        break;
    }
    case COVT_OP_VARINT_U32: rc = covt_oracle_decode_varint(blob, limit, pos, n, (int32_t*)out, &overlong); break;
    case COVT_OP_VARINT_ZZ: rc = covt_oracle_decode_zigzag_varint(blob, limit, pos, n, (int32_t*)out, &overlong); break;
    case COVT_OP_VARINT_ZZ_DELTA: rc = covt_oracle_decode_zigzag_delta_varint(blob, limit, pos, n, (int32_t*)out, &overlong); break;
    case COVT_OP_VARINT_ZZ_DELTA_XY: rc = covt_oracle_decode_zigzag_delta_varint_coordinates(blob, limit, pos, n, (int32_t*)out, &overlong); break;
    case COVT_OP_VARINT_DELTA_MORTON: rc = covt_oracle_decode_delta_varint_morton_codes(blob, limit, pos, n, num_bits, no_shift, (int32_t*)out, &overlong); break;
    case COVT_OP_VARINT_U64: rc = covt_oracle_decode_varint64(blob, limit, pos, n, (int64_t*)out, &overlong); break;
    case COVT_OP_VARINT_ZZ_DELTA_64: rc = covt_oracle_decode_zigzag_delta_varint64(blob, limit, pos, n, (int64_t*)out, &overlong); break;
    case COVT_OP_VARINT_U32_AS_I64: case COVT_OP_VARINT_ZZ_DELTA_AS_I64: case COVT_OP_VARINT_ZZ_AS_I64: {
        /* Arrays.stream(int[]).mapToLong(i -> i), CovtParser.java:560,566 (ids) and :305-306,310-311 (INT_64 property data) */
        int32_t* tmp = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
        rc = op == COVT_OP_VARINT_U32_AS_I64 ? covt_oracle_decode_varint(blob, limit, pos, n, tmp, &overlong)
             : op == COVT_OP_VARINT_ZZ_AS_I64 ? covt_oracle_decode_zigzag_varint(blob, limit, pos, n, tmp, &overlong)
                                              : covt_oracle_decode_zigzag_delta_varint(blob, limit, pos, n, tmp, &overlong);
        if (rc == COVT_OK) for (uint32_t i = 0; i < n; i++) ((int64_t*)out)[i] = (int64_t)tmp[i];
        free(tmp);
        break;
    }
    case COVT_OP_PFOR_ZZ_DELTA: rc = covt_oracle_decode_fastpfor_zigzag_delta(blob, limit, pos, n, byte_length, (int32_t*)out); break;
    case COVT_OP_PFOR_ZZ_DELTA_XY: rc = covt_oracle_decode_fastpfor_delta_coordinates(blob, limit, pos, n, byte_length, (int32_t*)out); break;
    case COVT_OP_PFOR_DELTA_MORTON: rc = covt_oracle_decode_fastpfor_delta_morton_codes(blob, limit, pos, n, byte_length, num_bits, no_shift, (int32_t*)out); break;
    default: return COVT_ERR_UNSUPPORTED_ENCODING;
    }
    if (rc == COVT_OK && overlong) rc = COVT_ERR_VARINT_OVERLONG;
    return rc;
}

int32_t covt_oracle_decode_stream(const uint8_t* blob, uint64_t blob_len, covt_stream_desc* d, uint32_t flags,
                                  void* out, uint64_t out_cap_bytes)
{
    uint32_t op = d->op ? d->op : (uint32_t)covt_oracle_resolve_op(d->stream_type, d->encoding, d->column_type, flags);
    d->status = COVT_OK;
    d->bytes_consumed = 0;
    d->out_count = 0;
    if (op == COVT_OP_NONE || op >= COVT_NUM_OPS) { d->status = COVT_ERR_UNSUPPORTED_ENCODING; return COVT_OK; }
    if (d->byte_offset > blob_len || d->byte_length > blob_len - d->byte_offset) { d->status = COVT_ERR_TRUNCATED; return COVT_OK; }
    /* library policy shared with the product (see plausible_count): more than 256 values per payload byte cannot decode */
    if ((uint64_t)d->num_values > 256ull * ((uint64_t)d->byte_length + 16ull)) { d->status = COVT_ERR_TRUNCATED; return COVT_OK; }
    uint64_t cnt = covt_oracle_op_out_count(op, d->num_values);
    if (cnt * covt_oracle_op_elem_size(op) > out_cap_bytes) return COVT_ERR_INVALID_ARG;
    uint64_t pos = d->byte_offset;
    int32_t rc = decode_op(blob, d->byte_offset + d->byte_length, &pos, op, d->num_values, d->byte_length, d->num_bits, flags, out);
    d->status = (uint32_t)rc;
    d->bytes_consumed = (uint32_t)(pos - d->byte_offset);
    d->out_count = rc == COVT_OK || rc == COVT_ERR_VARINT_OVERLONG ? cnt : 0;
    return COVT_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Container walkers
 * ------------------------------------------------------------------------------------------- */
typedef struct { const uint8_t* b; uint64_t p, end; int err; } cur_t;

static uint32_t c_varint(cur_t* c)
{
    int32_t v = 0;
    if (varint_java(c->b, c->end, &c->p, &v, NULL)) { c->err = 1; return 0; }
    return (uint32_t)v;
}
static uint32_t c_byte(cur_t* c)
{
    if (c->p >= c->end) { c->err = 1; return 0; }
    return c->b[c->p++];
}
/* decodeString, DecodingUtils.java:21-26: varint length + UTF-8 */
static void c_string(cur_t* c, uint64_t* off, uint32_t* len)
{
    uint32_t n = c_varint(c);
    if (c->err || c->p + n > c->end) { c->err = 1; *off = 0; *len = 0; return; }
    *off = c->p;
    *len = n;
    c->p += n;
}
static int name_is(const uint8_t* b, uint64_t off, uint32_t len, const char* s)
{
    return strlen(s) == len && memcmp(b + off, s, len) == 0;
}

static void layer_init(covt_layer* L, uint32_t tile, uint32_t idx)
{
    memset(L, 0, sizeof(*L));
    L->tile = tile;
    L->layer_index = idx;
    for (int s = 0; s < COVT_NUM_SLOTS; s++) L->streams[s].encoding = COVT_ENC_ABSENT;
}

static const uint8_t slot_stream_type[COVT_NUM_SLOTS] = {
    COVT_ST_DATA, COVT_ST_GEOMETRY_TYPES, COVT_ST_GEOMETRY_OFFSETS, COVT_ST_PART_OFFSETS,
    COVT_ST_RING_OFFSETS, COVT_ST_VERTEX_OFFSETS, COVT_ST_VERTEX_BUFFER, COVT_ST_INDEX_BUFFER};

/* Payload placement. The reference consumes the columns IN METADATA ORDER (CovtParser.java:64-85 iterates a LinkedHashMap),
 * and the streams of the geometry column in the fixed order types, geometry_offsets, part_offsets, ring_offsets,
 * vertex_offsets, vertex_buffer [, index_buffer] (CovtParser.java:405-510). */
static void place_id(covt_layer* L, uint64_t* p)
{
    covt_stream_ref* r = &L->streams[COVT_SLOT_ID];
    if (r->encoding == COVT_ENC_ABSENT) return;
    r->byte_offset = *p;
    *p += r->byte_length;
}
static void place_geometry(covt_layer* L, uint64_t* p)
{
    for (int s = COVT_SLOT_TYPES; s < COVT_NUM_SLOTS; s++) {
        covt_stream_ref* r = &L->streams[s];
        if (r->encoding == COVT_ENC_ABSENT) continue;
        r->byte_offset = *p;
        *p += r->byte_length;
    }
}
/* dispatch of every present stream (CovtParser.decodeGeometryColumn :392-511, decodedIds :552-572) */
static void layer_resolve_ops(covt_layer* L, uint32_t flags)
{
    for (int s = 0; s < COVT_NUM_SLOTS; s++) {
        covt_stream_ref* r = &L->streams[s];
        if (r->encoding == COVT_ENC_ABSENT) continue;
        r->op = (uint8_t)covt_oracle_resolve_op(slot_stream_type[s], r->encoding, L->geom_column_type, flags);
        if (r->op == COVT_OP_NONE) {
            r->status = COVT_ERR_UNSUPPORTED_ENCODING;
            if (!L->status) L->status = COVT_ERR_UNSUPPORTED_ENCODING;
        }
    }
}
/* one entry per column, in metadata order */
enum { COL_ID = 0, COL_GEOMETRY = 1, COL_PROPERTY = 2 };
typedef struct { uint8_t kind; uint8_t data_type; uint64_t listed_bytes; uint64_t meta_pos; } col_t;

/* property-column sink (defined with the property code at the end of this file); NULL = hop over the property payloads */
typedef struct prop_sink prop_sink_t;
static void sink_set_layer(prop_sink_t* k, uint32_t li);
static void sink_begin_column(prop_sink_t* k, uint64_t noff, uint32_t nlen, uint32_t dt, uint32_t ct, uint32_t F);
static void sink_stream(prop_sink_t* k, uint32_t st, uint64_t sub_off, uint32_t sub_len, uint32_t nv, uint32_t bl, uint32_t enc, uint64_t off);
static void sink_end_column(prop_sink_t* k);
/* property columns exist for COMPLETE layers only: a layer whose walk fails takes back what it emitted */
typedef struct { uint64_t n[11]; uint64_t payload_bytes; } sink_mark_t;
static void sink_mark(prop_sink_t* k, sink_mark_t* m);
static void sink_rollback(prop_sink_t* k, const sink_mark_t* m);
static int name_starts(const uint8_t* b, uint64_t off, uint32_t len, const char* s)
{
    size_t n = strlen(s);
    return len >= n && memcmp(b + off, s, n) == 0;
}
/* gen-2 data type byte of the column header -> COVT_DT_*; 0xFF = unknown */
static uint32_t dt_of_gen2(uint32_t g)
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

/* Library policy shared with the product (result buffers are sized from numValues before anything is decoded): a stream that
 * claims more than 256 values per payload byte cannot decode with any codec of the path (the densest, FastPFOR at bit width 0,
 * holds 128 per byte) - the reference would run off the end of its array (ArrayIndexOutOfBounds), so the tile fails. */
static int plausible_count(uint32_t num_values, uint32_t byte_length)
{
    return (uint64_t)num_values <= 256ull * ((uint64_t)byte_length + 16ull);
}

static uint32_t nlz32(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }

/* gen-2b: varint version, varint numLayers | per layer: string name, varint extent, numFeatures, numColumns |
 * per column: string name, byte dataType, byte columnType, varint numStreams | per stream: string name,
 * varint numValues, varint byteLength, byte StreamEncoding. Payload follows each layer's metadata. */
static int32_t parse_gen2b(const uint8_t* blob, uint64_t begin, uint64_t end, uint32_t flags, uint32_t tile,
                           covt_layer* layers, uint32_t cap, uint32_t* n_layers, uint64_t* end_pos, prop_sink_t* sink)
{
    cur_t c = {blob, begin, end, 0};
    *n_layers = 0;
    (void)c_varint(&c); /* version */
    uint32_t num_layers = c_varint(&c);
    if (c.err) return COVT_ERR_TRUNCATED;
    for (uint32_t li = 0; li < num_layers; li++) {
        if (*n_layers >= cap) return COVT_ERR_OOM;
        covt_layer* L = &layers[*n_layers];
        layer_init(L, tile, li);
        sink_set_layer(sink, li);
        const uint64_t layer_start = c.p;
        c_string(&c, &L->name_offset, &L->name_length);
        L->extent = c_varint(&c);
        L->num_features = c_varint(&c);
        L->num_columns = c_varint(&c);
        if (c.err) return COVT_ERR_TRUNCATED;
        L->num_bits = (uint8_t)(32 - nlz32(L->extent));
        if ((uint64_t)L->num_columns > end - c.p) return COVT_ERR_TRUNCATED; /* every column takes bytes: bounds the table below */
        col_t* cols = (col_t*)calloc((size_t)L->num_columns + 1, sizeof(col_t));
        if (!cols) return COVT_ERR_OOM;
        int have_geometry = 0, have_id_column = 0;
        int32_t rc = COVT_OK;
        for (uint32_t ci = 0; ci < L->num_columns && rc == COVT_OK; ci++) {
            uint64_t noff; uint32_t nlen;
            cols[ci].meta_pos = c.p;
            c_string(&c, &noff, &nlen);
            uint32_t data_type = c_byte(&c);
            uint32_t column_type = c_byte(&c);
            uint32_t num_streams = c_varint(&c);
            if (c.err) { rc = COVT_ERR_TRUNCATED; break; }
            (void)data_type;
            int is_id = name_is(blob, noff, nlen, "id");
            int is_geom = name_is(blob, noff, nlen, "geometry");
            if (ci == 0 && !is_id && !is_geom) { rc = COVT_ERR_BAD_METADATA; break; } /* CovtParser.java:67-69 */
            /* the reference keeps the columns in a map keyed by name: one id and one geometry column */
            if ((is_id && have_id_column) || (is_geom && have_geometry)) { rc = COVT_ERR_BAD_METADATA; break; }
            if (is_id) have_id_column = 1;
            if (is_geom) { L->geom_column_type = (uint8_t)column_type; have_geometry = 1; }
            if (is_geom && column_type > COVT_CT_ICE_MORTON_CODE) { rc = COVT_ERR_BAD_METADATA; break; }
            cols[ci].kind = is_id ? COL_ID : (is_geom ? COL_GEOMETRY : COL_PROPERTY);
            for (uint32_t si = 0; si < num_streams; si++) {
                uint64_t soff; uint32_t slen;
                c_string(&c, &soff, &slen);
                uint32_t nv = c_varint(&c);
                uint32_t bl = c_varint(&c);
                uint32_t enc = c_byte(&c);
                if (c.err) { rc = COVT_ERR_TRUNCATED; break; }
                int slot = -1;
                if (is_id) { if (name_is(blob, soff, slen, "data")) slot = COVT_SLOT_ID; }
                else if (is_geom) {
                    if (name_is(blob, soff, slen, "geometry_types")) slot = COVT_SLOT_TYPES;
                    else if (name_is(blob, soff, slen, "geometry_offsets")) slot = COVT_SLOT_GEOM;
                    else if (name_is(blob, soff, slen, "part_offsets")) slot = COVT_SLOT_PART;
                    else if (name_is(blob, soff, slen, "ring_offsets")) slot = COVT_SLOT_RING;
                    else if (name_is(blob, soff, slen, "vertex_offsets")) slot = COVT_SLOT_VOFF;
                    else if (name_is(blob, soff, slen, "vertex_buffer")) slot = COVT_SLOT_VBUF;
                    else if (name_is(blob, soff, slen, "index_buffer")) slot = COVT_SLOT_INDEX;
                }
                if (slot >= 0) {
                    if (enc > COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) { rc = COVT_ERR_BAD_METADATA; break; }
                    if (!plausible_count(nv, bl)) { rc = COVT_ERR_TRUNCATED; break; }
                    L->streams[slot].num_values = nv;
                    L->streams[slot].byte_length = bl;
                    L->streams[slot].encoding = (uint8_t)enc;
                    if (slot == COVT_SLOT_ID) L->has_id = 1;
                } else if (is_id || is_geom) {
                    rc = COVT_ERR_BAD_METADATA;
                    break;
                } else {
                    cols[ci].listed_bytes += bl;
                }
            }
        }
        if (rc == COVT_OK && (!have_geometry || L->streams[COVT_SLOT_TYPES].encoding == COVT_ENC_ABSENT ||
                              L->streams[COVT_SLOT_VBUF].encoding == COVT_ENC_ABSENT))
            rc = COVT_ERR_BAD_METADATA;
        /* payloads in column-metadata order */
        uint64_t p = c.p;
        sink_mark_t mark;
        sink_mark(sink, &mark);
        for (uint32_t ci = 0; ci < L->num_columns && rc == COVT_OK; ci++) {
            if (cols[ci].kind == COL_ID) place_id(L, &p);
            else if (cols[ci].kind == COL_GEOMETRY) place_geometry(L, &p);
            else if (!sink) p += cols[ci].listed_bytes;
            else {
                /* the column's metadata once more (it parsed fine a moment ago): its streams in payload order. Stream names of
                 * property columns: present, data, length, dictionary; localized dictionaries list pairs (present_<s>, <s>) and
                 * share one length + dictionary (oracle/properties.py) */
                cur_t m = {blob, cols[ci].meta_pos, end, 0};
                uint64_t noff; uint32_t nlen;
                c_string(&m, &noff, &nlen);
                uint32_t dt2 = c_byte(&m);
                uint32_t column_type = c_byte(&m);
                uint32_t num_streams = c_varint(&m);
                sink_begin_column(sink, noff, nlen, dt_of_gen2(dt2), column_type, L->num_features);
                for (uint32_t si = 0; si < num_streams; si++) {
                    uint64_t soff; uint32_t slen;
                    c_string(&m, &soff, &slen);
                    uint32_t nv = c_varint(&m);
                    uint32_t bl = c_varint(&m);
                    uint32_t enc = c_byte(&m);
                    uint32_t st = COVT_ST_DATA;
                    uint64_t sub_off = 0; uint32_t sub_len = 0;
                    if (name_is(blob, soff, slen, "present")) st = COVT_ST_PRESENT;
                    else if (name_is(blob, soff, slen, "data")) st = COVT_ST_DATA;
                    else if (name_is(blob, soff, slen, "length")) st = COVT_ST_LENGTH;
                    else if (name_is(blob, soff, slen, "dictionary")) st = COVT_ST_DICTIONARY;
                    else if (name_starts(blob, soff, slen, "present_")) { st = COVT_ST_PRESENT; sub_off = soff + 8; sub_len = slen - 8; }
                    else { st = COVT_ST_DATA; sub_off = soff; sub_len = slen; }
                    sink_stream(sink, st, sub_off, sub_len, nv, bl, enc, p);
                    p += bl;
                }
                sink_end_column(sink);
            }
            if (p > end) rc = COVT_ERR_TRUNCATED;
        }
        free(cols);
        if (rc != COVT_OK) { sink_rollback(sink, &mark); return rc; }
        L->header_offset = layer_start;
        layer_resolve_ops(L, flags);
        c.p = p;
        (*n_layers)++;
    }
    *end_pos = c.p;
    return COVT_OK;
}

/* length in bytes of a Byte-RLE stream that decodes to n bytes (present streams of gen-3 property
 * columns are not listed in the metadata, CovtConverter.java:434-436; DecodingUtils.java:290-306) */
static int byte_rle_span(const uint8_t* b, uint64_t p, uint64_t end, uint32_t n, uint64_t* out_end)
{
    uint32_t done = 0;
    while (done < n) {
        if (p >= end) return -1;
        uint8_t c = b[p++];
        if (c < 0x80) { done += (uint32_t)c + 3; p += 1; }
        else { done += 256u - c; p += 256u - c; }
        if (p > end) return -1;
    }
    *out_end = p;
    return 0;
}

/* gen-3: CovtParser.decodeLayerMetadata, CovtParser.java:574-652 (no tile header; loop until EOF :56). */
static int32_t parse_gen3(const uint8_t* blob, uint64_t begin, uint64_t end, const covt_tilejson* tj, uint32_t flags,
                          uint32_t tile, covt_layer* layers, uint32_t cap, uint32_t* n_layers, uint64_t* end_pos, prop_sink_t* sink)
{
    cur_t c = {blob, begin, end, 0};
    *n_layers = 0;
    uint32_t li = 0;
    while (c.p < end) {
        if (*n_layers >= cap) return COVT_ERR_OOM;
        covt_layer* L = &layers[*n_layers];
        layer_init(L, tile, li);
        sink_set_layer(sink, li);
        const uint64_t layer_start = c.p;
        uint32_t header = c_byte(&c);
        int optimized = header & 1; /* :575-578 */
        uint32_t n_fields = 0;
        if (optimized) {
            uint32_t layer_id = c_varint(&c); /* :584-589 */
            if (c.err) return COVT_ERR_TRUNCATED;
            if (!tj || layer_id >= tj->n_vector_layers) return COVT_ERR_BAD_METADATA; /* Java: NPE / IndexOutOfBounds */
            n_fields = tj->n_fields[layer_id];
            L->name_offset = layer_id;
            L->name_length = 0;
        } else {
            c_string(&c, &L->name_offset, &L->name_length); /* :592 */
        }
        L->extent = c_varint(&c); /* :595-598 */
        L->num_features = c_varint(&c);
        L->num_columns = c_varint(&c);
        if (c.err) return COVT_ERR_TRUNCATED;
        L->num_bits = (uint8_t)(32 - nlz32(L->extent)); /* CovtParser.java:77 */
        if ((uint64_t)L->num_columns > end - c.p) return COVT_ERR_TRUNCATED; /* every column takes bytes: bounds the table below */
        col_t* cols = (col_t*)calloc((size_t)L->num_columns + 1, sizeof(col_t));
        if (!cols) return COVT_ERR_OOM;
        int have_geometry = 0, have_id_column = 0;
        int32_t rc = COVT_OK;
        for (uint32_t ci = 0; ci < L->num_columns && rc == COVT_OK; ci++) {
            int is_id = 0, is_geom = 0;
            cols[ci].meta_pos = c.p;
            if (optimized || ci == 0) { /* :604-614 */
                uint32_t column_id = c_varint(&c);
                if (column_id > 1) {
                    if (!optimized || column_id - 2 >= n_fields) { rc = COVT_ERR_BAD_METADATA; break; } /* fields == null -> NPE */
                } else if (column_id == 0) is_id = 1;
                else is_geom = 1;
            } else {
                uint64_t noff; uint32_t nlen;
                c_string(&c, &noff, &nlen); /* :616 */
                if (!c.err) { is_id = name_is(blob, noff, nlen, "id"); is_geom = name_is(blob, noff, nlen, "geometry"); }
            }
            uint32_t column_desc = c_byte(&c); /* :619-624 */
            if (c.err) { rc = COVT_ERR_TRUNCATED; break; }
            uint32_t data_type = (column_desc >> 3) & 0xF;
            uint32_t column_type = column_desc & 0x7;
            if (column_type > COVT_CT_ICE_MORTON_CODE) { rc = COVT_ERR_BAD_METADATA; break; } /* ColumnType.values()[..] throws */
            if (ci == 0 && !is_id && !is_geom) { rc = COVT_ERR_BAD_METADATA; break; } /* :67-69 */
            /* the reference keeps the columns in a map keyed by name: one id and one geometry column */
            if ((is_id && have_id_column) || (is_geom && have_geometry)) { rc = COVT_ERR_BAD_METADATA; break; }
            if (is_id) have_id_column = 1;
            if (is_geom) { L->geom_column_type = (uint8_t)column_type; have_geometry = 1; }
            cols[ci].kind = is_id ? COL_ID : (is_geom ? COL_GEOMETRY : COL_PROPERTY);
            cols[ci].data_type = (uint8_t)data_type;
            for (;;) { /* :628-648 */
                uint32_t stream_desc = c_byte(&c);
                uint32_t stream_type = stream_desc >> 4;
                uint32_t enc = stream_desc & 0xF;
                uint32_t nv = c_varint(&c);
                uint32_t bl = c_varint(&c);
                if (c.err) { rc = COVT_ERR_TRUNCATED; break; }
                if (stream_type > COVT_ST_INDEX_BUFFER || enc > COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) { rc = COVT_ERR_BAD_METADATA; break; }
                int slot = -1;
                if (is_id && stream_type == COVT_ST_DATA) slot = COVT_SLOT_ID;
                else if (is_geom && stream_type >= COVT_ST_GEOMETRY_TYPES && stream_type <= COVT_ST_VERTEX_BUFFER)
                    slot = COVT_SLOT_TYPES + (int)(stream_type - COVT_ST_GEOMETRY_TYPES);
                else if (is_geom && stream_type == COVT_ST_INDEX_BUFFER) slot = COVT_SLOT_INDEX;
                if (slot >= 0) {
                    if (!plausible_count(nv, bl)) { rc = COVT_ERR_TRUNCATED; break; }
                    L->streams[slot].num_values = nv;
                    L->streams[slot].byte_length = bl;
                    L->streams[slot].encoding = (uint8_t)enc;
                    if (slot == COVT_SLOT_ID) L->has_id = 1;
                } else if (is_id || is_geom) { rc = COVT_ERR_BAD_METADATA; break; }
                else cols[ci].listed_bytes += bl;
                /* last stream of the column, :639-647. (INDEX_BUFFER, when present, precedes VERTEX_BUFFER in the
                 * metadata so that the reference terminator still ends the column.) */
                if (data_type == COVT_DT_GEOMETRY && stream_type == COVT_ST_VERTEX_BUFFER) break;
                else if (stream_type == COVT_ST_DATA && column_type == COVT_CT_PLAIN) break;
                else if (stream_type == COVT_ST_DICTIONARY) break;
            }
        }
        if (rc == COVT_OK && (!have_geometry || L->streams[COVT_SLOT_TYPES].encoding == COVT_ENC_ABSENT ||
                              L->streams[COVT_SLOT_VBUF].encoding == COVT_ENC_ABSENT))
            rc = COVT_ERR_BAD_METADATA;
        /* payloads in column-metadata order. Property columns: BOOLEAN = listed data stream only (CovtParser.java:280-290);
         * every other type = unlisted Byte-RLE present stream of ceil(numFeatures/8) bytes (:295) + its listed streams */
        uint64_t p = c.p;
        sink_mark_t mark;
        sink_mark(sink, &mark);
        for (uint32_t ci = 0; ci < L->num_columns && rc == COVT_OK; ci++) {
            if (cols[ci].kind == COL_ID) place_id(L, &p);
            else if (cols[ci].kind == COL_GEOMETRY) place_geometry(L, &p);
            else if (!sink) {
                if (cols[ci].data_type != COVT_DT_BOOLEAN) {
                    uint32_t nbytes = (L->num_features + 7) / 8;
                    if (byte_rle_span(blob, p, end, nbytes, &p)) { rc = COVT_ERR_TRUNCATED; break; }
                }
                p += cols[ci].listed_bytes;
            } else {
                /* the column's metadata once more (it parsed fine a moment ago): its streams in payload order */
                cur_t m = {blob, cols[ci].meta_pos, end, 0};
                uint64_t noff = 0; uint32_t nlen = 0;
                if (optimized || ci == 0) noff = c_varint(&m) - 2; /* index into the TileJSON fields of the layer */
                else c_string(&m, &noff, &nlen);
                uint32_t column_desc = c_byte(&m);
                uint32_t data_type = (column_desc >> 3) & 0xF, column_type = column_desc & 0x7;
                sink_begin_column(sink, noff, nlen, data_type, column_type, L->num_features);
                if (data_type != COVT_DT_BOOLEAN) {
                    uint64_t q = p;
                    if (byte_rle_span(blob, p, end, (L->num_features + 7) / 8, &q)) { rc = COVT_ERR_TRUNCATED; break; }
                    sink_stream(sink, COVT_ST_PRESENT, 0, 0, L->num_features, (uint32_t)(q - p), COVT_ENC_BOOLEAN_RLE, p);
                    p = q;
                }
                for (;;) {
                    uint32_t stream_desc = c_byte(&m);
                    uint32_t stream_type = stream_desc >> 4;
                    uint32_t nv = c_varint(&m);
                    uint32_t bl = c_varint(&m);
                    /* BOOLEAN data = one bit per FEATURE whatever numValues says (CovtParser.java:280-283) */
                    sink_stream(sink, stream_type, 0, 0, (data_type == COVT_DT_BOOLEAN && stream_type == COVT_ST_DATA) ? L->num_features : nv, bl,
                                stream_desc & 0xF, p);
                    p += bl;
                    if (m.err) break;
                    if (data_type == COVT_DT_GEOMETRY && stream_type == COVT_ST_VERTEX_BUFFER) break;
                    else if (stream_type == COVT_ST_DATA && column_type == COVT_CT_PLAIN) break;
                    else if (stream_type == COVT_ST_DICTIONARY) break;
                }
                sink_end_column(sink);
            }
            if (p > end) rc = COVT_ERR_TRUNCATED;
        }
        free(cols);
        if (rc != COVT_OK) { sink_rollback(sink, &mark); return rc; }
        L->header_offset = layer_start;
        layer_resolve_ops(L, flags);
        c.p = p;
        (*n_layers)++;
        li++;
    }
    *end_pos = c.p;
    return COVT_OK;
}

int32_t covt_oracle_parse_tile(const uint8_t* blob, uint64_t begin, uint64_t end, uint32_t container,
                               const covt_tilejson* tj, uint32_t flags, uint32_t tile_index,
                               covt_layer* layers, uint32_t cap, uint32_t* n_layers, uint64_t* end_pos)
{
    uint64_t ep = begin;
    int32_t rc;
    if (container == COVT_CONTAINER_GEN2B) rc = parse_gen2b(blob, begin, end, flags, tile_index, layers, cap, n_layers, &ep, NULL);
    else if (container == COVT_CONTAINER_GEN3) rc = parse_gen3(blob, begin, end, tj, flags, tile_index, layers, cap, n_layers, &ep, NULL);
    else return COVT_ERR_INVALID_ARG;
    if (end_pos) *end_pos = ep;
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * Result layout (DESIGN.md "result layout"): shared rule with libcovt_b200
 * ------------------------------------------------------------------------------------------- */
static const uint8_t buf_elem_size[COVT_NUM_BUFFERS] = {1, 8, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1};
uint32_t covt_oracle_buffer_elem_size(uint32_t which) { return which < COVT_NUM_BUFFERS ? buf_elem_size[which] : 0; }

static inline uint64_t align_elems(uint64_t n, uint32_t elem_size)
{
    uint64_t per = 16 / elem_size;
    return (n + per - 1) / per * per;
}

/* slice sizes (elements, before 16-byte rounding) of one layer in every result buffer */
static void layer_slice_sizes(covt_layer* L, uint32_t flags, uint64_t sz[COVT_NUM_BUFFERS])
{
    memset(sz, 0, COVT_NUM_BUFFERS * sizeof(uint64_t));
    if (L->status == COVT_ERR_BAD_METADATA) return;
#define NV(slot) (L->streams[slot].encoding == COVT_ENC_ABSENT ? 0ull : (uint64_t)L->streams[slot].num_values)
    uint64_t F = NV(COVT_SLOT_TYPES);
    sz[COVT_BUF_S_GEOMETRY_TYPES] = F;
    sz[COVT_BUF_S_IDS] = NV(COVT_SLOT_ID);
    sz[COVT_BUF_S_GEOMETRY_OFFSETS] = NV(COVT_SLOT_GEOM);
    sz[COVT_BUF_S_PART_OFFSETS] = NV(COVT_SLOT_PART);
    sz[COVT_BUF_S_RING_OFFSETS] = NV(COVT_SLOT_RING);
    sz[COVT_BUF_S_VERTEX_OFFSETS] = NV(COVT_SLOT_VOFF);
    uint64_t vb_ints = NV(COVT_SLOT_VBUF);
    if (L->geom_column_type == COVT_CT_ICE_MORTON_CODE) vb_ints *= 2;
    else if (L->geom_column_type == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) vb_ints *= 2;
    sz[COVT_BUF_S_VERTEX_BUFFER] = vb_ints;
    sz[COVT_BUF_S_INDEX_BUFFER] = NV(COVT_SLOT_INDEX);
    if (!(flags & COVT_FLAG_SKIP_ASSEMBLY)) {
        uint64_t V = L->streams[COVT_SLOT_VOFF].encoding != COVT_ENC_ABSENT ? NV(COVT_SLOT_VOFF) : vb_ints / 2;
        uint64_t cap_parts = F + NV(COVT_SLOT_PART);
        uint64_t cap_rings = cap_parts + NV(COVT_SLOT_RING);
        L->cap_parts = (uint32_t)cap_parts;
        L->cap_rings = (uint32_t)cap_rings;
        sz[COVT_BUF_A_GEOM_OFFSETS] = F + 1;
        sz[COVT_BUF_A_PART_OFFSETS] = cap_parts + 1;
        sz[COVT_BUF_A_RING_OFFSETS] = cap_rings + 1;
        sz[COVT_BUF_A_COORDS] = 2 * (V + ((flags & COVT_FLAG_CLOSE_RINGS) ? NV(COVT_SLOT_RING) : 0));
    }
#undef NV
}

/* ---------------------------------------------------------------------------------------------
 * Assembler: intended semantics of CovtParser.convertGeometryColumn (CovtParser.java:135-274,
 * 513-550) = what the encoder wrote (CovtConverter.java:580-639,689-758) = JS LayerTable
 * (JS/src/decoder/layerTable.ts:100-209); SURVEY §A.7. Normal form: every feature -> parts ->
 * rings -> vertices, degenerate levels have length 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* types; uint32_t F;
    const int32_t *geom, *part, *ring, *voff, *vbuf;
    uint32_t n_geom, n_part, n_ring, n_voff; uint64_t vbuf_ints;
    int32_t *a_geom, *a_part, *a_ring, *a_coords;
    uint32_t cap_parts, cap_rings; uint64_t cap_coords_ints;
} asm_in_t;

static int32_t assemble_layer(const asm_in_t* a, uint32_t flags, uint32_t* n_parts, uint32_t* n_rings,
                              uint32_t* n_vertices, uint32_t* n_coords)
{
    const int close = (flags & COVT_FLAG_CLOSE_RINGS) != 0;
    const int ice = a->voff != NULL;
    const uint64_t src_total = ice ? a->n_voff : a->vbuf_ints / 2;
    const uint64_t dict = a->vbuf_ints / 2;
    uint32_t gc = 0, pc = 0, rc = 0, p = 0, r = 0;
    uint64_t s = 0, v = 0;
    int32_t status = COVT_OK;
    a->a_geom[0] = 0;
    a->a_part[0] = 0;
    a->a_ring[0] = 0;
#define FAIL(code) do { status = (code); goto out; } while (0)
#define EMIT_RING(n_, closed_)                                                                  \
    do {                                                                                        \
        int64_t n__ = (n_);                                                                     \
        if (n__ < 0 || s + (uint64_t)n__ > src_total) FAIL(COVT_ERR_TOPOLOGY);                  \
        uint64_t extra__ = ((closed_) && n__ > 0) ? 1 : 0;                                      \
        if (r >= a->cap_rings || 2 * (v + (uint64_t)n__ + extra__) > a->cap_coords_ints) FAIL(COVT_ERR_TOPOLOGY); \
        for (int64_t i__ = 0; i__ < n__ + (int64_t)extra__; i__++) {                            \
            uint64_t si__ = s + (uint64_t)(i__ == n__ ? 0 : i__);                               \
            uint64_t vi__ = si__;                                                               \
            if (ice) {                                                                          \
                int32_t o__ = a->voff[si__];                                                    \
                if (o__ < 0 || (uint64_t)o__ >= dict) FAIL(COVT_ERR_TOPOLOGY);                  \
                vi__ = (uint64_t)o__;                                                           \
            }                                                                                   \
            a->a_coords[2 * v] = a->vbuf[2 * vi__];                                             \
            a->a_coords[2 * v + 1] = a->vbuf[2 * vi__ + 1];                                     \
            v++;                                                                                \
        }                                                                                       \
        s += (uint64_t)n__;                                                                     \
        a->a_ring[++r] = (int32_t)v;                                                            \
    } while (0)
#define END_PART() do { if (p >= a->cap_parts) FAIL(COVT_ERR_TOPOLOGY); a->a_part[++p] = (int32_t)r; } while (0)
#define NEXT(arr, cur, n, dst) do { if ((cur) >= (n)) FAIL(COVT_ERR_TOPOLOGY); (dst) = (arr)[(cur)++]; } while (0)

    for (uint32_t f = 0; f < a->F; f++) {
        uint8_t t = a->types[f];
        int32_t cnt, nr, nparts;
        switch (t) {
        case COVT_GT_POINT: /* CovtParser.java:153-167 */
            EMIT_RING(1, 0);
            END_PART();
            break;
        case COVT_GT_LINESTRING: /* :168-181 */
            NEXT(a->part, pc, a->n_part, cnt);
            EMIT_RING(cnt, 0);
            END_PART();
            break;
        case COVT_GT_POLYGON: /* :182-206 */
            NEXT(a->part, pc, a->n_part, nr);
            if (nr < 0) FAIL(COVT_ERR_TOPOLOGY);
            for (int32_t k = 0; k < nr; k++) {
                NEXT(a->ring, rc, a->n_ring, cnt);
                EMIT_RING(cnt, close);
            }
            END_PART();
            break;
        case COVT_GT_MULTILINESTRING: /* :207-228 */
            NEXT(a->geom, gc, a->n_geom, nparts);
            if (nparts < 0) FAIL(COVT_ERR_TOPOLOGY);
            for (int32_t k = 0; k < nparts; k++) {
                NEXT(a->part, pc, a->n_part, cnt);
                EMIT_RING(cnt, 0);
                END_PART();
            }
            break;
        case COVT_GT_MULTIPOLYGON: /* :229-267 (intended semantics; the Java branch is buggy) */
            NEXT(a->geom, gc, a->n_geom, nparts);
            if (nparts < 0) FAIL(COVT_ERR_TOPOLOGY);
            for (int32_t k = 0; k < nparts; k++) {
                NEXT(a->part, pc, a->n_part, nr);
                if (nr < 0) FAIL(COVT_ERR_TOPOLOGY);
                for (int32_t j = 0; j < nr; j++) {
                    NEXT(a->ring, rc, a->n_ring, cnt);
                    EMIT_RING(cnt, close);
                }
                END_PART();
            }
            break;
        default: /* :268-270 */
            FAIL(COVT_ERR_UNSUPPORTED_GEOMETRY);
        }
        a->a_geom[f + 1] = (int32_t)p;
    }
out:
#undef FAIL
#undef EMIT_RING
#undef END_PART
#undef NEXT
    *n_parts = p;
    *n_rings = r;
    *n_vertices = (uint32_t)s;
    *n_coords = (uint32_t)v;
    return status;
}

/* decode every stream of a layer into the result buffers and assemble */
static void decode_layer(const uint8_t* blob, uint64_t tile_end, covt_layer* L, uint32_t flags, void* const bufs[COVT_NUM_BUFFERS])
{
    static const uint8_t slot_buf[COVT_NUM_SLOTS] = {
        COVT_BUF_S_IDS, COVT_BUF_S_GEOMETRY_TYPES, COVT_BUF_S_GEOMETRY_OFFSETS, COVT_BUF_S_PART_OFFSETS,
        COVT_BUF_S_RING_OFFSETS, COVT_BUF_S_VERTEX_OFFSETS, COVT_BUF_S_VERTEX_BUFFER, COVT_BUF_S_INDEX_BUFFER};
    if (L->status == COVT_ERR_BAD_METADATA) return;
    void* dst[COVT_NUM_SLOTS];
    for (int s = 0; s < COVT_NUM_SLOTS; s++) {
        covt_stream_ref* r = &L->streams[s];
        dst[s] = NULL;
        if (r->encoding == COVT_ENC_ABSENT) continue;
        uint32_t b = slot_buf[s];
        dst[s] = (uint8_t*)bufs[b] + L->out[b] * buf_elem_size[b];
        if (r->op == COVT_OP_NONE) continue;
        uint32_t n = r->num_values;
        if (s == COVT_SLOT_VBUF && L->geom_column_type == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) n *= 2;
        uint64_t pos = r->byte_offset;
        uint64_t limit = r->byte_offset + r->byte_length;
        if (limit > tile_end) limit = tile_end;
        int32_t rc = decode_op(blob, limit, &pos, r->op, n, r->byte_length, L->num_bits, flags, dst[s]);
        r->status = (uint32_t)rc;
        if (rc != COVT_OK && !L->status) L->status = (uint32_t)rc;
    }
    if (flags & COVT_FLAG_SKIP_ASSEMBLY) return;
    if (L->status != COVT_OK) return;
    asm_in_t a;
    memset(&a, 0, sizeof(a));
    a.types = (const uint8_t*)dst[COVT_SLOT_TYPES];
    a.F = L->streams[COVT_SLOT_TYPES].num_values;
    a.geom = (const int32_t*)dst[COVT_SLOT_GEOM];
    a.n_geom = a.geom ? L->streams[COVT_SLOT_GEOM].num_values : 0;
    a.part = (const int32_t*)dst[COVT_SLOT_PART];
    a.n_part = a.part ? L->streams[COVT_SLOT_PART].num_values : 0;
    a.ring = (const int32_t*)dst[COVT_SLOT_RING];
    a.n_ring = a.ring ? L->streams[COVT_SLOT_RING].num_values : 0;
    a.voff = (const int32_t*)dst[COVT_SLOT_VOFF];
    a.n_voff = a.voff ? L->streams[COVT_SLOT_VOFF].num_values : 0;
    a.vbuf = (const int32_t*)dst[COVT_SLOT_VBUF];
    uint64_t vb_ints = L->streams[COVT_SLOT_VBUF].num_values;
    if (L->geom_column_type == COVT_CT_ICE_MORTON_CODE) vb_ints *= 2;
    else if (L->geom_column_type == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) vb_ints *= 2;
    a.vbuf_ints = vb_ints;
    a.a_geom = (int32_t*)bufs[COVT_BUF_A_GEOM_OFFSETS] + L->out[COVT_BUF_A_GEOM_OFFSETS];
    a.a_part = (int32_t*)bufs[COVT_BUF_A_PART_OFFSETS] + L->out[COVT_BUF_A_PART_OFFSETS];
    a.a_ring = (int32_t*)bufs[COVT_BUF_A_RING_OFFSETS] + L->out[COVT_BUF_A_RING_OFFSETS];
    a.a_coords = (int32_t*)bufs[COVT_BUF_A_COORDS] + L->out[COVT_BUF_A_COORDS];
    a.cap_parts = L->cap_parts;
    a.cap_rings = L->cap_rings;
    uint64_t V = a.voff ? a.n_voff : vb_ints / 2;
    a.cap_coords_ints = 2 * (V + ((flags & COVT_FLAG_CLOSE_RINGS) ? a.n_ring : 0));
    int32_t rc = assemble_layer(&a, flags, &L->n_parts, &L->n_rings, &L->n_vertices, &L->n_coords);
    if (rc != COVT_OK) L->status = (uint32_t)rc;
}

/* ---------------------------------------------------------------------------------------------
 * Batch driver
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* blob; const uint64_t* tile_offsets; uint32_t container; const covt_tilejson* tj; uint32_t flags;
    covt_oracle_result* R; covt_layer** per_tile; uint32_t* per_tile_n;
} batch_arg_t;

static void parse_body(void* p, int64_t b, int64_t e, uint32_t thread)
{
    batch_arg_t* A = (batch_arg_t*)p;
    (void)thread;
    for (int64_t t = b; t < e; t++) {
        uint32_t cap = 64;
        for (;;) {
            covt_layer* ls = (covt_layer*)malloc((size_t)cap * sizeof(covt_layer));
            uint32_t n = 0;
            uint64_t ep = 0;
            int32_t rc = covt_oracle_parse_tile(A->blob, A->tile_offsets[t], A->tile_offsets[t + 1], A->container, A->tj,
                                                A->flags, (uint32_t)t, ls, cap, &n, &ep);
            if (rc == COVT_ERR_OOM && cap < (1u << 20)) { free(ls); cap *= 4; continue; }
            if (rc == COVT_OK && ep != A->tile_offsets[t + 1]) rc = COVT_ERR_TRUNCATED;
            A->R->tile_status[t] = (uint32_t)rc;
            A->per_tile[t] = ls;
            A->per_tile_n[t] = n;
            break;
        }
    }
}

static void decode_body(void* p, int64_t b, int64_t e, uint32_t thread)
{
    batch_arg_t* A = (batch_arg_t*)p;
    (void)thread;
    for (int64_t l = b; l < e; l++) {
        covt_layer* L = &A->R->layers[l];
        decode_layer(A->blob, A->tile_offsets[L->tile + 1], L, A->flags, A->R->buffers);
    }
}

int32_t covt_oracle_decode_batch(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                 const covt_tilejson* tj, uint32_t flags, uint32_t n_threads, covt_oracle_result** out)
{
    covt_oracle_result* R = (covt_oracle_result*)calloc(1, sizeof(*R));
    if (!R) return COVT_ERR_OOM;
    R->n_tiles = n_tiles;
    R->tile_status = (uint32_t*)calloc((size_t)n_tiles + 1, sizeof(uint32_t));
    R->first_layer = (uint32_t*)calloc((size_t)n_tiles + 2, sizeof(uint32_t));
    batch_arg_t A = {blob, tile_offsets, container, tj, flags, R, NULL, NULL};
    /* pass 1: parse every tile (per-tile layer lists), then concatenate in tile order */
    A.per_tile = (covt_layer**)calloc((size_t)n_tiles + 1, sizeof(covt_layer*));
    A.per_tile_n = (uint32_t*)calloc((size_t)n_tiles + 1, sizeof(uint32_t));
    parallel_for(n_tiles, 64, n_threads, parse_body, &A);
    uint64_t total_layers = 0;
    for (uint32_t t = 0; t < n_tiles; t++) { R->first_layer[t] = (uint32_t)total_layers; total_layers += A.per_tile_n[t]; }
    R->first_layer[n_tiles] = (uint32_t)total_layers;
    R->n_layers = (uint32_t)total_layers;
    R->layers = (covt_layer*)malloc((size_t)(total_layers + 1) * sizeof(covt_layer));
    for (uint32_t t = 0; t < n_tiles; t++) {
        memcpy(R->layers + R->first_layer[t], A.per_tile[t], (size_t)A.per_tile_n[t] * sizeof(covt_layer));
        free(A.per_tile[t]);
    }
    free(A.per_tile);
    free(A.per_tile_n);
    /* pass 2: result layout = exclusive prefix sums of 16-byte-rounded slice sizes in layer order */
    uint64_t run[COVT_NUM_BUFFERS];
    memset(run, 0, sizeof(run));
    for (uint32_t l = 0; l < R->n_layers; l++) {
        covt_layer* L = &R->layers[l];
        uint64_t sz[COVT_NUM_BUFFERS];
        layer_slice_sizes(L, flags, sz);
        for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
            L->out[b] = run[b];
            run[b] += align_elems(sz[b], buf_elem_size[b]);
        }
        for (int s = 0; s < COVT_NUM_SLOTS; s++)
            if (L->streams[s].encoding != COVT_ENC_ABSENT) R->payload_bytes += L->streams[s].byte_length;
    }
    for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
        R->counts[b] = run[b];
        R->buffers[b] = calloc((size_t)run[b] + 16, buf_elem_size[b]);
        if (!R->buffers[b]) { covt_oracle_result_free(R); return COVT_ERR_OOM; }
    }
    /* pass 3: decode + assemble */
    parallel_for(R->n_layers, 16, n_threads, decode_body, &A);
    for (uint32_t l = 0; l < R->n_layers; l++) {
        covt_layer* L = &R->layers[l];
        R->vertices += L->n_vertices;
        if (L->status && !R->tile_status[L->tile]) R->tile_status[L->tile] = L->status;
    }
    *out = R;
    return COVT_OK;
}

/* Timed variant for the CPU baseline: same work per tile (parse, decode every stream, assemble), outputs
 * written to per-thread scratch that is reused from tile to tile, as a Java caller would let the GC do. */
typedef struct {
    covt_layer* ls; uint32_t cap_layers; void* bufs[COVT_NUM_BUFFERS]; uint64_t caps[COVT_NUM_BUFFERS];
    uint64_t pb, vx, cs; int32_t status; char pad[64];
} timed_thread_t;
typedef struct {
    const uint8_t* blob; const uint64_t* tile_offsets; uint32_t container; const covt_tilejson* tj; uint32_t flags;
    timed_thread_t* th;
} timed_arg_t;

static void timed_body(void* p, int64_t b, int64_t e, uint32_t thread)
{
    timed_arg_t* A = (timed_arg_t*)p;
    timed_thread_t* T = &A->th[thread];
    if (!T->ls) { T->cap_layers = 256; T->ls = (covt_layer*)malloc((size_t)T->cap_layers * sizeof(covt_layer)); }
    for (int64_t t = b; t < e; t++) {
        uint32_t n = 0;
        uint64_t ep = 0;
        int32_t rc;
        for (;;) {
            rc = covt_oracle_parse_tile(A->blob, A->tile_offsets[t], A->tile_offsets[t + 1], A->container, A->tj, A->flags,
                                        (uint32_t)t, T->ls, T->cap_layers, &n, &ep);
            if (rc == COVT_ERR_OOM && T->cap_layers < (1u << 20)) {
                T->cap_layers *= 4;
                T->ls = (covt_layer*)realloc(T->ls, (size_t)T->cap_layers * sizeof(covt_layer));
                continue;
            }
            break;
        }
        if (rc != COVT_OK) { T->status = rc; continue; }
        for (uint32_t l = 0; l < n; l++) {
            covt_layer* L = &T->ls[l];
            uint64_t sz[COVT_NUM_BUFFERS];
            layer_slice_sizes(L, A->flags, sz);
            for (int k = 0; k < COVT_NUM_BUFFERS; k++) {
                L->out[k] = 0;
                if (sz[k] + 16 > T->caps[k]) {
                    T->caps[k] = (sz[k] + 16) * 2;
                    free(T->bufs[k]);
                    T->bufs[k] = malloc((size_t)T->caps[k] * buf_elem_size[k]);
                }
            }
            decode_layer(A->blob, A->tile_offsets[t + 1], L, A->flags, T->bufs);
            for (int s = 0; s < COVT_NUM_SLOTS; s++)
                if (L->streams[s].encoding != COVT_ENC_ABSENT) T->pb += L->streams[s].byte_length;
            T->vx += L->n_vertices;
            if (L->n_coords) T->cs += (uint64_t)(uint32_t)((int32_t*)T->bufs[COVT_BUF_A_COORDS])[2 * (L->n_coords - 1)] + L->n_rings;
            if (L->status) T->status = (int32_t)L->status;
        }
    }
}

int32_t covt_oracle_decode_batch_timed(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                       const covt_tilejson* tj, uint32_t flags, uint32_t n_threads,
                                       uint64_t* payload_bytes, uint64_t* vertices, uint64_t* checksum)
{
    if (n_threads == 0) n_threads = pf_default_threads();
    if (n_threads > 256) n_threads = 256;
    timed_thread_t* th = (timed_thread_t*)calloc(n_threads, sizeof(timed_thread_t));
    timed_arg_t A = {blob, tile_offsets, container, tj, flags, th};
    parallel_for(n_tiles, 16, n_threads, timed_body, &A);
    uint64_t pb = 0, vx = 0, cs = 0;
    int32_t status = COVT_OK;
    for (uint32_t i = 0; i < n_threads; i++) {
        pb += th[i].pb; vx += th[i].vx; cs += th[i].cs;
        if (th[i].status) status = th[i].status;
        for (int k = 0; k < COVT_NUM_BUFFERS; k++) free(th[i].bufs[k]);
        free(th[i].ls);
    }
    free(th);
    if (payload_bytes) *payload_bytes = pb;
    if (vertices) *vertices = vx;
    if (checksum) *checksum = cs;
    return status;
}

void covt_oracle_result_free(covt_oracle_result* r)
{
    if (!r) return;
    for (int b = 0; b < COVT_NUM_BUFFERS; b++) free(r->buffers[b]);
    free(r->layers);
    free(r->tile_status);
    free(r->first_layer);
    free(r);
}

/* ---------------------------------------------------------------------------------------------
 * Property columns                               J/decoder/CovtParser.java:276-390
 * Layout facts of the gen-2b container established on the 129 fixtures and pinned on the partner MVT tiles by
 * tests/test_oracle_properties.py (see oracle/properties.py, the first statement of this code): property payloads follow the
 * geometry payload of their layer in column-metadata order. gen-3 (HEAD): PRESENT streams are not listed
 * (CovtConverter.java:434-436), BOOLEAN data holds one bit per feature (CovtParser.java:280-290).
 *
 * The container walkers above call the sink once per property column, in payload order. Column status = the first failure in
 * this order: (1) metadata-only checks (missing stream, unsupported type / encoding, a stream that leaves its tile, counts no
 * codec can reach, FLOAT size mismatch), (2) the dictionary's status, (3) the PRESENT stream (decode + exact consumption),
 * (4) set bits of the validity bitmap != numValues of the data stream, (5) the DATA stream, (6) a dictionary index outside the
 * dictionary. Every slice starts on a 16-byte boundary of its buffer, like the result layout of the geometry path.
 * ------------------------------------------------------------------------------------------- */
typedef struct { void* p; uint64_t n, cap; size_t elem; } vec_t;
static void* vec_grow(vec_t* v, uint64_t add)
{
    if (v->n + add > v->cap) {
        uint64_t cap = v->cap ? v->cap : 1024;
        while (cap < v->n + add) cap *= 2;
        void* q = realloc(v->p, (size_t)cap * v->elem);
        if (!q) return NULL;
        v->p = q;
        v->cap = cap;
    }
    void* at = (uint8_t*)v->p + (size_t)v->n * v->elem;
    memset(at, 0, (size_t)add * v->elem);
    v->n += add;
    return at;
}
/* a slice of n elements that starts (and ends) on a 16-byte boundary; returns its element offset */
static int64_t vec_slice(vec_t* v, uint64_t n)
{
    uint64_t per = 16 / v->elem;
    uint64_t padded = (n + per - 1) / per * per;
    uint64_t at = v->n;
    if (padded && !vec_grow(v, padded)) return -1;
    return (int64_t)at;
}

typedef struct { uint64_t off; uint32_t nv, bl, enc; int have; } ps_t;
typedef struct { uint32_t st_present, st_data, fill_ones, present_decoded; } paux_t;
typedef struct {
    vec_t cols, aux, dicts, dict_st, validity, i64, f32, f64, bools, didx, doff;
    uint64_t payload_bytes;
    int oom;
} props_build_t;

struct prop_sink {
    const uint8_t* blob;
    uint64_t tile_end;
    uint32_t tile;
    int gen3;
    props_build_t* B;
    uint32_t layer, F, dt, ct, nlen, dict_index;
    uint64_t noff;
    int has_dict, localized;
    ps_t P, D, L, Y, pend;
    uint64_t pend_sub_off;
    uint32_t pend_sub_len;
};

static int ps_in_tile(const prop_sink_t* k, const ps_t* s) { return s->off <= k->tile_end && (uint64_t)s->bl <= k->tile_end - s->off; }

/* decode + exact consumption of one Byte-RLE stream into out[n] */
static uint32_t prop_byte_rle(const uint8_t* blob, const ps_t* s, uint32_t n, uint8_t* out)
{
    uint64_t pos = s->off;
    int32_t rc = covt_oracle_decode_byte_rle(blob, s->off + s->bl, &pos, n, out);
    if (rc != COVT_OK) return (uint32_t)rc;
    return pos == s->off + s->bl ? COVT_OK : COVT_ERR_COUNT_MISMATCH;
}
/* decode + exact consumption of one RLE stream; out64 (n values) or out32 (narrowed like the Java (int) casts, :357, :383) */
static uint32_t prop_rle(const uint8_t* blob, const ps_t* s, uint32_t n, int is_signed, int64_t* out64, int32_t* out32)
{
    int64_t* tmp = out64 ? out64 : (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
    if (!tmp) return COVT_ERR_OOM;
    uint64_t pos = s->off;
    int32_t rc = covt_oracle_decode_rle(blob, s->off + s->bl, &pos, n, is_signed, tmp);
    if (out32) for (uint32_t i = 0; i < n; i++) out32[i] = (int32_t)(uint32_t)(uint64_t)tmp[i];
    if (!out64) free(tmp);
    if (rc != COVT_OK) return (uint32_t)rc;
    return pos == s->off + s->bl ? COVT_OK : COVT_ERR_COUNT_MISMATCH;
}

static void prop_emit(prop_sink_t* k, const ps_t* Ps, const ps_t* Ds, uint64_t sub_off, uint32_t sub_len, int orphan)
{
    props_build_t* B = k->B;
    uint32_t st = COVT_OK, kind = COVT_PV_NONE, F = k->F, dt = k->dt, ct = k->ct;
    uint32_t VB = (F + 7u) / 8u;
    int use_p = 0, fill_ones = 0;
    if (orphan || dt == 0xFFu || ct > COVT_CT_ICE_MORTON_CODE || !Ds->have) st = COVT_ERR_BAD_METADATA;
    else if (dt == COVT_DT_BOOLEAN) {
        kind = COVT_PV_BOOL;
        if (!ps_in_tile(k, Ds) || (Ps->have && !ps_in_tile(k, Ps))) st = COVT_ERR_TRUNCATED;
        else if (!plausible_count((Ds->nv + 7u) / 8u, Ds->bl)) st = COVT_ERR_TRUNCATED;
        else if (Ps->have) { if (!plausible_count(VB, Ps->bl)) st = COVT_ERR_TRUNCATED; else if (Ds->nv > F) st = COVT_ERR_COUNT_MISMATCH; else use_p = 1; }
        else if (Ds->nv != F) st = COVT_ERR_COUNT_MISMATCH; /* no present stream: every feature has a value (:280-290) */
        else fill_ones = 1;
    } else if (!Ps->have) st = COVT_ERR_BAD_METADATA;
    else if (dt == COVT_DT_STRING) {
        kind = COVT_PV_DICT_INDEX;
        if (!k->has_dict) st = COVT_ERR_UNSUPPORTED_ENCODING; /* CovtParser.java:345-347 */
        else if (!k->localized && (!k->L.have || !k->Y.have)) st = COVT_ERR_BAD_METADATA;
        else if (!ps_in_tile(k, Ps) || !ps_in_tile(k, Ds) || !plausible_count(VB, Ps->bl) || !plausible_count(Ds->nv, Ds->bl)) st = COVT_ERR_TRUNCATED;
        else if (Ds->nv > F) st = COVT_ERR_COUNT_MISMATCH;
        use_p = 1;
    } else if (dt == COVT_DT_INT_64 || dt == COVT_DT_UINT_64) {
        kind = COVT_PV_I64;
        if (Ds->enc != COVT_ENC_RLE && Ds->enc != COVT_ENC_VARINT_ZIG_ZAG && Ds->enc != COVT_ENC_VARINT_DELTA_ZIG_ZAG && Ds->enc != COVT_ENC_VARINT)
            st = COVT_ERR_UNSUPPORTED_ENCODING; /* :313-315 */
        else if (!ps_in_tile(k, Ps) || !ps_in_tile(k, Ds) || !plausible_count(VB, Ps->bl) || !plausible_count(Ds->nv, Ds->bl)) st = COVT_ERR_TRUNCATED;
        else if (Ds->nv > F) st = COVT_ERR_COUNT_MISMATCH;
        use_p = 1;
    } else if (dt == COVT_DT_FLOAT || dt == COVT_DT_DOUBLE) {
        uint32_t es = dt == COVT_DT_FLOAT ? 4u : 8u;
        kind = dt == COVT_DT_FLOAT ? COVT_PV_F32 : COVT_PV_F64;
        if (!ps_in_tile(k, Ps) || !ps_in_tile(k, Ds) || !plausible_count(VB, Ps->bl)) st = COVT_ERR_TRUNCATED;
        else if ((uint64_t)Ds->nv * es != Ds->bl || Ds->nv > F) st = COVT_ERR_COUNT_MISMATCH;
        use_p = 1;
    } else st = COVT_ERR_UNSUPPORTED_ENCODING; /* "Data type not supported", :368-370 */

    covt_prop_column* c = (covt_prop_column*)vec_grow(&B->cols, 1);
    paux_t* a = (paux_t*)vec_grow(&B->aux, 1);
    if (!c || !a) { B->oom = 1; return; }
    c->tile = k->tile;
    c->layer = k->layer;
    c->name_offset = k->noff;
    c->name_length = k->nlen;
    c->sub_offset = sub_len ? sub_off : 0;
    c->sub_length = sub_len;
    c->data_type = (uint8_t)dt;
    c->column_type = (uint8_t)ct;
    c->value_kind = (uint8_t)kind;
    c->status = st;
    c->num_features = F;
    c->dictionary = (kind == COVT_PV_DICT_INDEX && k->has_dict) ? k->dict_index : 0; /* (no dictionary of its own: 0, status UNSUPPORTED_ENCODING) */
    c->validity_offset = B->validity.n;
    vec_t* vb = kind == COVT_PV_BOOL ? &B->bools : kind == COVT_PV_DICT_INDEX ? &B->didx : kind == COVT_PV_I64 ? &B->i64 :
                kind == COVT_PV_F32 ? &B->f32 : kind == COVT_PV_F64 ? &B->f64 : NULL;
    c->values_offset = vb ? vb->n : 0;
    if (st != COVT_OK) return;
    c->data_num_values = Ds->nv;
    const uint8_t* blob = k->blob;
    int64_t v_at = vec_slice(&B->validity, VB);
    uint64_t n_alloc = kind == COVT_PV_BOOL ? VB : F; /* one slot per feature (the dense values are spread out by props_finish) */
    int64_t d_at = vec_slice(vb, n_alloc);
    if (v_at < 0 || d_at < 0) { B->oom = 1; c->status = COVT_ERR_OOM; return; }
    a->fill_ones = (uint32_t)fill_ones;
    B->payload_bytes += (use_p ? Ps->bl : 0u) + (uint64_t)Ds->bl;
    if (use_p) {
        a->st_present = prop_byte_rle(blob, Ps, VB, (uint8_t*)B->validity.p + v_at);
        a->present_decoded = 1;
    }
    if (kind == COVT_PV_BOOL) a->st_data = prop_byte_rle(blob, Ds, (Ds->nv + 7u) / 8u, (uint8_t*)B->bools.p + d_at);
    else if (kind == COVT_PV_DICT_INDEX) a->st_data = prop_rle(blob, Ds, Ds->nv, 0, NULL, (int32_t*)B->didx.p + d_at);
    else if (kind == COVT_PV_I64) {
        int64_t* out = (int64_t*)B->i64.p + d_at;
        if (Ds->enc == COVT_ENC_RLE) a->st_data = prop_rle(blob, Ds, Ds->nv, dt == COVT_DT_INT_64, out, NULL); /* :299-301 */
        else {
            /* int varints widened to long (CovtParser.java:303-311, "TODO: refactor to use long instead of int") */
            int32_t* tmp = (int32_t*)malloc(((size_t)Ds->nv + 1) * sizeof(int32_t));
            uint64_t pos = Ds->off;
            int overlong = 0;
            int32_t rc;
            if (!tmp) { B->oom = 1; return; }
            if (Ds->enc == COVT_ENC_VARINT_ZIG_ZAG) rc = covt_oracle_decode_zigzag_varint(blob, Ds->off + Ds->bl, &pos, Ds->nv, tmp, &overlong);
            else if (Ds->enc == COVT_ENC_VARINT_DELTA_ZIG_ZAG) rc = covt_oracle_decode_zigzag_delta_varint(blob, Ds->off + Ds->bl, &pos, Ds->nv, tmp, &overlong);
            else rc = covt_oracle_decode_varint(blob, Ds->off + Ds->bl, &pos, Ds->nv, tmp, &overlong);
            if (rc == COVT_OK && pos != Ds->off + Ds->bl) rc = COVT_ERR_COUNT_MISMATCH;
            if (rc == COVT_OK && overlong) rc = COVT_ERR_VARINT_OVERLONG;
            for (uint32_t i = 0; i < Ds->nv; i++) out[i] = (int64_t)tmp[i];
            free(tmp);
            a->st_data = (uint32_t)rc;
        }
    } else {
        /* DecodingUtils.decodeFloatsLE :446-453: little-endian IEEE values of the present features (the hosts this runs on are little-endian) */
        memcpy((uint8_t*)vb->p + (size_t)d_at * vb->elem, blob + Ds->off, Ds->bl);
    }
}

static void prop_emit_dictionary(prop_sink_t* k)
{
    props_build_t* B = k->B;
    covt_prop_dictionary* d = (covt_prop_dictionary*)B->dicts.p + k->dict_index;
    uint32_t* lst = (uint32_t*)B->dict_st.p + k->dict_index;
    uint32_t st = COVT_OK, n = 0;
    if (!k->L.have || !k->Y.have) st = COVT_ERR_BAD_METADATA;
    else if (!ps_in_tile(k, &k->L) || !ps_in_tile(k, &k->Y)) st = COVT_ERR_TRUNCATED;
    else {
        n = k->gen3 ? k->Y.nv : k->L.nv; /* CovtParser.java:352: the DICTIONARY stream's numValues counts the entries */
        if (!plausible_count(n, k->L.bl)) { st = COVT_ERR_TRUNCATED; n = 0; }
    }
    d->n_entries = n;
    d->status = st;
    d->offsets_offset = B->doff.n;
    d->bytes_offset = st == COVT_OK ? k->Y.off : 0;
    d->n_bytes = st == COVT_OK ? k->Y.bl : 0;
    if (st != COVT_OK) return;
    int64_t at = vec_slice(&B->doff, (uint64_t)n + 1);
    if (at < 0) { B->oom = 1; d->status = COVT_ERR_OOM; return; }
    B->payload_bytes += (uint64_t)k->L.bl + k->Y.bl;
    *lst = prop_rle(k->blob, &k->L, n, 0, NULL, (int32_t*)B->doff.p + at + 1);
}

static void sink_set_layer(prop_sink_t* k, uint32_t li) { if (k) k->layer = li; }
static void sink_mark(prop_sink_t* k, sink_mark_t* m)
{
    if (!k) return;
    props_build_t* B = k->B;
    vec_t* v[11] = {&B->cols, &B->aux, &B->dicts, &B->dict_st, &B->validity, &B->i64, &B->f32, &B->f64, &B->bools, &B->didx, &B->doff};
    for (int i = 0; i < 11; i++) m->n[i] = v[i]->n;
    m->payload_bytes = B->payload_bytes;
}
static void sink_rollback(prop_sink_t* k, const sink_mark_t* m)
{
    if (!k) return;
    props_build_t* B = k->B;
    vec_t* v[11] = {&B->cols, &B->aux, &B->dicts, &B->dict_st, &B->validity, &B->i64, &B->f32, &B->f64, &B->bools, &B->didx, &B->doff};
    for (int i = 0; i < 11; i++) v[i]->n = m->n[i];
    B->payload_bytes = m->payload_bytes;
}
static void sink_begin_column(prop_sink_t* k, uint64_t noff, uint32_t nlen, uint32_t dt, uint32_t ct, uint32_t F)
{
    k->noff = noff; k->nlen = nlen; k->dt = dt; k->ct = ct; k->F = F;
    k->P.have = k->D.have = k->L.have = k->Y.have = k->pend.have = 0;
    k->has_dict = dt == COVT_DT_STRING && (ct == COVT_CT_DICTIONARY || ct == COVT_CT_LOCALIZED_DICTIONARY);
    k->localized = dt == COVT_DT_STRING && ct == COVT_CT_LOCALIZED_DICTIONARY;
    if (k->has_dict) {
        /* the record exists from here on (sub-columns refer to it); it stays BAD_METADATA if the walk never ends the column */
        k->dict_index = (uint32_t)k->B->dicts.n;
        covt_prop_dictionary* d = (covt_prop_dictionary*)vec_grow(&k->B->dicts, 1);
        uint32_t* lst = (uint32_t*)vec_grow(&k->B->dict_st, 1);
        if (!d || !lst) { k->B->oom = 1; k->has_dict = 0; return; }
        d->tile = k->tile;
        d->layer = k->layer;
        d->status = COVT_ERR_BAD_METADATA;
    }
}
static void sink_stream(prop_sink_t* k, uint32_t st, uint64_t sub_off, uint32_t sub_len, uint32_t nv, uint32_t bl, uint32_t enc, uint64_t off)
{
    ps_t s = {off, nv, bl, enc, 1};
    if (sub_len == 0) { /* present / data / length / dictionary: the first of each counts */
        if (st == COVT_ST_PRESENT) { if (!k->P.have) k->P = s; }
        else if (st == COVT_ST_DATA) { if (!k->D.have) k->D = s; }
        else if (st == COVT_ST_LENGTH) { if (!k->L.have) k->L = s; }
        else if (st == COVT_ST_DICTIONARY) { if (!k->Y.have) k->Y = s; }
        return;
    }
    if (!k->localized) return; /* a stray named stream of a plain column: hopped over */
    /* localized dictionary (gen-2b fixtures): pairs (present_<s>, <s>), adjacent, sharing the column's dictionary */
    if (st == COVT_ST_PRESENT) {
        if (k->pend.have) prop_emit(k, &k->pend, &k->pend, k->pend_sub_off, k->pend_sub_len, 1); /* no partner */
        k->pend = s;
        k->pend_sub_off = sub_off;
        k->pend_sub_len = sub_len;
        return;
    }
    if (!k->pend.have || k->pend_sub_len != sub_len || memcmp(k->blob + k->pend_sub_off, k->blob + sub_off, sub_len) != 0) return;
    prop_emit(k, &k->pend, &s, k->pend_sub_off, k->pend_sub_len, 0);
    k->pend.have = 0;
}
static void sink_end_column(prop_sink_t* k)
{
    if (k->localized) {
        if (k->pend.have) prop_emit(k, &k->pend, &k->pend, k->pend_sub_off, k->pend_sub_len, 1);
    } else prop_emit(k, &k->P, &k->D, 0, 0, 0);
    if (k->has_dict) prop_emit_dictionary(k);
}

/* after every tile: dictionary offsets, then column statuses in the documented order */
static void props_finish(props_build_t* B)
{
    covt_prop_dictionary* dicts = (covt_prop_dictionary*)B->dicts.p;
    for (uint64_t i = 0; i < B->dicts.n; i++) {
        covt_prop_dictionary* d = &dicts[i];
        if (d->status != COVT_OK) continue;
        uint32_t st = ((uint32_t*)B->dict_st.p)[i];
        if (st == COVT_OK) {
            int32_t* off = (int32_t*)B->doff.p + d->offsets_offset;
            uint64_t run = 0;
            int bad = 0;
            for (uint32_t e = 0; e < d->n_entries; e++) {
                int32_t len = off[e + 1]; /* (int)lengthStream[i], CovtParser.java:383 */
                if (len < 0) { bad = 1; len = 0; }
                run += (uint64_t)len;
                if (run > d->n_bytes) bad = 1;
                off[e + 1] = (int32_t)run;
            }
            off[0] = 0;
            if (bad) st = COVT_ERR_TRUNCATED; /* a string would run past the dictionary bytes */
            else if (run != d->n_bytes) st = COVT_ERR_COUNT_MISMATCH;
        }
        d->status = st;
    }
    covt_prop_column* cols = (covt_prop_column*)B->cols.p;
    for (uint64_t i = 0; i < B->cols.n; i++) {
        covt_prop_column* c = &cols[i];
        const paux_t* a = (const paux_t*)B->aux.p + i;
        if (c->status != COVT_OK) continue;
        uint32_t st = COVT_OK, n_valid = 0, n_entries = 0, F = c->num_features, VB = (F + 7u) / 8u;
        int present_ok = 1;
        uint8_t* validity = (uint8_t*)B->validity.p + c->validity_offset;
        if (c->value_kind == COVT_PV_DICT_INDEX) { st = dicts[c->dictionary].status; n_entries = dicts[c->dictionary].n_entries; }
        if (a->fill_ones) {
            for (uint32_t b = 0; b < F; b++) validity[b >> 3] |= (uint8_t)(1u << (b & 7));
            n_valid = F;
        } else if (a->st_present != COVT_OK) {
            present_ok = 0;
            if (st == COVT_OK) st = a->st_present;
        } else {
            for (uint32_t b = 0; b < F; b++) n_valid += (validity[b >> 3] >> (b & 7)) & 1u;
        }
        if (st == COVT_OK && n_valid != c->data_num_values) st = COVT_ERR_COUNT_MISMATCH;
        if (st == COVT_OK && a->st_data != COVT_OK) st = a->st_data;
        if (st == COVT_OK && c->value_kind == COVT_PV_DICT_INDEX) {
            const int32_t* idx = (const int32_t*)B->didx.p + c->values_offset;
            for (uint32_t v = 0; v < c->data_num_values; v++)
                if (idx[v] < 0 || (uint32_t)idx[v] >= n_entries) { st = COVT_ERR_TOPOLOGY; break; } /* ArrayIndexOutOfBounds, :357-358 */
        }
        c->status = st;
        c->num_values = present_ok ? n_valid : 0;
        /* null expansion (CovtParser.java:317-326, 331-340, 354-364: one Optional per feature): the dense values sit at the front
         * of the column's F-slot slice; spread them to their features' slots back to front, zero the others (the Arrow layout) */
        if (st == COVT_OK && n_valid != F) {
            uint32_t r = n_valid;
            if (c->value_kind == COVT_PV_BOOL) {
                uint8_t* bits = (uint8_t*)B->bools.p + c->values_offset;
                uint8_t* tmp = (uint8_t*)calloc((size_t)VB + 1, 1);
                if (!tmp) { B->oom = 1; return; }
                for (uint32_t f = F; f-- > 0;)
                    if ((validity[f >> 3] >> (f & 7)) & 1u) { r--; if ((bits[r >> 3] >> (r & 7)) & 1u) tmp[f >> 3] |= (uint8_t)(1u << (f & 7)); }
                memcpy(bits, tmp, VB);
                free(tmp);
            } else if (c->value_kind == COVT_PV_I64 || c->value_kind == COVT_PV_F64) {
                uint64_t* a = (uint64_t*)(c->value_kind == COVT_PV_I64 ? B->i64.p : B->f64.p) + c->values_offset;
                for (uint32_t f = F; f-- > 0;) a[f] = ((validity[f >> 3] >> (f & 7)) & 1u) ? a[--r] : 0;
            } else {
                uint32_t* a = (uint32_t*)(c->value_kind == COVT_PV_F32 ? B->f32.p : B->didx.p) + c->values_offset;
                for (uint32_t f = F; f-- > 0;) a[f] = ((validity[f >> 3] >> (f & 7)) & 1u) ? a[--r] : 0;
            }
        }
    }
}

int32_t covt_oracle_decode_properties(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                      const covt_tilejson* tj, uint32_t flags, covt_oracle_props** out)
{
    covt_oracle_props* R = (covt_oracle_props*)calloc(1, sizeof(covt_oracle_props));
    if (!R) return COVT_ERR_OOM;
    props_build_t B;
    memset(&B, 0, sizeof(B));
    B.cols.elem = sizeof(covt_prop_column);
    B.aux.elem = sizeof(paux_t);
    B.dicts.elem = sizeof(covt_prop_dictionary);
    B.dict_st.elem = sizeof(uint32_t);
    B.validity.elem = 1;
    B.i64.elem = sizeof(int64_t);
    B.f32.elem = sizeof(float);
    B.f64.elem = sizeof(double);
    B.bools.elem = 1;
    B.didx.elem = sizeof(int32_t);
    B.doff.elem = sizeof(int32_t);
    R->n_tiles = n_tiles;
    R->tile_status = (uint32_t*)calloc((size_t)n_tiles + 1, sizeof(uint32_t));
    uint32_t cap = 4096;
    covt_layer* layers = (covt_layer*)malloc(cap * sizeof(covt_layer));
    int32_t rc = (R->tile_status && layers) ? COVT_OK : COVT_ERR_OOM;
    for (uint32_t t = 0; rc == COVT_OK && t < n_tiles; t++) {
        prop_sink_t k;
        memset(&k, 0, sizeof(k));
        k.blob = blob;
        k.tile_end = tile_offsets[t + 1];
        k.tile = t;
        k.gen3 = container == COVT_CONTAINER_GEN3;
        k.B = &B;
        uint32_t nl = 0;
        uint64_t ep = 0;
        int32_t st;
        for (;;) {
            uint64_t n_cols0 = B.cols.n, n_dicts0 = B.dicts.n;
            if (container == COVT_CONTAINER_GEN2B) st = parse_gen2b(blob, tile_offsets[t], tile_offsets[t + 1], flags, t, layers, cap, &nl, &ep, &k);
            else if (container == COVT_CONTAINER_GEN3) st = parse_gen3(blob, tile_offsets[t], tile_offsets[t + 1], tj, flags, t, layers, cap, &nl, &ep, &k);
            else { st = COVT_ERR_INVALID_ARG; rc = st; }
            if (st != COVT_ERR_OOM || cap >= (1u << 24) || B.oom) break;
            /* more layers than the scratch table holds: grow it and walk the tile again (its columns so far are dropped; the value
             * arenas keep the abandoned slices, which nothing refers to) */
            B.cols.n = n_cols0; B.aux.n = n_cols0; B.dicts.n = n_dicts0; B.dict_st.n = n_dicts0;
            cap *= 4;
            free(layers);
            layers = (covt_layer*)malloc((size_t)cap * sizeof(covt_layer));
            if (!layers) { rc = COVT_ERR_OOM; break; }
        }
        if (st == COVT_OK && container == COVT_CONTAINER_GEN2B && ep != tile_offsets[t + 1]) st = COVT_ERR_TRUNCATED;
        R->tile_status[t] = (uint32_t)st;
    }
    free(layers);
    if (B.oom) rc = COVT_ERR_OOM;
    if (rc == COVT_OK) props_finish(&B);
    free(B.aux.p);
    free(B.dict_st.p);
    R->columns = (covt_prop_column*)B.cols.p;        R->n_columns = (uint32_t)B.cols.n;
    R->dictionaries = (covt_prop_dictionary*)B.dicts.p; R->n_dictionaries = (uint32_t)B.dicts.n;
    R->validity = (uint8_t*)B.validity.p;            R->validity_bytes = B.validity.n;
    R->i64 = (int64_t*)B.i64.p;                      R->n_i64 = B.i64.n;
    R->f32 = (float*)B.f32.p;                        R->n_f32 = B.f32.n;
    R->f64 = (double*)B.f64.p;                       R->n_f64 = B.f64.n;
    R->bools = (uint8_t*)B.bools.p;                  R->bool_bytes = B.bools.n;
    R->dict_index = (int32_t*)B.didx.p;              R->n_dict_index = B.didx.n;
    R->dict_offsets = (int32_t*)B.doff.p;            R->n_dict_offsets = B.doff.n;
    R->payload_bytes = B.payload_bytes;
    if (rc != COVT_OK) { covt_oracle_props_free(R); return rc; }
    *out = R;
    return COVT_OK;
}

void covt_oracle_props_free(covt_oracle_props* p)
{
    if (!p) return;
    free(p->tile_status); free(p->columns); free(p->dictionaries); free(p->validity); free(p->i64); free(p->f32); free(p->f64);
    free(p->bools); free(p->dict_index); free(p->dict_offsets);
    free(p);
}
