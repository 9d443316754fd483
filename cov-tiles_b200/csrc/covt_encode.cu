// covt_encode.cu — stream ENCODERS on the GPU: the inverse of every decode op of the path.
//
// Replaces, one to one, the static encoders of J/converter/EncodingUtils.java — encodeVarints :39-55 (zigzag :65-71, delta :73-93),
// encodeZigZagDeltaCoordinates :190-211, encodeRle :123-134 (orc-core 1.8.1 RunLengthIntegerWriter), encodeByteRle :136-147
// (RunLengthByteWriter), encodeFastPfor128 :149-188 (JavaFastPFOR 0.1.12 Composition(FastPFOR, VariableByte), big-endian words) —
// and GeometryUtils.encodeMorton (J/converter/GeometryUtils.java:23-32). Output is byte-identical to what those write
// (tests/test_gpu_encode.py: every stream of the 129 fixture tiles decodes and re-encodes to its own bytes on the GPU).
//
// Work unit = a PIECE: up to ENC_VARINT_PIECE values of a varint stream (a varint's bytes depend on its own value and on the value
// before it, which is in the input — so pieces are independent and a 1 GiB stream spreads over the whole GPU), or one whole
// RLE / Byte-RLE / FastPFOR stream (the ORC writers are greedy state machines: a run can only be closed knowing everything
// before it, so a stream is one thread's work and the parallelism is across the streams of a batch; a FastPFOR stream is one
// warp's work, 256-value blocks with all 32 lanes). Every piece writes into its own bounded scratch range; the host sums the piece
// lengths and k_enc_compact moves the pieces to their final, gap-free place.
#include "covt_device.cuh"

namespace covt {

// ---- value transforms (what EncodingUtils does before it writes bytes) ------------------------------------------------------
__device__ __forceinline__ uint64_t zz64(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
__device__ __forceinline__ uint32_t zz32(int32_t v) { return (uint32_t)(v >> 31) ^ ((uint32_t)v << 1); }

// GeometryUtils.encodeMorton :23-32
__device__ __forceinline__ int32_t morton_encode(int32_t x, int32_t y, uint32_t num_bits, bool no_shift)
{
    if (!no_shift) {
        const int32_t half = (int32_t)(2u << ((num_bits - 2u) & 31u)) / 2;
        x += half;
        y += half;
    }
    uint32_t code = 0;
    for (uint32_t i = 0; i < num_bits; i++) code |= (((uint32_t)x & (1u << i)) << i) | (((uint32_t)y & (1u << i)) << (i + 1));
    return (int32_t)code;
}

// The unsigned number whose LEB128 bytes value #i of the piece's stream becomes. `values` = the stream's first value.
__device__ __forceinline__ uint64_t enc_varint_code(const void* values, uint64_t i, uint32_t op, uint32_t num_bits, bool no_shift)
{
    switch (op) {
    case COVT_OP_VARINT_U32: return (uint64_t)(uint32_t) reinterpret_cast<const int32_t*>(values)[i];
    case COVT_OP_VARINT_ZZ: return zz64((int64_t) reinterpret_cast<const int32_t*>(values)[i]);
    case COVT_OP_VARINT_ZZ_DELTA: {  // encodeVarints(long[], zigZag, delta): 64-bit arithmetic on the widened ints
        const int32_t* v = reinterpret_cast<const int32_t*>(values);
        const int64_t prev = i ? (int64_t)v[i - 1] : 0;
        return zz64((int64_t)v[i] - prev);
    }
    case COVT_OP_VARINT_ZZ_DELTA_XY: {  // encodeZigZagDeltaCoordinates: int deltas per axis, int zigzag
        const int32_t* v = reinterpret_cast<const int32_t*>(values);
        const uint32_t prev = i >= 2 ? (uint32_t)v[i - 2] : 0u;
        return (uint64_t)zz32((int32_t)((uint32_t)v[i] - prev));
    }
    case COVT_OP_VARINT_DELTA_MORTON: {  // Morton codes of the vertices, deltas WITHOUT zigzag (CovtConverter.java:939-948)
        const int2* v = reinterpret_cast<const int2*>(values);
        const int64_t code = morton_encode(v[i].x, v[i].y, num_bits, no_shift);
        const int64_t prev = i ? (int64_t)morton_encode(v[i - 1].x, v[i - 1].y, num_bits, no_shift) : 0;
        return (uint64_t)(code - prev);
    }
    case COVT_OP_VARINT_U64: return (uint64_t) reinterpret_cast<const int64_t*>(values)[i];
    default: {  // COVT_OP_VARINT_ZZ_DELTA_64
        const int64_t* v = reinterpret_cast<const int64_t*>(values);
        const uint64_t prev = i ? (uint64_t)v[i - 1] : 0ull;
        return zz64((int64_t)((uint64_t)v[i] - prev));
    }
    }
}

// =================================================================================================
// LEB128 writer: one warp per piece, 128 values per trip (four consecutive values per lane). Lengths come from clz, places from a
// warp scan; the bytes cross a small shared-memory stage so that they leave as 16-byte vectors (the remainder of a trip waits at
// the front of the stage).
// =================================================================================================
constexpr int ENC_WARPS = 4;
constexpr int ENC_VPL = 4;                                       // values per lane and trip
constexpr int ENC_STAGE_BYTES = 16 + 32 * ENC_VPL * 10 + 16;

__global__ void __launch_bounds__(ENC_WARPS * 32) k_enc_varint(const uint8_t* values, EncPiece* pieces, uint32_t n_pieces, uint32_t flags)
{
    __shared__ __align__(16) uint8_t s_stage[ENC_WARPS][ENC_STAGE_BYTES];
    const unsigned lane = lane_id();
    const uint32_t g = blockIdx.x * ENC_WARPS + (threadIdx.x >> 5);
    if (g >= n_pieces) return;
    const EncPiece P = pieces[g];
    uint8_t* stage = s_stage[threadIdx.x >> 5];
    const void* stream_values = values + P.stream_values;
    const bool no_shift = (flags & COVT_FLAG_MORTON_NO_SHIFT) != 0;
    uint32_t fill = 0;
    uint64_t flushed = 0;
    uint8_t* out = P.scratch;
    for (uint32_t i0 = 0; i0 < P.num_values; i0 += 32 * ENC_VPL) {
        // four consecutive values per lane: one warp scan and one flush per 128 values
        uint64_t u[ENC_VPL];
        uint32_t len[ENC_VPL], lane_len = 0;
#pragma unroll
        for (int q = 0; q < ENC_VPL; q++) {
            const uint32_t i = i0 + ENC_VPL * lane + q;
            const bool valid = i < P.num_values;
            u[q] = valid ? enc_varint_code(stream_values, (uint64_t)P.first_index + i, P.op, P.num_bits, no_shift) : 0ull;
            len[q] = valid ? (u[q] ? (uint32_t)(70 - __clzll((long long)u[q])) / 7u : 1u) : 0u;
            lane_len += len[q];
        }
        uint32_t total;
        uint32_t at = fill + warp_exclusive_scan(lane_len, total);
#pragma unroll
        for (int q = 0; q < ENC_VPL; q++) {
            uint64_t x = u[q];
            for (uint32_t k = 0; k < len[q]; k++) {
                stage[at + k] = (uint8_t)((x & 0x7fu) | (k + 1 < len[q] ? 0x80u : 0u));
                x >>= 7;
            }
            at += len[q];
        }
        fill += total;
        __syncwarp();
        const uint32_t nvec = fill >> 4;
        for (uint32_t i = lane; i < nvec; i += 32) reinterpret_cast<uint4*>(out + flushed)[i] = reinterpret_cast<const uint4*>(stage)[i];
        const uint32_t rem = fill & 15u;
        uint8_t keep = 0;
        if (lane < rem) keep = stage[16u * nvec + lane];
        __syncwarp();
        if (lane < rem) stage[lane] = keep;
        flushed += 16ull * nvec;
        fill = rem;
        __syncwarp();
    }
    if (lane < fill) out[flushed + lane] = stage[lane];
    if (lane == 0) {
        pieces[g].byte_length = (uint32_t)(flushed + fill);
        pieces[g].status = COVT_OK;
    }
}

// =================================================================================================
// ORC RLE v1 writers (orc-core 1.8.1 RunLengthIntegerWriter / RunLengthByteWriter via EncodingUtils.encodeRle :123-134,
// encodeByteRle :136-147): one thread per stream. The writers buffer up to 128 literals; here a literal's bytes
// go out at once, the group's header byte is patched when the group closes, and the two literals that turn out to open a run
// (the writer notices a run at its third value) are taken back by rewinding the output position to where they started.
// =================================================================================================
struct ByteSink {
    uint8_t* out;
    uint32_t o;
    __device__ __forceinline__ void put(uint32_t b) { out[o++] = (uint8_t)b; }
    __device__ __forceinline__ void vulong(uint64_t v)
    {
        while (v & ~0x7full) { put(0x80u | (uint32_t)(v & 0x7fu)); v >>= 7; }
        put((uint32_t)v);
    }
};

template <bool BYTES>
__device__ void thread_enc_rle(const void* values, uint32_t n, bool is_signed, uint32_t op, ByteSink& s)
{
    // num / repeat / tail / delta / literals[0] = the writer's fields; last1 = literals[num - 1], last2 = literals[num - 2];
    // head = position of the open literal group's header byte; p1 / p2 = where the bytes of last1 / last2 start
    int num = 0, tail = 0;
    bool repeat = false;
    int64_t delta = 0, first = 0, last1 = 0, last2 = 0;
    uint32_t head = 0, p1 = 0, p2 = 0;
    auto value = [&](uint32_t i) -> int64_t {
        if (BYTES) return (int64_t) reinterpret_cast<const uint8_t*>(values)[i];
        if (op == COVT_OP_RLE_U32) return (int64_t) reinterpret_cast<const int32_t*>(values)[i];
        return reinterpret_cast<const int64_t*>(values)[i];
    };
    auto emit = [&](int64_t v) {
        if (BYTES) s.put((uint32_t)v);
        else s.vulong(is_signed ? zz64(v) : (uint64_t)v);
    };
    auto flush = [&]() {
        if (num == 0) return;
        if (repeat) {
            s.put((uint32_t)(num - 3));
            if (BYTES) s.put((uint32_t)first);
            else { s.put((uint32_t)(uint8_t)(int8_t)delta); emit(first); }
        } else {
            s.out[head] = (uint8_t)(-num);
        }
        repeat = false;
        num = 0;
        tail = 0;
    };
    auto literal = [&](int64_t v) {  // literals[num++] = v
        if (num == 0) { head = s.o; s.put(0); first = v; }
        p2 = p1; p1 = s.o;
        last2 = last1; last1 = v;
        emit(v);
        num++;
    };
    for (uint32_t i = 0; i < n; i++) {
        const int64_t v = value(i);
        if (num == 0) {
            literal(v);
            tail = 1;
        } else if (repeat) {
            const bool same = BYTES ? v == first : v == (int64_t)((uint64_t)first + (uint64_t)delta * (uint64_t)num);
            if (same) {
                num++;
                if (num == 130) flush();
            } else {
                flush();
                literal(v);
                tail = 1;
            }
        } else {
            if (BYTES) {
                if (v == last1) tail++; else tail = 1;
            } else if (tail == 1 || v != (int64_t)((uint64_t)last1 + (uint64_t)delta)) {
                delta = (int64_t)((uint64_t)v - (uint64_t)last1);
                tail = (delta < -128 || delta > 127) ? 1 : 2;
            } else {
                tail++;
            }
            if (tail == 3) {
                if (num + 1 == 3) {  // the group IS the run: take its two literals (and the header) back
                    s.o = head;
                    repeat = true;
                    num = 3;
                } else {             // close the literals before the run's first two values, which become the run
                    num -= 2;
                    const int64_t base = last2;
                    s.o = p2;
                    flush();
                    first = base;
                    repeat = true;
                    num = 3;
                }
                if (BYTES) first = v;  // (a byte run repeats v itself)
            } else {
                literal(v);
                if (num == 128) flush();
            }
        }
    }
    flush();
}

__global__ void __launch_bounds__(128) k_enc_rle(const uint8_t* values, EncPiece* pieces, uint32_t n_pieces)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pieces) return;
    const EncPiece P = pieces[g];
    ByteSink s = {P.scratch, 0u};
    const void* v = values + P.stream_values;
    if (P.op == COVT_OP_BYTE_RLE) thread_enc_rle<true>(v, P.num_values, false, P.op, s);
    else thread_enc_rle<false>(v, P.num_values, P.op == COVT_OP_RLE_S64, P.op, s);
    pieces[g].byte_length = s.o;
    pieces[g].status = COVT_OK;
}

// =================================================================================================
// Composition(FastPFOR, VariableByte).compress (JavaFastPFOR 0.1.12 via EncodingUtils.encodeFastPfor128 :149-188):
// one warp per stream. Pages of 65 536 values, blocks of 256 (eight values per lane, value 32 g + lane of the block in register g).
// A page is walked twice: pass 1 picks every block's bit width (getBestBFromData: the histogram of bit lengths in shared memory,
// the cost of every candidate width on its own lane, one warp minimum) and with it the size of every part of the page; pass 2
// packs the blocks, the byte container and the exception arrays straight into their final words. Words are built native-endian
// in a zeroed scratch range and byte-swapped at the end (the reference serialises its int[] big-endian).
// =================================================================================================
constexpr int PF_META_WORDS = 256;  // blocks per page
struct PforSmem {
    uint32_t meta[PF_META_WORDS];  // per block: bestb | bestc << 8 | maxb << 16
    uint32_t hist[36];
    uint32_t cnt[36];              // exceptions per bit-width difference (array index) in this page
    uint32_t start[36];            // first word of every exception array (its count word)
    uint32_t run[36];              // exceptions already written per array
    uint32_t pk[32];               // one group of 32 values being packed
};

// the unsigned value FastPFOR sees for value #i of the stream
__device__ __forceinline__ uint32_t enc_pfor_code(const void* values, uint64_t i, uint32_t op, uint32_t num_bits, bool no_shift)
{
    if (op == COVT_OP_PFOR_DELTA_MORTON) {
        const int2* v = reinterpret_cast<const int2*>(values);
        const uint32_t code = (uint32_t)morton_encode(v[i].x, v[i].y, num_bits, no_shift);
        const uint32_t prev = i ? (uint32_t)morton_encode(v[i - 1].x, v[i - 1].y, num_bits, no_shift) : 0u;
        return code - prev;
    }
    const int32_t* v = reinterpret_cast<const int32_t*>(values);
    if (op == COVT_OP_PFOR_ZZ_DELTA_XY) return zz32((int32_t)((uint32_t)v[i] - (i >= 2 ? (uint32_t)v[i - 2] : 0u)));
    return zz32((int32_t)((uint32_t)v[i] - (i ? (uint32_t)v[i - 1] : 0u)));
}
__device__ __forceinline__ uint32_t bits32(uint32_t v) { return v ? 32u - (uint32_t)__clz((int)v) : 0u; }

// packs value `v` (bit width b, index j of its array / group) into the LSB-first bit string that starts at word W[0]
__device__ __forceinline__ void pf_or_bits(uint32_t* W, uint64_t j, uint32_t b, uint32_t v)
{
    const uint64_t bit = j * b;
    const uint32_t sh = (uint32_t)(bit & 31u);
    uint32_t* w = W + (bit >> 5);
    atomicOr(w, v << sh);
    if (sh + b > 32u) atomicOr(w + 1, v >> (32u - sh));
}

__global__ void __launch_bounds__(ENC_WARPS * 32) k_enc_pfor(const uint8_t* values, EncPiece* pieces, uint32_t n_pieces, uint32_t flags)
{
    __shared__ PforSmem s_sm[ENC_WARPS];
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_le() ^ (1u << lane);
    const uint32_t g = blockIdx.x * ENC_WARPS + (threadIdx.x >> 5);
    if (g >= n_pieces) return;
    const EncPiece P = pieces[g];
    PforSmem& S = s_sm[threadIdx.x >> 5];
    const void* v = values + P.stream_values;
    const bool no_shift = (flags & COVT_FLAG_MORTON_NO_SHIFT) != 0;
    uint32_t* W = reinterpret_cast<uint32_t*>(P.scratch);  // zeroed by the host
    const uint32_t n = P.num_values, n256 = n & ~255u;
    uint64_t wpos = 0;
    if (n256) {
        if (lane == 0) W[0] = n256;
        wpos = 1;
        for (uint32_t p0 = 0; p0 < n256; p0 += 65536u) {
            const uint32_t nblk = min(65536u, n256 - p0) >> 8;
            // ---- pass 1: bit width of every block, sizes of the page's parts
            for (uint32_t k = lane; k < 36; k += 32) { S.cnt[k] = 0; S.run[k] = 0; }
            uint32_t bc_len = 0, packed_words = 0;
            for (uint32_t blk = 0; blk < nblk; blk++) {
                for (uint32_t k = lane; k < 36; k += 32) S.hist[k] = 0;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) atomicAdd(&S.hist[bits32(enc_pfor_code(v, (uint64_t)p0 + blk * 256u + 32u * q + lane, P.op, P.num_bits, no_shift))], 1u);
                __syncwarp();
                const uint32_t f = S.hist[lane], f32 = S.hist[32];
                const unsigned nz = __ballot_sync(FULL, f != 0u);
                const uint32_t maxb = f32 ? 32u : (nz ? 31u - (uint32_t)__clz((int)nz) : 0u);
                uint32_t suf = f;  // sum of hist[lane .. 31]
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_down_sync(FULL, suf, d);
                    if (lane + d < 32u) suf += t;
                }
                const uint32_t cex = suf - f + f32;  // values longer than `lane` bits
                // getBestBFromData: candidates maxb (no exceptions) and every b < maxb that leaves a value inside; the loop runs
                // downwards and only a strictly smaller cost wins -> the smallest cost, the LARGEST b among equals
                uint32_t key = 0xffffffffu;
                if (lane < maxb && cex < 256u) {
                    uint32_t cost = cex * 8u + cex * (maxb - lane) + lane * 256u + 8u;
                    if (maxb - lane == 1u) cost -= cex;
                    key = (cost << 6) | (63u - lane);
                }
                key = min(key, ((maxb * 256u) << 6) | (63u - maxb));
                key = __reduce_min_sync(FULL, key);
                const uint32_t bestb = 63u - (key & 63u);
                const uint32_t bestc = __shfl_sync(FULL, cex, bestb & 31u) * (bestb < maxb ? 1u : 0u);
                if (lane == 0) {
                    S.meta[blk] = bestb | (bestc << 8) | (maxb << 16);
                    if (bestc) S.cnt[maxb - bestb] += bestc;
                }
                bc_len += 2u + (bestc ? 1u + bestc : 0u);
                packed_words += 8u * bestb;
                __syncwarp();
            }
            // ---- layout of the page: header | packed blocks | bytesize | byte container | bitmap | exception arrays
            const uint64_t header_pos = wpos, packed_pos = wpos + 1, bytesize_pos = packed_pos + packed_words;
            const uint64_t bc_pos = bytesize_pos + 1, bitmap_pos = bc_pos + ((bc_len + 3u) >> 2);
            if (lane == 0) {
                uint64_t at = bitmap_pos + 1;
                uint32_t bitmap = 0;
                for (uint32_t k = 2; k <= 32; k++) {
                    S.start[k] = (uint32_t)(at - bitmap_pos);
                    if (S.cnt[k]) {
                        bitmap |= 1u << (k - 1);
                        W[at] = S.cnt[k];
                        at += 1 + (((uint64_t)S.cnt[k] * k + 31u) >> 5);
                    }
                }
                S.start[33] = (uint32_t)(at - bitmap_pos);
                W[header_pos] = (uint32_t)(bytesize_pos - header_pos);
                W[bytesize_pos] = bc_len;
                W[bitmap_pos] = bitmap;
            }
            __syncwarp();
            // ---- pass 2: pack
            uint8_t* bc = reinterpret_cast<uint8_t*>(W + bc_pos);  // (bytes little-endian inside native words, like the Java ByteBuffer)
            uint32_t bc_at = 0;
            uint64_t pk_at = packed_pos;
            for (uint32_t blk = 0; blk < nblk; blk++) {
                const uint32_t m = S.meta[blk];
                const uint32_t bestb = m & 0xffu, bestc = (m >> 8) & 0xffu, maxb = m >> 16;
                const uint32_t idx = maxb - bestb;
                if (lane == 0) {
                    bc[bc_at] = (uint8_t)bestb;
                    bc[bc_at + 1] = (uint8_t)bestc;
                    if (bestc) bc[bc_at + 2] = (uint8_t)maxb;
                }
                uint32_t seen = 0;
                const uint32_t mask = bestb == 32u ? 0xffffffffu : ((1u << bestb) - 1u);
#pragma unroll 1
                for (int q = 0; q < 8; q++) {
                    const uint32_t x = enc_pfor_code(v, (uint64_t)p0 + blk * 256u + 32u * q + lane, P.op, P.num_bits, no_shift);
                    const uint32_t hi = bestb == 32u ? 0u : x >> bestb;
                    const unsigned em = __ballot_sync(FULL, hi != 0u);
                    if (hi) {
                        const uint32_t r = seen + (uint32_t)__popc(em & lt);
                        bc[bc_at + 3 + r] = (uint8_t)(32 * q + lane);
                        // (a difference of ONE bit is not stored: the exception's value can only be 1 — JavaFastPFOR keeps positions only)
                        if (idx >= 2u) pf_or_bits(W + bitmap_pos + S.start[idx] + 1, (uint64_t)S.run[idx] + r, idx, hi);
                    }
                    seen += (uint32_t)__popc(em);
                    if (bestb) {
                        S.pk[lane] = 0;
                        __syncwarp();
                        pf_or_bits(S.pk, lane, bestb, x & mask);
                        __syncwarp();
                        if (lane < bestb) W[pk_at + lane] = S.pk[lane];
                        pk_at += bestb;
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (lane == 0 && bestc) S.run[idx] += bestc;
                bc_at += 2u + (bestc ? 1u + bestc : 0u);
                __syncwarp();
            }
            wpos = bitmap_pos + S.start[33];
            __syncwarp();
        }
    } else if (n) {
        if (lane == 0) W[0] = 0;  // Composition writes a literal 0 when FastPFOR had nothing to compress
        wpos = 1;
    }
    // ---- VariableByte tail: 7 bits per byte LSB first, MSB set on the LAST byte, zero padded to a word
    if (n > n256) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(W + wpos);
        uint32_t at0 = 0;
        for (uint32_t i0 = n256; i0 < n; i0 += 32) {
            const bool valid = i0 + lane < n;
            uint32_t x = valid ? enc_pfor_code(v, (uint64_t)i0 + lane, P.op, P.num_bits, no_shift) : 0u;
            const uint32_t len = valid ? (x ? (bits32(x) + 6u) / 7u : 1u) : 0u;
            uint32_t total;
            const uint32_t at = at0 + warp_exclusive_scan(len, total);
            for (uint32_t k = 0; k < len; k++) {
                vb[at + k] = (uint8_t)((x & 0x7fu) | (k + 1 == len ? 0x80u : 0u));
                x >>= 7;
            }
            at0 += total;
        }
        wpos += (at0 + 3u) >> 2;
    }
    __syncwarp();
    __threadfence_block();
    for (uint64_t i = lane; i < wpos; i += 32) W[i] = __byte_perm(W[i], 0, 0x0123);  // EncodingUtils.java:174-185: big-endian
    if (lane == 0) {
        pieces[g].byte_length = (uint32_t)(4u * wpos);
        pieces[g].status = COVT_OK;
    }
}

// =================================================================================================
// pieces -> their final place (warp per piece; 16-byte vectors once the destination is aligned)
// =================================================================================================
__global__ void __launch_bounds__(ENC_WARPS * 32) k_enc_compact(const EncPiece* pieces, uint32_t n_pieces, uint8_t* arena)
{
    const unsigned lane = lane_id();
    const uint32_t g = blockIdx.x * ENC_WARPS + (threadIdx.x >> 5);
    if (g >= n_pieces) return;
    const EncPiece P = pieces[g];
    const uint8_t* src = P.scratch;
    uint8_t* dst = arena + P.out_offset;
    const uint32_t n = P.byte_length;
    const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    if (lane < head) dst[lane] = src[lane];
    const uint32_t nvec = (n - head) >> 4;
    for (uint32_t i = lane; i < nvec; i += 32) {
        // 16 destination-aligned bytes from an arbitrarily aligned source: the four words that cover them, funnel-shifted
        const uint8_t* s = src + head + 16u * i;
        uint4 o;
        o.x = ld_u32_unaligned(s); o.y = ld_u32_unaligned(s + 4); o.z = ld_u32_unaligned(s + 8); o.w = ld_u32_unaligned(s + 12);
        reinterpret_cast<uint4*>(dst + head)[i] = o;
    }
    const uint32_t done = head + 16u * nvec;
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

// ---- launchers --------------------------------------------------------------------------------------------------------------
cudaError_t launch_encode_pieces(const uint8_t* values, EncPiece* varint_pieces, uint32_t n_varint, EncPiece* rle_pieces, uint32_t n_rle,
                                 EncPiece* pfor_pieces, uint32_t n_pfor, uint32_t flags, cudaStream_t st)
{
    if (n_varint) k_enc_varint<<<(n_varint + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, st>>>(values, varint_pieces, n_varint, flags);
    if (n_rle) k_enc_rle<<<(n_rle + 127) / 128, 128, 0, st>>>(values, rle_pieces, n_rle);
    if (n_pfor) k_enc_pfor<<<(n_pfor + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, st>>>(values, pfor_pieces, n_pfor, flags);
    return cudaGetLastError();
}
cudaError_t launch_encode_compact(const EncPiece* pieces, uint32_t n_pieces, uint8_t* arena, cudaStream_t st)
{
    if (n_pieces) k_enc_compact<<<(n_pieces + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, st>>>(pieces, n_pieces, arena);
    return cudaGetLastError();
}

}  // namespace covt
