// covt_streams.cuh — warp-per-stream decoders for the four COVT column-stream codecs.
//
// Each function is executed by ONE full warp with uniform control flow and decodes one stream
// described by a StreamTask. They replace, one to one, the static codecs of the reference
// J/decoder/DecodingUtils.java (cited per function). Large delta-varint streams of the stream API take the
// multi-CTA two-pass kernels (k1a_aggregate / k1b_decode in covt_kernels.cu) instead.
#pragma once
#include "covt_device.cuh"
#include "covt_varint.cuh"

namespace covt {

// bytes of shared memory one warp needs for any stream op (u64 stage for the 64-bit id ops)
constexpr int WARP_SMEM_BYTES = STAGE_WORDS * 8;

// =================================================================================================
// 32-bit varints: DecodingUtils.decodeVarint :35, decodeZigZagVarint :46, decodeZigZagDeltaVarint :55,
// decodeZigZagDeltaVarintCoordinates :95, decodeDeltaVarintMortonCodes :394
// =================================================================================================
// Position (1-based, in chunk-window bytes) right after the `limit`-th emitted value of this chunk, or 0.
__device__ __forceinline__ uint32_t chunk_cut_position(uint32_t emit, uint32_t lane_excl, uint32_t limit)
{
    uint32_t cnt = __popc(emit);
    bool mine = limit > 0 && lane_excl < limit && limit <= lane_excl + cnt;
    uint32_t pos = 0;
    if (mine) pos = lane_id() * 16u + __fns(emit, 0, (int)(limit - lane_excl)) + 1u;
    unsigned b = __ballot_sync(FULL, mine);
    if (!b) return 0;
    return __shfl_sync(FULL, pos, __ffs(b) - 1);
}

// The Java int-varint reader (DecodingUtils.java:157-186: at most 4 bytes, the 4th byte ends the value whatever its MSB says) over
// a stream that holds a value with four continuation bytes, sequentially: COVT_ERR_TRUNCATED when the bytes end before value
// #num_values does, else COVT_ERR_VARINT_OVERLONG and the bytes consumed. Error path only (one thread).
__device__ __noinline__ void java_varint32_recount(const uint8_t* src, uint64_t byte_length, uint32_t num_values, uint32_t& status, uint32_t& consumed)
{
    uint64_t p = 0;
    for (uint32_t i = 0; i < num_values; i++) {
        for (int k = 0; k < 4; k++) {
            if (p >= byte_length) { status = COVT_ERR_TRUNCATED; consumed = (uint32_t)p; return; }
            if (!(src[p++] & 0x80u)) break;
        }
    }
    status = COVT_ERR_VARINT_OVERLONG;
    consumed = (uint32_t)p;
}

__device__ __noinline__ void warp_varint32_stream(const StreamTask& t, uint32_t* stage, StreamOutcome& out, const int post, const bool widen)
{
    const unsigned lane = lane_id();
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(t.src) & ~uintptr_t(15);
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(t.src) - a0);
    const uint64_t total = (uint64_t)head + t.byte_length;
    uint32_t carry_halo = 0, ov = 0, produced = 0, consumed = 0;
    int32_t cx = 0, cy = 0;
    for (uint64_t base = 0; base < total && produced < t.num_values; base += WARP_CHUNK_BYTES) {
        const uint64_t off = base + lane * 16u;
        uint4 win = make_uint4(0, 0, 0, 0);
        if (off < total) win = ldg_stream128(reinterpret_cast<const void*>(a0 + off));
        const uint32_t head_f = base == 0 ? head : 0u;
        uint32_t end_in_chunk = (uint32_t)umin64(WARP_CHUNK_BYTES, total - base);  // window-relative end of the stream
        const uint32_t remaining = t.num_values - produced;
        uint32_t w[4], acc, mul, lane_ov = 0;
        uint32_t halo_in = carry_halo;
        auto bounds = [&](uint32_t& lo16, uint32_t& hi16) {
            lo16 = head_f > lane * 16u ? min(16u, head_f - lane * 16u) : 0u;
            hi16 = end_in_chunk > lane * 16u ? min(16u, end_in_chunk - lane * 16u) : 0u;
        };
        uint32_t lo16, hi16;
        bounds(lo16, hi16);
        LeanLane L = lean_front(win, head_f != 0u || end_in_chunk < WARP_CHUNK_BYTES, lo16, hi16, halo_in, w, acc, mul, lane_ov);
        L.excl = warp_exclusive_scan(L.cnt, L.total);
        if (t.exact_length != 1 || L.total - head_f - (WARP_CHUNK_BYTES - end_in_chunk) > remaining) {
            // DecodingUtils "pos" semantics: bytes after the last requested value belong to somebody else (and a stream with an
            // exact length that holds MORE values than numValues is read no further either: what follows must not raise the
            // overlong flag). Find the terminator of value #remaining; if it lies in this chunk, the stream ends there.
            const int32_t first = (int32_t)L.excl - (int32_t)head_f;  // chunk-local index of the lane's first value
            const bool mine = remaining > 0u && (int32_t)remaining > first && (int32_t)remaining <= first + (int32_t)L.cnt;
            const unsigned b = __ballot_sync(FULL, mine);
            uint32_t pos = 0;
            if (mine) pos = lane * 16u + lean_nth_terminator(L.cm, remaining - (uint32_t)first);
            pos = __shfl_sync(FULL, pos, b ? __ffs(b) - 1 : 0);
            // a terminator behind the stream's bytes is one of the fake zeros, not the value asked for: the stream is too short
            if (b && pos <= end_in_chunk) {
                end_in_chunk = pos;
                consumed = (uint32_t)(base + end_in_chunk - head);
                bounds(lo16, hi16);
                halo_in = carry_halo;
                lane_ov = 0;
                L = lean_front(win, true, lo16, hi16, halo_in, w, acc, mul, lane_ov);
                L.excl = warp_exclusive_scan(L.cnt, L.total);
            }
        }
        carry_halo = halo_in;
        ov |= lane_ov;
        const uint32_t tail_f = WARP_CHUNK_BYTES - end_in_chunk;
        const uint32_t s4 = produced & 3u;  // see lean_rows4: the chunk's values are staged at A[s4 + i]
        uint32_t* A = stage + LEAN_FRONT;
        if (lane < s4) A[lane] = 0;
        lean_stage_lane(w, L.cm, acc, mul, A + s4 + L.excl - head_f, ov);
        __syncwarp();
        const uint32_t n = min(L.total - head_f - tail_f, remaining);
        lean_rows4_dispatch(post, widen, A, s4, n, t.dst, produced, cx, cy, t.num_bits, t.no_shift != 0);
        __syncwarp();
        produced += n;
    }
    out.consumed = t.exact_length == 1 ? t.byte_length : consumed;
    if (__any_sync(FULL, (ov >> 28) & 1u)) {
        // Error path. A value with four continuation bytes: the Java reader stops after the 4th byte whatever its MSB says
        // (DecodingUtils.java:157-186), so it counts MORE values than there are terminators and may reach numValues where the
        // terminator count falls short (or end earlier). Status and bytes consumed follow the Java reader exactly: one lane walks
        // the stream the way it does. (The values of such a stream are unspecified, see COVT_ERR_VARINT_OVERLONG.)
        uint32_t st = 0, cons = 0;
        if (lane == 0) java_varint32_recount(t.src, t.byte_length, t.num_values, st, cons);
        out.status = __shfl_sync(FULL, st, 0);
        cons = __shfl_sync(FULL, cons, 0);
        if (t.exact_length != 1) out.consumed = cons;
    } else if (produced < t.num_values) out.status = COVT_ERR_TRUNCATED;  // Java: ArrayIndexOutOfBounds
    else out.status = COVT_OK;
}

// =================================================================================================
// 64-bit varints for ids (ID_WIDTH 64): inverse of EncodingUtils.encodeVarints (EncodingUtils.java:39-55)
// =================================================================================================
template <bool ZZ_DELTA>
__device__ __noinline__ void warp_varint64_stream(const StreamTask& t, uint64_t* stage, StreamOutcome& out)
{
    const unsigned lane = lane_id();
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(t.src) & ~uintptr_t(15);
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(t.src) - a0);
    const uint64_t total = (uint64_t)head + t.byte_length;
    uint32_t produced = 0;
    int64_t running = 0;
    uint32_t h1 = 0, h2 = 0, h3 = 0;  // previous 12 bytes for lane 0
    bool overlong = false;
    uint32_t consumed = 0;
    for (uint64_t base = 0; base < total && produced < t.num_values; base += WARP_CHUNK_BYTES) {
        const uint64_t off = base + lane * 16u;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (off < total) w = ldg_stream128(reinterpret_cast<const void*>(a0 + off));
        const uint32_t lo16 = off >= head ? 0u : (uint32_t)umin64(16, head - off);
        const uint32_t hi16 = off >= total ? 0u : (uint32_t)umin64(16, total - off);
        uint32_t valid16 = ((1u << hi16) - 1u) & ~((1u << lo16) - 1u);
        uint32_t words[4] = {w.x, w.y, w.z, w.w};
        const uint32_t remaining = t.num_values - produced;
        uint32_t emit = (~gather_msb16(words) & 0xffffu) & valid16;
        uint32_t ctotal;
        uint32_t excl = warp_exclusive_scan(__popc(emit), ctotal);
        uint32_t cut = chunk_cut_position(emit, excl, remaining);
        if (cut) {
            consumed = (uint32_t)(base + cut - head);
            uint32_t keep = cut > lane * 16u ? min(16u, cut - lane * 16u) : 0u;
            valid16 &= (1u << keep) - 1u;
            emit &= valid16;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) words[q] &= nibble_to_bytemask((valid16 >> (4 * q)) & 0xfu);
        // halo: the 12 bytes before this lane's window
        uint32_t p1 = __shfl_up_sync(FULL, words[1], 1), p2 = __shfl_up_sync(FULL, words[2], 1), p3 = __shfl_up_sync(FULL, words[3], 1);
        if (lane == 0) { p1 = h1; p2 = h2; p3 = h3; }
        h1 = __shfl_sync(FULL, words[1], 31); h2 = __shfl_sync(FULL, words[2], 31); h3 = __shfl_sync(FULL, words[3], 31);
        const uint32_t hw[3] = {p1, p2, p3};
        uint32_t k = 0;  // trailing continuation bytes of the halo
#pragma unroll
        for (int i = 11; i >= 0; i--) {
            uint32_t b = (hw[i >> 2] >> (8 * (i & 3))) & 0xffu;
            if (k == (uint32_t)(11 - i) && (b & 0x80u)) k++;
        }
        uint64_t acc = 0;
        uint32_t shift = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if ((uint32_t)i >= 12u - k) {
                uint32_t b = (hw[i >> 2] >> (8 * (i & 3))) & 0xffu;
                acc |= (uint64_t)(b & 0x7fu) << (shift & 63u);
                shift += 7;
            }
        }
        if (k >= 10) overlong = true;
        uint32_t idx = excl;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint32_t b = (words[j >> 2] >> (8 * (j & 3))) & 0xffu;
            acc |= (uint64_t)(b & 0x7fu) << (shift & 63u);
            if (b & 0x80u) {
                if (shift >= 63) overlong = true;  // a 10th byte that still continues
                shift += 7;
            } else {
                if ((emit >> j) & 1u) {
                    if (idx < remaining) stage[stage_index(idx)] = ZZ_DELTA ? (uint64_t)zigzag_decode64(acc) : acc;
                    idx++;
                }
                acc = 0;
                shift = 0;
            }
        }
        const uint32_t n = min(ctotal, remaining);
        __syncwarp();
        if (ZZ_DELTA) {
            int64_t v[16];
            int64_t a = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t i = lane * 16 + j;
                v[j] = i < n ? (int64_t)stage[stage_index(i)] : 0;
                a += v[j];
            }
            uint64_t tot;
            int64_t pa = (int64_t)warp_exclusive_scan_u64((uint64_t)a, tot) + running;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t i = lane * 16 + j;
                pa += v[j];
                if (i < n) stage[stage_index(i)] = (uint64_t)pa;
            }
            running += (int64_t)tot;
            __syncwarp();
        }
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            uint32_t i = lane + 32u * kk;
            if (i < n) reinterpret_cast<int64_t*>(t.dst)[produced + i] = (int64_t)stage[stage_index(i)];
        }
        __syncwarp();
        produced += n;
    }
    out.consumed = consumed;
    if (produced < t.num_values) out.status = COVT_ERR_TRUNCATED;
    else if (__any_sync(FULL, overlong)) out.status = COVT_ERR_VARINT_OVERLONG;
    else out.status = COVT_OK;
}

// =================================================================================================
// ORC RLE v1: DecodingUtils.decodeRle :257 (orc-core RunLengthIntegerReader) and decodeByteRle :275/:290
// (RunLengthByteReader). The header chain is sequential; all lanes walk it redundantly (broadcast
// loads) and expand runs / store literal groups cooperatively with coalesced stores.
// =================================================================================================
__device__ __forceinline__ void warp_touch_lines(const uint8_t* src, uint32_t len)
{
    // pull the stream's cache lines in with independent loads so that the dependent header walk hits L1/L2
    uint32_t acc = 0;
    for (uint32_t o = lane_id() * 128u; o < len; o += 32u * 128u) acc += __ldg(src + o);
    if (acc == 0xffffffffu) __nanosleep(1);  // keep the loads alive
}

// unsigned LEB128 up to 10 bytes at src[pos..len); uniform across the warp. Returns false on truncation.
__device__ __forceinline__ bool read_vulong(const uint8_t* src, uint32_t len, uint32_t& pos, uint64_t& v)
{
    v = 0;
    uint32_t shift = 0;
    for (int i = 0; i < 10; i++) {
        if (pos >= len) return false;
        uint32_t b = __ldg(src + pos);
        pos++;
        v |= (uint64_t)(b & 0x7fu) << (shift & 63u);
        shift += 7;
        if (!(b & 0x80u)) break;
    }
    return true;
}

// A literal group of `lit` LEB128 values (up to 10 bytes each) at src[pos], decoded by the whole warp at once: 256 bytes are staged
// in shared memory, every lane finds the terminators among its 8 bytes, a warp scan numbers them, and each terminator's lane walks
// back to the start of its value (the window starts at a value boundary) and assembles it — instead of every lane reading every
// byte one after the other (a group of 128 values cost ~2 400 warp instructions that way, ~200 this way). Returns 0 = done,
// 1 = truncated, 2 = the bytes are not a plain sequence of <= 10-byte values (corrupt): the caller falls back to the sequential reader.
constexpr int RLE_WIN_WORDS = 68;  // 256 bytes + slack
#ifndef RLE_PARALLEL_LITERALS
#define RLE_PARALLEL_LITERALS 24u  // literal groups of at least this many values take the window
#endif
template <typename OutT, bool SIGNED>
__device__ __forceinline__ int warp_rle_literals(const uint8_t* src, uint32_t len, uint32_t& pos, uint32_t lit, OutT* dst, uint32_t done, uint32_t n,
                                                 uint32_t* win)
{
    const unsigned lane = lane_id();
    uint32_t left = lit, idx0 = done;
    auto byte_at = [&](uint32_t p) { return (win[p >> 2] >> (8u * (p & 3u))) & 0xffu; };
    while (left) {
        if (pos >= len) return 1;
        const uint32_t avail = min(256u, len - pos);
        __syncwarp();
        for (uint32_t w = lane; w < 64u; w += 32) win[w] = 4u * w < avail ? ld_u32_unaligned(src + pos + 4u * w) : 0u;
        __syncwarp();
        const uint32_t w0 = win[2u * lane], w1 = win[2u * lane + 1u];
        // bit b = byte b of this lane's 8 ends a value (and lies inside the stream)
        uint32_t t8 = ((((~w0 & 0x80808080u) >> 7) * 0x00204081u) >> 21 & 0xfu) | (((((~w1 & 0x80808080u) >> 7) * 0x00204081u) >> 21 & 0xfu) << 4);
        const uint32_t first = 8u * lane;
        t8 &= avail > first ? (avail - first >= 8u ? 0xffu : (1u << (avail - first)) - 1u) : 0u;
        uint32_t total;
        const uint32_t excl = warp_exclusive_scan((uint32_t)__popc(t8), total);
        const uint32_t take = min(total, left);
        if (take == 0u) return avail < 256u ? 1 : 2;
        bool weird = false;
        uint32_t k = 0, end_pos = 0;
        for (uint32_t bits = t8; bits; bits &= bits - 1u, k++) {
            const uint32_t idx = excl + k;
            if (idx >= take) break;
            const uint32_t q = first + (uint32_t)__ffs(bits) - 1u;
            uint32_t s0 = q, nb = 1;
            while (s0 > 0u && (byte_at(s0 - 1u) & 0x80u) && nb <= 10u) { s0--; nb++; }
            if (nb > 10u) { weird = true; break; }  // the sequential reader ends a value after 10 bytes whatever they hold
            uint64_t v = 0;
            for (uint32_t j = 0; j < nb; j++) v |= (uint64_t)(byte_at(s0 + j) & 0x7fu) << (7u * j);
            if (idx0 + idx < n) dst[idx0 + idx] = (OutT)(SIGNED ? (uint64_t)zigzag_decode64(v) : v);
            if (idx + 1u == take) end_pos = q + 1u;
        }
        if (__any_sync(FULL, weird)) return 2;
        const unsigned owner = __ballot_sync(FULL, end_pos != 0u);
        pos += __shfl_sync(FULL, end_pos, __ffs(owner) - 1);
        left -= take;
        idx0 += take;
    }
    return 0;
}

// win: RLE_WIN_WORDS words of warp-private shared memory
template <typename OutT, bool SIGNED>
__device__ __noinline__ void warp_rle_stream(const StreamTask& t, StreamOutcome& out, uint32_t* win)
{
    const unsigned lane = lane_id();
    const uint8_t* src = t.src;
    const uint32_t len = t.byte_length, n = t.num_values;
    OutT* dst = reinterpret_cast<OutT*>(t.dst);
    warp_touch_lines(src, len);
    uint32_t pos = 0, done = 0;
    uint32_t status = COVT_OK;
    while (done < n) {
        if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
        const uint32_t c = __ldg(src + pos);
        pos++;
        if (c < 0x80u) {
            const uint32_t run = c + 3u;
            if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
            const int64_t delta = (int8_t)__ldg(src + pos);
            pos++;
            uint64_t raw;
            if (!read_vulong(src, len, pos, raw)) { status = COVT_ERR_TRUNCATED; break; }
            const uint64_t base = SIGNED ? (uint64_t)zigzag_decode64(raw) : raw;
            const uint32_t m = min(run, n - done);
            for (uint32_t i = lane; i < m; i += 32) dst[done + i] = (OutT)(base + (uint64_t)i * (uint64_t)delta);
            done += m;
        } else {
            const uint32_t lit = 256u - c;
            uint32_t p2 = pos;
            // (short groups: the window costs ~150 instructions whatever it holds, a value read by all lanes ~20)
            const int r = lit >= RLE_PARALLEL_LITERALS ? warp_rle_literals<OutT, SIGNED>(src, len, p2, lit, dst, done, n, win) : 2;
            if (r == 1) { status = COVT_ERR_TRUNCATED; break; }
            if (r == 0) pos = p2;
            else {
                // short group, or corrupt bytes (a "value" longer than 10 bytes): exactly what the sequential reader does
                uint64_t mine = 0;
                bool bad = false;
                for (uint32_t i = 0; i < lit; i++) {
                    uint64_t raw;
                    if (!read_vulong(src, len, pos, raw)) { bad = true; break; }
                    if ((i & 31u) == lane) mine = SIGNED ? (uint64_t)zigzag_decode64(raw) : raw;
                    if ((i & 31u) == 31u || i + 1 == lit) {
                        const uint32_t idx = done + (i & ~31u) + lane;
                        if (lane <= (i & 31u) && idx < n) dst[idx] = (OutT)mine;
                    }
                }
                if (bad) { status = COVT_ERR_TRUNCATED; break; }
            }
            done += min(lit, n - done);
        }
    }
    out.status = status;
    out.consumed = pos;
}

__device__ __noinline__ void warp_byte_rle_stream(const StreamTask& t, StreamOutcome& out)
{
    const unsigned lane = lane_id();
    const uint8_t* src = t.src;
    const uint32_t len = t.byte_length, n = t.num_values;
    uint8_t* dst = reinterpret_cast<uint8_t*>(t.dst);
    warp_touch_lines(src, len);
    uint32_t pos = 0, done = 0;
    uint32_t status = COVT_OK;
    while (done < n) {
        if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
        const uint32_t c = __ldg(src + pos);
        pos++;
        if (c < 0x80u) {
            const uint32_t run = c + 3u;
            if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
            const uint8_t v = __ldg(src + pos);
            pos++;
            const uint32_t m = min(run, n - done);
            for (uint32_t i = lane; i < m; i += 32) dst[done + i] = v;
            done += m;
        } else {
            const uint32_t lit = 256u - c;
            if (pos + lit > len) { status = COVT_ERR_TRUNCATED; break; }
            const uint32_t m = min(lit, n - done);
            for (uint32_t i = lane; i < m; i += 32) dst[done + i] = __ldg(src + pos + i);
            pos += lit;
            done += m;
        }
    }
    out.status = status;
    out.consumed = pos;
}

// =================================================================================================
// Thread-per-stream flavours of the sequential codecs for SMALL streams (the common case in tile batches:
// ~50 values per topology / id stream). A warp-cooperative walk spends a whole warp instruction per
// header byte (ncu: k_decode_rle 82 % issue-bound at 650 warp-instructions per stream); here 32 streams advance
// in one instruction stream. Stores go out 4/8 bytes at a time and merge in L2.
// =================================================================================================
constexpr uint32_t SMALL_STREAM_VALUES = 256;  // streams up to this many values take the thread-per-stream path
constexpr uint32_t SMALL_STREAM_BYTES = 2048;

// A thread's sequential view of its stream through a two-block register window: 16 bytes per global load instead of one
// (the per-byte loads of 32 different streams made every load instruction 32 sector requests and every byte a dependent
// round trip), and the next block is already in flight when the current one runs out.
struct ByteReader {
    const uint4* base;  // 16-byte aligned block 0
    uint32_t ofs;       // offset of the stream's byte 0 inside block 0
    uint32_t blk;       // index of the block held in cur
    uint4 cur, nxt;
    __device__ __forceinline__ void init(const uint8_t* src)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        base = reinterpret_cast<const uint4*>(a & ~uintptr_t(15));
        ofs = (uint32_t)(a & 15u);
        blk = 0;
        cur = __ldg(base);
        nxt = __ldg(base + 1);
    }
    // byte `pos` of the stream; positions are visited in non-decreasing order (the batch blob is padded by 256 bytes, so the
    // look-ahead block never leaves the allocation)
    __device__ __forceinline__ uint32_t get(uint32_t pos)
    {
        const uint32_t p = pos + ofs, b = p >> 4;
        if (b != blk) {
            if (b == blk + 1) cur = nxt;
            else cur = __ldg(base + b);
            blk = b;
            nxt = __ldg(base + b + 1);
        }
        const uint32_t k = (p >> 2) & 3u;
        const uint32_t w = k < 2 ? (k == 0 ? cur.x : cur.y) : (k == 2 ? cur.z : cur.w);
        return (w >> (8u * (p & 3u))) & 0xffu;
    }
};

// The same interface, one byte per load: the RLE / Byte-RLE streams of a tile are ~40 bytes long, where the two up-front
// block loads and the extra selects of ByteReader cost more than they save (measured: k_decode_rle 0.35 -> 0.46 ms, k_decode_byte_rle
// 0.22 -> 0.34 ms per 262k tiles with the window; k_decode_varint64, ~150-byte streams, 0.49 -> 0.34 ms).
struct PlainByteReader {
    const uint8_t* src;
    __device__ __forceinline__ void init(const uint8_t* s) { src = s; }
    __device__ __forceinline__ uint32_t get(uint32_t pos) { return __ldg(src + pos); }
};

template <class Reader>
__device__ __forceinline__ bool thread_read_vulong(Reader& rd, uint32_t len, uint32_t& pos, uint64_t& v)
{
    v = 0;
    uint32_t shift = 0;
    for (int i = 0; i < 10; i++) {
        if (pos >= len) return false;
        const uint32_t b = rd.get(pos);
        pos++;
        v |= (uint64_t)(b & 0x7fu) << (shift & 63u);
        shift += 7;
        if (!(b & 0x80u)) break;
    }
    return true;
}

// DecodingUtils.decodeRle :257 — one thread, one stream
template <typename OutT>
__device__ __forceinline__ void thread_rle_stream(const StreamTask& t, bool is_signed, StreamOutcome& out)
{
    PlainByteReader rd;
    rd.init(t.src);
    const uint32_t len = t.byte_length, n = t.num_values;
    OutT* dst = reinterpret_cast<OutT*>(t.dst);
    uint32_t pos = 0, done = 0, status = COVT_OK;
    while (done < n) {
        if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
        const uint32_t c = rd.get(pos);
        pos++;
        if (c < 0x80u) {
            if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
            const int64_t delta = (int8_t)rd.get(pos);
            pos++;
            uint64_t raw;
            if (!thread_read_vulong(rd, len, pos, raw)) { status = COVT_ERR_TRUNCATED; break; }
            uint64_t v = is_signed ? (uint64_t)zigzag_decode64(raw) : raw;
            const uint32_t m = min(c + 3u, n - done);
            for (uint32_t i = 0; i < m; i++) { dst[done + i] = (OutT)v; v += (uint64_t)delta; }
            done += m;
        } else {
            const uint32_t lit = 256u - c;
            bool bad = false;
            for (uint32_t i = 0; i < lit; i++) {
                uint64_t raw;
                if (!thread_read_vulong(rd, len, pos, raw)) { bad = true; break; }
                if (done + i < n) dst[done + i] = (OutT)(is_signed ? (uint64_t)zigzag_decode64(raw) : raw);
            }
            if (bad) { status = COVT_ERR_TRUNCATED; break; }
            done += min(lit, n - done);
        }
    }
    out.status = status;
    out.consumed = pos;
}

// DecodingUtils.decodeByteRle :275/:290 — one thread, one stream; output packed into 32-bit stores (dst is 16-byte aligned)
__device__ __forceinline__ void thread_byte_rle_stream(const StreamTask& t, StreamOutcome& out)
{
    PlainByteReader rd;
    rd.init(t.src);
    const uint32_t len = t.byte_length, n = t.num_values;
    uint8_t* dst = reinterpret_cast<uint8_t*>(t.dst);
    uint32_t pos = 0, done = 0, status = COVT_OK;
    uint32_t word = 0;  // bytes [done & ~3, done) already produced
#define BRLE_PUT(val)                                                              \
    {                                                                              \
        word |= (uint32_t)(val) << (8u * (done & 3u));                             \
        done++;                                                                    \
        if ((done & 3u) == 0u) { *reinterpret_cast<uint32_t*>(dst + done - 4) = word; word = 0; } \
    }
    while (done < n) {
        if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
        const uint32_t c = rd.get(pos);
        pos++;
        if (c < 0x80u) {
            if (pos >= len) { status = COVT_ERR_TRUNCATED; break; }
            const uint32_t v = rd.get(pos);
            pos++;
            const uint32_t m = min(c + 3u, n - done);
            for (uint32_t i = 0; i < m; i++) BRLE_PUT(v);
        } else {
            const uint32_t lit = 256u - c;
            if (pos + lit > len) { status = COVT_ERR_TRUNCATED; break; }
            const uint32_t m = min(lit, n - done);
            for (uint32_t i = 0; i < m; i++) BRLE_PUT(rd.get(pos + i));
            pos += lit;
        }
    }
#undef BRLE_PUT
    // tail bytes of the last partial word (never crosses the slice: slices are 16-byte padded)
    for (uint32_t k = done & ~3u; k < done; k++) dst[k] = (uint8_t)(word >> (8u * (k & 3u)));
    out.status = status;
    out.consumed = pos;
}

// 64-bit LEB128 ids (ID_WIDTH 64) — one thread, one stream
__device__ __forceinline__ void thread_varint64_stream(const StreamTask& t, bool zz_delta, StreamOutcome& out)
{
    ByteReader rd;
    rd.init(t.src);
    const uint32_t len = t.byte_length, n = t.num_values;
    int64_t* dst = reinterpret_cast<int64_t*>(t.dst);
    uint32_t pos = 0, status = COVT_OK;
    int64_t running = 0;
    bool overlong = false;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v = 0;
        uint32_t shift = 0;
        bool ok = false;
        for (int k = 0; k < 10; k++) {
            if (pos >= len) break;
            const uint32_t b = rd.get(pos);
            pos++;
            v |= (uint64_t)(b & 0x7fu) << (shift & 63u);
            shift += 7;
            if (!(b & 0x80u)) { ok = true; break; }
            if (k == 9) { overlong = true; ok = true; }
        }
        if (!ok) { status = COVT_ERR_TRUNCATED; break; }
        if (zz_delta) { running += zigzag_decode64(v); dst[i] = running; }
        else dst[i] = (int64_t)v;
    }
    if (status == COVT_OK && overlong) status = COVT_ERR_VARINT_OVERLONG;
    out.status = status;
    out.consumed = t.exact_length == 1 ? t.byte_length : pos;
}

// =================================================================================================
// Composition(FastPFOR-256, VariableByte) over big-endian words + fused post pass:
// DecodingUtils.decodeFastPfor128ZigZagDelta :316, decodeFastPfor128DeltaCoordinates :349,
// decodeFastPfor128DeltaMortonCodes :411 (JavaFastPFOR 0.1.12)
// =================================================================================================
// k-bit value number q of an array packed in groups of 32 (BitPacking.fastpack layout) starting at word `w0`
__device__ __forceinline__ uint32_t unpack_packed(const uint8_t* base, uint32_t w0, uint32_t q, uint32_t k)
{
    if (k == 0) return 0;
    const uint32_t bo = (q & 31u) * k;
    const uint32_t wi = w0 + (q >> 5) * k + (bo >> 5);
    const uint32_t sh = bo & 31u;
    const uint32_t lo = ld_be_word(base, wi);
    const uint32_t hi = (sh + k > 32u) ? ld_be_word(base, wi + 1) : 0u;
    const uint32_t v = __funnelshift_r(lo, hi, sh);
    return k >= 32u ? v : (v & ((1u << k) - 1u));
}

// Streams larger than the shared-memory window (PFOR_SMEM_WORDS): the page is walked through global memory, but everything the
// inner loop touches is staged first with coalesced loads — the 8*b packed words of the current block, and a sliding 512-byte
// window of the byte container (consumed front to back) — so that no lane chases dependent unaligned global loads. Only the
// exception VALUES (a few percent of the stream) are fetched from global memory directly.
// wsm: PFOR_WARP_SMEM bytes of warp-private shared memory.
__device__ __noinline__ void warp_pfor_stream(const StreamTask& t, uint32_t* wsm, StreamOutcome& out, const int post)
{
    uint32_t* pk = wsm;           // [272] packed words of one block (+1 for the funnel shift)
    uint32_t* bcw = wsm + 272;    // [132] byte-container window
    uint32_t* stage = wsm + 404;  // [LEAN_STAGE_WORDS] values, 16-byte aligned
    const unsigned lane = lane_id();
    const uint8_t* base = t.src;
    const uint32_t n_words = t.byte_length / 4u;  // (int)Math.ceil(byteLength / 4): integer division, DecodingUtils.java:324
    const uint32_t n = t.num_values;
    uint32_t produced = 0;
    int32_t cx = 0, cy = 0;
    uint32_t status = COVT_OK;
    uint32_t inpos = 0;
    out.consumed = t.byte_length;
#define PFOR_FAIL(code) { status = (code); goto finish; }
    if (n_words == 0) PFOR_FAIL(COVT_ERR_TRUNCATED);  // cannot happen: streams this large have words (kept for symmetry)
    {
        const uint32_t mynvalue = ld_be_word(base, 0);
        inpos = 1;
        if (mynvalue > n || (mynvalue & 255u)) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);
        uint32_t outpos = 0;
        while (outpos < mynvalue) {
            const uint32_t thissize = min(65536u, mynvalue - outpos);
            // ---- page header (FastPFOR.decodePage) ----
            const uint32_t initpos = inpos;
            if (initpos >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t wheremeta = ld_be_word(base, initpos);
            uint64_t inexcept = (uint64_t)initpos + wheremeta;
            if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bytesize = ld_be_word(base, (uint32_t)inexcept);
            inexcept++;
            const uint64_t bc_words = ((uint64_t)bytesize + 3u) / 4u;
            if (inexcept + bc_words > n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bc_word0 = (uint32_t)inexcept;
            inexcept += bc_words;
            if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bitmap = ld_be_word(base, (uint32_t)inexcept);
            inexcept++;
            // exception array of width k lives in lane k-1: first word, size, cursor (cursors reset per page)
            uint32_t exc_base = 0, exc_size = 0, exc_ptr = 0;
            for (uint32_t rest = bitmap & ~1u; rest; rest &= rest - 1u) {
                const uint32_t k = (uint32_t)__ffs(rest);  // bit k-1 set
                if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
                const uint32_t size = ld_be_word(base, (uint32_t)inexcept);
                inexcept++;
                const uint64_t need = ((uint64_t)size * k + 31u) / 32u;
                if (inexcept + need > n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
                if (lane == k - 1) { exc_base = (uint32_t)inexcept; exc_size = size; }
                inexcept += need;
            }
            uint32_t tmpin = initpos + 1;
            uint32_t bcpos = 0;
            uint32_t bc_lo = 0, bc_hi = 0;  // byte range of the container held in bcw (bc_lo a multiple of 4)
            // byte i of the byte container: bytes are little-endian inside each (big-endian serialised) word
#define BC_BYTE(i) ((bcw[((i) - bc_lo) >> 2] >> (8u * ((i) & 3u))) & 0xffu)
            for (uint32_t run = 0, run_end = thissize / 256u; run < run_end; run++) {
                if (bcpos + 2 > bytesize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                if (bcpos < bc_lo || bcpos + 260u > bc_hi) {  // a block reads at most 2 + 1 + 255 container bytes
                    __syncwarp();
                    bc_lo = bcpos & ~3u;
                    bc_hi = bc_lo + 512u;
                    for (uint32_t i = lane; i < 128u; i += 32) {
                        const uint64_t wi = (uint64_t)bc_word0 + (bc_lo >> 2) + i;
                        bcw[i] = wi < (uint64_t)bc_word0 + bc_words ? ld_be_word(base, (uint32_t)wi) : 0u;
                    }
                    __syncwarp();
                }
                const uint32_t b = BC_BYTE(bcpos);
                const uint32_t cexcept = BC_BYTE(bcpos + 1);
                bcpos += 2;
                if (b > 32u) PFOR_FAIL(COVT_ERR_BAD_METADATA);
                if ((uint64_t)tmpin + 8ull * b > (uint64_t)initpos + wheremeta) PFOR_FAIL(COVT_ERR_TRUNCATED);
                // ---- stage the block's 8*b packed words, then unpack 256 values of b bits: value (g, lane) ----
                for (uint32_t i = lane; i < 8u * b + 1u; i += 32) pk[i] = tmpin + i < n_words ? ld_be_word(base, tmpin + i) : 0u;
                __syncwarp();
                {
                    const uint32_t bo = lane * b, wi = bo >> 5, sh = bo & 31u;
                    const uint32_t mask = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
#pragma unroll
                    for (int g = 0; g < 8; g++) stage[g * 32 + lane] = __funnelshift_r(pk[wi + g * b], pk[wi + g * b + 1], sh) & mask;
                }
                tmpin += 8u * b;
                if (cexcept > 0) {
                    if (bcpos + 1 + cexcept > bytesize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                    const uint32_t maxbits = BC_BYTE(bcpos);
                    bcpos++;
                    const int index = (int)maxbits - (int)b;
                    __syncwarp();
                    // A well-formed block lists every position once. A corrupt one may repeat a position, and the Java patch loop ORs
                    // every listed exception in: only then (two lanes of one batch on one word) must the update be atomic; a
                    // batch whose positions increase strictly (as every encoder writes them) cannot repeat one.
                    auto patch = [&](auto val_of) {
                        for (uint32_t e0 = 0; e0 < cexcept; e0 += 32) {
                            const uint32_t e = e0 + lane;
                            const bool have = e < cexcept;
                            const uint32_t pos = have ? BC_BYTE(bcpos + e) : 256u + lane;
                            const uint32_t before = __shfl_up_sync(FULL, pos, 1);
                            const bool repeated = !__all_sync(FULL, lane == 0 || pos > before);  // the encoder lists positions in increasing order
                            if (have) {
                                if (repeated) atomicOr(&stage[pos], val_of(e));
                                else stage[pos] |= val_of(e);
                            }
                        }
                    };
                    if (index == 1) {
                        patch([&](uint32_t) { return 1u << (b & 31u); });
                    } else {
                        if (index < 2 || index > 32 || !(bitmap & (1u << (index - 1)))) PFOR_FAIL(COVT_ERR_BAD_METADATA);
                        const uint32_t ebase = __shfl_sync(FULL, exc_base, index - 1);
                        const uint32_t esize = __shfl_sync(FULL, exc_size, index - 1);
                        const uint32_t eptr = __shfl_sync(FULL, exc_ptr, index - 1);
                        if (eptr + cexcept > esize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                        patch([&](uint32_t e) { return unpack_packed(base, ebase, eptr + e, (uint32_t)index) << (b & 31u); });
                        if (lane == (unsigned)(index - 1)) exc_ptr += cexcept;
                    }
                    bcpos += cexcept;
                }
                __syncwarp();
                lean_rows4_dispatch(post, false, stage, 0, 256, t.dst, produced, cx, cy, t.num_bits, t.no_shift != 0);
                __syncwarp();
                produced += 256;
            }
#undef BC_BYTE
            outpos += thissize;
            inpos = (uint32_t)inexcept;
        }
        // ---- VariableByte tail over ALL remaining words (see warp_pfor_stream_smem) ----
        uint32_t carry_halo = 0, ov = 0, vb_carry = 0;
        bool vb_long = false;
        for (uint32_t wbase = inpos; wbase < n_words; wbase += 128) {
            const uint32_t first_w = wbase + lane * 4u;
            uint4 win;
            win.x = first_w + 0 < n_words ? ld_be_word(base, first_w + 0) ^ 0x80808080u : 0u;
            win.y = first_w + 1 < n_words ? ld_be_word(base, first_w + 1) ^ 0x80808080u : 0u;
            win.z = first_w + 2 < n_words ? ld_be_word(base, first_w + 2) ^ 0x80808080u : 0u;
            win.w = first_w + 3 < n_words ? ld_be_word(base, first_w + 3) ^ 0x80808080u : 0u;
            const uint32_t valid_words = min(128u, n_words - wbase);
            const uint32_t tail_f = 512u - 4u * valid_words;
            uint32_t w[4], acc, mul;
            LeanLane L = lean_front(win, false, 0, 16, carry_halo, w, acc, mul, ov);
            vb_long |= vb_run_of_five(w, vb_carry);
            L.excl = warp_exclusive_scan(L.cnt, L.total);
            const uint32_t ctotal = L.total - tail_f;
            const uint32_t remaining = n - produced;
            if (ctotal > remaining) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);  // Java: ArrayIndexOutOfBounds
            const uint32_t s4 = produced & 3u;
            if (lane < s4) stage[lane] = 0;
            lean_stage_lane(w, L.cm, acc, mul, stage + s4 + L.excl, ov);
            __syncwarp();
            lean_rows4_dispatch(post, false, stage, s4, ctotal, t.dst, produced, cx, cy, t.num_bits, t.no_shift != 0);
            __syncwarp();
            produced += ctotal;
        }
        if (produced != n) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);
        if (__any_sync(FULL, vb_long)) PFOR_FAIL(COVT_ERR_VARINT_OVERLONG);
    }
finish:
#undef PFOR_FAIL
    out.status = status;
}

// =================================================================================================
// The same codec for streams that fit a warp's shared-memory window (PFOR_SMEM_WORDS words; every FastPFOR stream of a
// typical tile does): the payload is staged ONCE as native-endian words with coalesced loads, so that the page header, the
// byte container, the packed blocks and the exception arrays — which the format scatters over the page — are all read from
// shared memory instead of through dependent unaligned global loads, and zigzag/delta/Morton run striped (covt_varint.cuh).
// =================================================================================================
constexpr uint32_t PFOR_SMEM_WORDS = 1024;  // 4 KiB of payload per warp

__device__ __forceinline__ uint32_t sm_unpack(const uint32_t* sw, uint32_t w0, uint32_t q, uint32_t k)
{
    const uint32_t bo = (q & 31u) * k;
    const uint32_t wi = w0 + (q >> 5) * k + (bo >> 5);
    const uint32_t v = __funnelshift_r(sw[wi], sw[wi + 1], bo & 31u);
    return k >= 32u ? v : (v & ((1u << k) - 1u));
}

// sw: PFOR_SMEM_WORDS + 4 words, stage: LEAN_STAGE_WORDS words (both warp-private)
__device__ __noinline__ void warp_pfor_stream_smem(const StreamTask& t, uint32_t* sw, uint32_t* stage, StreamOutcome& out, const int post)
{
    const unsigned lane = lane_id();
    const uint8_t* base = t.src;
    const uint32_t n_words = t.byte_length / 4u;  // (int)Math.ceil(byteLength / 4): integer division, DecodingUtils.java:324
    const uint32_t n = t.num_values;
    uint32_t produced = 0;
    int32_t cx = 0, cy = 0;
    uint32_t status = COVT_OK;
    out.consumed = t.byte_length;
    // ---- stage the payload: word j of the stream, byte-swapped to native order ----
    for (uint32_t j = lane; j < n_words; j += 32) sw[j] = ld_be_word(base, j);
    if (lane < 4) sw[n_words + lane] = 0;
    __syncwarp();
#define PFOR_FAIL(code) { status = (code); goto finish; }
    if (n_words == 0) {
        // Composition.uncompress returns at once: the output stays all zeros, the post passes still run
        for (uint32_t i = lane; i < 512; i += 32) stage[i] = 0;
        __syncwarp();
        for (uint32_t b0 = 0; b0 < n; b0 += 512) {
            lean_rows4_dispatch(post, false, stage, 0, min(512u, n - b0), t.dst, b0, cx, cy, t.num_bits, t.no_shift != 0);
        }
        out.status = COVT_OK;
        return;
    }
    {
        const uint32_t mynvalue = sw[0];
        uint32_t inpos = 1;
        if (mynvalue > n || (mynvalue & 255u)) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);
        uint32_t outpos = 0;
        while (outpos < mynvalue) {
            const uint32_t thissize = min(65536u, mynvalue - outpos);
            // ---- page header (FastPFOR.decodePage) ----
            const uint32_t initpos = inpos;
            if (initpos >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t wheremeta = sw[initpos];
            uint64_t inexcept = (uint64_t)initpos + wheremeta;
            if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bytesize = sw[(uint32_t)inexcept];
            inexcept++;
            const uint64_t bc_words = ((uint64_t)bytesize + 3u) / 4u;
            if (inexcept + bc_words > n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bc_word0 = (uint32_t)inexcept;
            inexcept += bc_words;
            if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
            const uint32_t bitmap = sw[(uint32_t)inexcept];
            inexcept++;
            // exception array of width k lives in lane k-1: first word, size, cursor (cursors reset per page)
            uint32_t exc_base = 0, exc_size = 0, exc_ptr = 0;
            for (uint32_t rest = bitmap & ~1u; rest; rest &= rest - 1u) {
                const uint32_t k = (uint32_t)__ffs(rest);  // bit k-1 set
                if (inexcept >= n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
                const uint32_t size = sw[(uint32_t)inexcept];
                inexcept++;
                const uint64_t need = ((uint64_t)size * k + 31u) / 32u;
                if (inexcept + need > n_words) PFOR_FAIL(COVT_ERR_TRUNCATED);
                if (lane == k - 1) { exc_base = (uint32_t)inexcept; exc_size = size; }
                inexcept += need;
            }
            uint32_t tmpin = initpos + 1;
            uint32_t bcpos = 0;
            // byte i of the byte container: bytes are little-endian inside each (big-endian serialised) word
#define BC_BYTE(i) ((sw[bc_word0 + ((i) >> 2)] >> (8u * ((i) & 3u))) & 0xffu)
            for (uint32_t run = 0, run_end = thissize / 256u; run < run_end; run++) {
                if (bcpos + 2 > bytesize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                const uint32_t b = BC_BYTE(bcpos);
                const uint32_t cexcept = BC_BYTE(bcpos + 1);
                bcpos += 2;
                if (b > 32u) PFOR_FAIL(COVT_ERR_BAD_METADATA);
                if ((uint64_t)tmpin + 8ull * b > (uint64_t)initpos + wheremeta) PFOR_FAIL(COVT_ERR_TRUNCATED);
                // ---- unpack 256 values of b bits: value (g, lane) ----
                {
                    const uint32_t bo = lane * b, wi = tmpin + (bo >> 5), sh = bo & 31u;
                    const uint32_t mask = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
#pragma unroll
                    for (int g = 0; g < 8; g++) stage[g * 32 + lane] = __funnelshift_r(sw[wi + g * b], sw[wi + g * b + 1], sh) & mask;
                }
                tmpin += 8u * b;
                if (cexcept > 0) {
                    if (bcpos + 1 + cexcept > bytesize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                    const uint32_t maxbits = BC_BYTE(bcpos);
                    bcpos++;
                    const int index = (int)maxbits - (int)b;
                    __syncwarp();
                    // (see warp_pfor_stream: atomic only when a corrupt block repeats a position inside one batch of 32)
                    auto patch = [&](auto val_of) {
                        for (uint32_t e0 = 0; e0 < cexcept; e0 += 32) {
                            const uint32_t e = e0 + lane;
                            const bool have = e < cexcept;
                            const uint32_t pos = have ? BC_BYTE(bcpos + e) : 256u + lane;
                            const uint32_t before = __shfl_up_sync(FULL, pos, 1);
                            const bool repeated = !__all_sync(FULL, lane == 0 || pos > before);  // the encoder lists positions in increasing order
                            if (have) {
                                if (repeated) atomicOr(&stage[pos], val_of(e));
                                else stage[pos] |= val_of(e);
                            }
                        }
                    };
                    if (index == 1) {
                        patch([&](uint32_t) { return 1u << (b & 31u); });
                    } else {
                        if (index < 2 || index > 32 || !(bitmap & (1u << (index - 1)))) PFOR_FAIL(COVT_ERR_BAD_METADATA);
                        const uint32_t ebase = __shfl_sync(FULL, exc_base, index - 1);
                        const uint32_t esize = __shfl_sync(FULL, exc_size, index - 1);
                        const uint32_t eptr = __shfl_sync(FULL, exc_ptr, index - 1);
                        if (eptr + cexcept > esize) PFOR_FAIL(COVT_ERR_TRUNCATED);
                        patch([&](uint32_t e) { return sm_unpack(sw, ebase, eptr + e, (uint32_t)index) << (b & 31u); });
                        if (lane == (unsigned)(index - 1)) exc_ptr += cexcept;
                    }
                    bcpos += cexcept;
                }
                __syncwarp();
                lean_rows4_dispatch(post, false, stage, 0, 256, t.dst, produced, cx, cy, t.num_bits, t.no_shift != 0);
                __syncwarp();
                produced += 256;
            }
#undef BC_BYTE
            outpos += thissize;
            inpos = (uint32_t)inexcept;
        }
        // ---- VariableByte tail over ALL remaining words (VariableByte.uncompress): MSB SET ends a value, so flipping
        // the MSBs turns it into LEB128 for the chunk decoder (up to 5 bytes per value; the carry logic covers 4 + 1)
        uint32_t carry_halo = 0, ov = 0, vb_carry = 0;
        bool vb_long = false;
        for (uint32_t wbase = inpos; wbase < n_words; wbase += 128) {
            const uint32_t first_w = wbase + lane * 4u;
            uint4 win;
            win.x = first_w + 0 < n_words ? sw[first_w + 0] ^ 0x80808080u : 0u;
            win.y = first_w + 1 < n_words ? sw[first_w + 1] ^ 0x80808080u : 0u;
            win.z = first_w + 2 < n_words ? sw[first_w + 2] ^ 0x80808080u : 0u;
            win.w = first_w + 3 < n_words ? sw[first_w + 3] ^ 0x80808080u : 0u;
            const uint32_t valid_words = min(128u, n_words - wbase);
            const uint32_t tail_f = 512u - 4u * valid_words;  // zeroed bytes decode to fake 1-byte zeros behind the real values
            uint32_t w[4], acc, mul;
            LeanLane L = lean_front(win, false, 0, 16, carry_halo, w, acc, mul, ov);
            vb_long |= vb_run_of_five(w, vb_carry);
            L.excl = warp_exclusive_scan(L.cnt, L.total);
            const uint32_t ctotal = L.total - tail_f;
            const uint32_t remaining = n - produced;
            if (ctotal > remaining) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);  // Java: ArrayIndexOutOfBounds
            const uint32_t s4 = produced & 3u;
            if (lane < s4) stage[lane] = 0;
            lean_stage_lane(w, L.cm, acc, mul, stage + s4 + L.excl, ov);
            __syncwarp();
            lean_rows4_dispatch(post, false, stage, s4, ctotal, t.dst, produced, cx, cy, t.num_bits, t.no_shift != 0);
            __syncwarp();
            produced += ctotal;
        }
        if (produced != n) PFOR_FAIL(COVT_ERR_COUNT_MISMATCH);
        if (__any_sync(FULL, vb_long)) PFOR_FAIL(COVT_ERR_VARINT_OVERLONG);
    }
finish:
#undef PFOR_FAIL
    out.status = status;
}

}  // namespace covt
