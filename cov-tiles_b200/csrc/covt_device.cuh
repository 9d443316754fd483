// covt_device.cuh — device-side building blocks shared by the COVT decode kernels (sm_100a).
//
// Everything here is warp-granular: one warp owns one stream (or one 512-byte sub-chunk of a large
// stream) and moves data with 128-bit coalesced loads; intermediate values live in a warp-private
// shared-memory stage so that the final global stores are coalesced. No tensor cores: the path is
// byte/integer work bound by HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/covt_b200.h"
#include "covt_internal.h"

namespace covt {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARP_CHUNK_BYTES = 512;        // 32 lanes x 16 B
constexpr int STAGE_ROW = 34;                // words per row of the transposed stage (34 keeps every access pattern conflict-free)
constexpr int STAGE_WORDS = 16 * STAGE_ROW;  // value i lives at row i & 15, column i >> 4

// What one decode call has to do; derived from a covt_layer slot or a covt_stream_desc.
struct StreamTask {
    const uint8_t* src;    // first payload byte (any alignment)
    void* dst;             // output slice (16-byte aligned)
    uint32_t byte_length;  // bytes available to the stream
    uint32_t num_values;   // values to produce (vertices for the Morton ops)
    uint8_t op;            // covt_op
    uint8_t num_bits;      // Morton bits
    uint8_t no_shift;      // COVT_FLAG_MORTON_NO_SHIFT
    uint8_t exact_length;  // 1: byte_length is the stream's exact size (container path); 0: an upper bound (DecodingUtils "pos" semantics)
};
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

struct StreamOutcome {
    uint32_t status;    // covt_status
    uint32_t consumed;  // bytes the reference reader would have advanced `pos` by
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// 128-bit streaming load (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream128(void* p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// unaligned little-endian 32-bit load (reads the two aligned words that cover it; buffers are padded)
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p)
{
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    unsigned s = (unsigned)(a & 3u) * 8u;
    uint32_t lo = __ldg(w);
    uint32_t hi = s ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, s);
}
// word i of a big-endian serialised int[] that starts at `base` (FastPFOR payloads, DecodingUtils.java:319-327)
__device__ __forceinline__ uint32_t ld_be_word(const uint8_t* base, uint32_t i)
{
    return __byte_perm(ld_u32_unaligned(base + 4ull * i), 0, 0x0123);
}

__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t v, uint32_t& total)
{
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane_id() >= (unsigned)d) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}
__device__ __forceinline__ int32_t warp_exclusive_scan_i32(int32_t v, int32_t& total)
{
    uint32_t t;
    uint32_t e = warp_exclusive_scan((uint32_t)v, t);
    total = (int32_t)t;
    return (int32_t)e;
}
__device__ __forceinline__ uint64_t warp_exclusive_scan_u64(uint64_t v, uint64_t& total)
{
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(FULL, x, d);
        if (lane_id() >= (unsigned)d) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

// 4 validity bits -> 4 byte masks
__device__ __forceinline__ uint32_t nibble_to_bytemask(uint32_t nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xffu; }

__device__ __forceinline__ int32_t zigzag_decode32(uint32_t e) { return (int32_t)((e >> 1) ^ (0u - (e & 1u))); }
__device__ __forceinline__ int64_t zigzag_decode64(uint64_t e) { return (int64_t)((e >> 1) ^ (0ull - (e & 1ull))); }

// GeometryUtils.decodeMorton (J/converter/GeometryUtils.java:34-47) for num_bits <= 16: gather even bits.
__device__ __forceinline__ uint32_t compact_even_bits(uint32_t v)
{
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0f0f0f0fu;
    v = (v | (v >> 4)) & 0x00ff00ffu;
    v = (v | (v >> 8)) & 0x0000ffffu;
    return v;
}
// Exact restatement incl. the Java long/int quirks for num_bits > 16 (sign-extended code, 1L << 2i).
__device__ __forceinline__ int32_t morton_compact_java(int32_t code, uint32_t num_bits)
{
    if (num_bits <= 16) return (int32_t)(compact_even_bits((uint32_t)code) & ((1u << num_bits) - 1u));
    int64_t lc = (int64_t)code;
    int32_t c = 0;
    for (uint32_t i = 0; i < num_bits; i++) {
        int64_t bit = lc & (int64_t)(1ull << ((2 * i) & 63));
        c = (int32_t)((int64_t)c | (bit >> (i & 63)));
    }
    return c;
}
__device__ __forceinline__ int2 morton_decode(int32_t code, uint32_t num_bits, bool no_shift)
{
    int32_t cx = morton_compact_java(code, num_bits);
    int32_t cy = morton_compact_java(code >> 1, num_bits);
    if (no_shift) {  // fixture-era converter: sign-extend the num_bits-bit value (SURVEY §A.6 MORTON_SHIFT)
        if (num_bits >= 1 && num_bits < 32) {
            int sh = 32 - (int)num_bits;
            cx = (int32_t)((uint32_t)cx << sh) >> sh;
            cy = (int32_t)((uint32_t)cy << sh) >> sh;
        }
        return make_int2(cx, cy);
    }
    int32_t half = (int32_t)(2u << ((num_bits - 2u) & 31u)) / 2;
    return make_int2((int32_t)((uint32_t)cx - (uint32_t)half), (int32_t)((uint32_t)cy - (uint32_t)half));
}

// Transposed stage: value i of a chunk (i < 512) lives at (i & 15) * 34 + (i >> 4). The three access patterns of a chunk
// are then bank-conflict free with compile-time offsets: blocked (lane*16 + j -> j*34 + lane), blocked by 8
// (FastPFOR blocks) and striped (lane + 32k -> (lane & 15)*34 + (lane >> 4) + 2k).
__device__ __forceinline__ uint32_t stage_index(uint32_t i) { return (i & 15u) * STAGE_ROW + (i >> 4); }

// value post-processing selectors (PostKind) live in covt_internal.h: the host picks them for large streams

__device__ __forceinline__ int post_kind_of_op(uint32_t op)
{
    switch (op) {
    case COVT_OP_VARINT_U32: return POST_PLAIN;
    case COVT_OP_VARINT_ZZ: return POST_ZZ;
    case COVT_OP_VARINT_ZZ_DELTA: case COVT_OP_PFOR_ZZ_DELTA: return POST_ZZ_DELTA;
    case COVT_OP_VARINT_ZZ_DELTA_XY: case COVT_OP_PFOR_ZZ_DELTA_XY: return POST_ZZ_DELTA_XY;
    case COVT_OP_VARINT_DELTA_MORTON: case COVT_OP_PFOR_DELTA_MORTON: return POST_DELTA_MORTON;
    default: return POST_PLAIN;
    }
}

// Running state of a delta chain that is carried from chunk to chunk of one stream.
struct DeltaCarry {
    int32_t x, y;       // running sums (x only for single-accumulator ops)
    uint32_t produced;  // values emitted so far
};

// -----------------------------------------------------------------------------------------------
// One 512-byte chunk of 32-bit varints, decoded by one warp.
//   w          : this lane's 16 bytes as loaded (bytes outside the stream may hold anything)
//   valid16    : bit j set <=> byte j of this lane's window lies inside the stream
//   carry_halo : in: the 4 bytes preceding lane 0's window; out: lane 31's last word (next chunk's halo)
//   limit      : values with chunk-local index >= limit are counted but not staged
//   VB         : VariableByte convention (MSB SET terminates, up to 5 bytes) instead of
//                LEB128 with the Java reader's 4-byte cap (DecodingUtils.java:157-186)
// Stages raw (zigzag-decoded when ZZ) values compacted into stage[]; returns the lane's emit mask, its
// warp-exclusive count and the chunk total; flags values longer than the Java cap.
// -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather_msb16(const uint32_t words[4])
{
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t m = words[q] & 0x80808080u;
        uint32_t g = ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
        r |= g << (4 * q);
    }
    return r;
}

// Common prologue of a chunk: zero the bytes outside the stream, find this lane's terminators and halo.
template <bool VB>
__device__ __forceinline__ void varint32_chunk_prologue(uint4 w, uint32_t valid16, uint32_t& carry_halo, uint32_t words[4],
                                                        uint32_t& emit, uint32_t& acc, uint32_t& shift, bool& overlong)
{
    words[0] = w.x; words[1] = w.y; words[2] = w.z; words[3] = w.w;
#pragma unroll
    for (int q = 0; q < 4; q++) words[q] &= nibble_to_bytemask((valid16 >> (4 * q)) & 0xfu);
    const uint32_t msb16 = gather_msb16(words);
    emit = (VB ? msb16 : (~msb16 & 0xffffu)) & valid16;
    uint32_t halo = __shfl_up_sync(FULL, words[3], 1);
    if (lane_id() == 0) halo = carry_halo;
    carry_halo = __shfl_sync(FULL, words[3], 31);
    // carry-in from the halo: k trailing continuation bytes
    const uint32_t hcont = VB ? (~halo & 0x80808080u) : (halo & 0x80808080u);
    const uint32_t hterm = hcont ^ 0x80808080u;
    const uint32_t k = hterm ? (uint32_t)(__clz(hterm) >> 3) : 4u;
    acc = 0;
    shift = 0;
    if (k) {
        uint32_t hv = k >= 4 ? halo : (halo >> (8u * (4u - k)));
        acc = (hv & 0x7fu) | ((hv >> 1) & 0x3f80u) | ((hv >> 2) & 0x1fc000u) | ((hv >> 3) & 0x0fe00000u);
        if (k < 4) acc &= (1u << (7u * k)) - 1u;
        shift = 7u * k;
    }
    if (!VB) {
        // Java cap: a value may not have 4 continuation bytes. Look at halo bytes 1..3 + the 16 window bytes.
        const uint32_t h3 = ((hcont >> 15) & 1u) | ((hcont >> 22) & 2u) | ((hcont >> 29) & 4u);
        const uint32_t c19 = h3 | (msb16 << 3);
        if (c19 & (c19 >> 1) & (c19 >> 2) & (c19 >> 3)) overlong = true;
    }
}

// Decode + stage. RAW values are staged (zigzag is applied by the delta pass or the copy-out): the byte loop is
// branch-free — a predicated store plus selects — because every byte position has a terminator in SOME lane.
template <bool VB, bool LIMIT>
__device__ __forceinline__ void varint32_chunk_decode(uint4 w, uint32_t valid16, uint32_t& carry_halo, uint32_t limit,
                                                      uint32_t* stage, uint32_t& emit, uint32_t& lane_excl,
                                                      uint32_t& chunk_total, bool& overlong)
{
    uint32_t words[4], acc, shift;
    varint32_chunk_prologue<VB>(w, valid16, carry_halo, words, emit, acc, shift, overlong);
    lane_excl = warp_exclusive_scan(__popc(emit), chunk_total);
    uint32_t idx = lane_excl;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t b = __byte_perm(words[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3));
        const uint32_t v = acc | ((b & 0x7fu) << (shift & 31u));
        const bool term = VB ? (b & 0x80u) != 0u : (b & 0x80u) == 0u;
        const bool em = (emit >> j) & 1u;
        if (em && (!LIMIT || idx < limit)) stage[stage_index(idx)] = v;
        idx += em ? 1u : 0u;
        acc = term ? 0u : v;
        shift = term ? 0u : shift + 7u;
    }
}

// Pass-1 flavour for large streams: only this lane's (count, sum at even local positions, sum at odd local positions).
template <bool ZZ>
__device__ __forceinline__ void varint32_chunk_sums(uint4 w, uint32_t valid16, uint32_t head_fakes, uint32_t& carry_halo,
                                                    uint32_t& cnt, int32_t& a, int32_t& b, bool& overlong)
{
    uint32_t words[4], emit, acc, shift;
    varint32_chunk_prologue<false>(w, valid16, carry_halo, words, emit, acc, shift, overlong);
    cnt = __popc(emit);
    // Zeroed bytes outside the stream look like 1-byte zeros: they add nothing to the sums and only toggle the parity,
    // which is undone below (trailing ones never matter, leading ones are counted in head_fakes).
    int32_t s0 = 0, s1 = 0;  // s0 = accumulator of the NEXT value's parity class
    uint32_t toggles = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t bb = __byte_perm(words[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3));
        const uint32_t v = acc | ((bb & 0x7fu) << (shift & 31u));
        const bool term = (bb & 0x80u) == 0u;
        const int32_t d = ZZ ? zigzag_decode32(v) : (int32_t)v;
        const int32_t t = s0 + d;
        s0 = term ? s1 : s0;
        s1 = term ? t : s1;
        toggles += term ? 1u : 0u;
        acc = term ? 0u : v;
        shift = term ? 0u : shift + 7u;
    }
    // after an even number of toggles s0 is again the class of local position 0
    const bool flip = (toggles ^ head_fakes) & 1u;  // leading fakes shifted every real value by head_fakes positions
    const bool odd = toggles & 1u;
    const int32_t e0 = odd ? s1 : s0, e1 = odd ? s0 : s1;  // sums at even / odd positions counted from the first (fake or real) value
    (void)flip;
    a = (head_fakes & 1u) ? e1 : e0;
    b = (head_fakes & 1u) ? e0 : e1;
}

// zz_at_load: the staged values are still zigzag-encoded (FastPFOR path, DecodingUtils.java:335-343).
// `post` is warp-uniform, so the branches below do not diverge.
template <int PER>
__device__ __forceinline__ void warp_delta_pass(uint32_t* stage, uint32_t n, DeltaCarry& carry, const int post, const bool zz_at_load)
{
    if (post == POST_PLAIN || post == POST_ZZ) return;
    const unsigned lane = lane_id();
    int32_t v[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        uint32_t i = lane * PER + j;
        uint32_t raw = i < n ? stage[stage_index(i)] : 0u;
        v[j] = zz_at_load ? zigzag_decode32(raw) : (int32_t)raw;
    }
    if (post == POST_ZZ_DELTA_XY) {
        // global index parity of position i is (produced + j) & 1 because lane*PER is even
        const bool swap = carry.produced & 1u;
        int32_t a = 0, b = 0;  // a: even j, b: odd j
#pragma unroll
        for (int j = 0; j < PER; j++) {
            if (j & 1) b += v[j]; else a += v[j];
        }
        int32_t ta, tb;
        int32_t ea = warp_exclusive_scan_i32(a, ta);
        int32_t eb = warp_exclusive_scan_i32(b, tb);
        int32_t pa = ea + (swap ? carry.y : carry.x);
        int32_t pb = eb + (swap ? carry.x : carry.y);
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint32_t i = lane * PER + j;
            if (j & 1) { pb += v[j]; v[j] = pb; } else { pa += v[j]; v[j] = pa; }
            if (i < n) stage[stage_index(i)] = (uint32_t)v[j];
        }
        if (swap) { carry.y += ta; carry.x += tb; } else { carry.x += ta; carry.y += tb; }
    } else {
        int32_t a = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) a += v[j];
        int32_t ta;
        int32_t pa = warp_exclusive_scan_i32(a, ta) + carry.x;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            uint32_t i = lane * PER + j;
            pa += v[j];
            if (i < n) stage[stage_index(i)] = (uint32_t)pa;
        }
        carry.x += ta;
    }
}

// Coalesced copy of stage[0..n) to dst[first ..]: int32, widened to int64 (ids), or expanded Morton (x,y) pairs.
enum CopyKind { COPY_I32 = 0, COPY_I64 = 1, COPY_MORTON = 2, COPY_I32_ZZ = 3 };
template <int PER>
__device__ __forceinline__ void warp_copy_out(const uint32_t* stage, uint32_t n, void* dst, uint64_t first, const int kind,
                                              uint32_t num_bits, bool no_shift)
{
    const unsigned lane = lane_id();
    if (kind == COPY_MORTON) {
#pragma unroll 4
        for (int k = 0; k < PER; k++) {
            uint32_t i = lane + 32u * k;
            if (i < n) reinterpret_cast<int2*>(dst)[first + i] = morton_decode((int32_t)stage[stage_index(i)], num_bits, no_shift);
        }
    } else if (kind == COPY_I64) {
#pragma unroll 4
        for (int k = 0; k < PER; k++) {
            uint32_t i = lane + 32u * k;
            if (i < n) reinterpret_cast<int64_t*>(dst)[first + i] = (int64_t)(int32_t)stage[stage_index(i)];
        }
    } else if (kind == COPY_I32_ZZ) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            uint32_t i = lane + 32u * k;
            if (i < n) reinterpret_cast<int32_t*>(dst)[first + i] = zigzag_decode32(stage[stage_index(i)]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            uint32_t i = lane + 32u * k;
            if (i < n) reinterpret_cast<int32_t*>(dst)[first + i] = (int32_t)stage[stage_index(i)];
        }
    }
}

}  // namespace covt
