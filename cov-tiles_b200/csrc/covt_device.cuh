// covt_device.cuh — device-side building blocks shared by the COVT decode kernels (sm_100a).
//
// Everything here is warp-granular: one warp owns one stream (or one 512-byte sub-chunk of a large
// stream) and moves data with 128-bit coalesced loads; intermediate values live in a warp-private
// shared-memory stage so that the final global stores are coalesced. No tensor cores: the path is
// byte/integer work. The 32-bit LEB128 chunk decoder lives in covt_varint.cuh.
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
    uint8_t exact_length;  // 1: byte_length is the stream's exact size (container path); 0: an upper bound (DecodingUtils "pos" semantics);
                           // 2: an upper bound, and ending anywhere else is COVT_ERR_COUNT_MISMATCH (property streams)
};
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

struct StreamOutcome {
    uint32_t status;    // covt_status
    uint32_t consumed;  // bytes the reference reader would have advanced `pos` by
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_le()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// 128-bit streaming load (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
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

// the 16 continuation bits of a lane window in byte order (64-bit varint path)
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

}  // namespace covt
