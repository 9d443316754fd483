// covt_varint.cuh — the 32-bit LEB128 chunk decoder shared by the per-stream varint kernel and the large-stream kernels.
//
// DecodingUtils.decodeVarint :35/:157-186, decodeZigZagVarint :46, decodeZigZagDeltaVarint :55,
// decodeZigZagDeltaVarintCoordinates :95, decodeDeltaVarintMortonCodes :394 (J/decoder/DecodingUtils.java).
//
// One warp owns one 512-byte window (32 lanes x one 128-bit load). The kernels are issue-bound, not HBM-bound, until the
// per-byte instruction count is small (profiles/r01_b_k1_two_pass_ncu_summary.txt: 642 + 931 warp instructions per chunk), so
// this version is written for instruction count:
//   * continuation bits are gathered into a SCATTERED 16-bit mask (4 shift+and pairs, no per-nibble transposition): the byte
//     loop tests compile-time bit positions, `16 - popc` is the lane's value count;
//   * the byte loop keeps (acc, mul = 1 << 7k): value = payload * mul + acc is one IMAD; a terminator stores the value at a
//     running shared-memory pointer (linear stage, no index arithmetic);
//   * bytes outside the stream are zeroed: they decode to fake 1-byte zeros that add nothing to the sums; leading fakes land
//     in a front slack of the stage, trailing fakes behind the real values;
//   * zigzag + delta run on the STRIPED layout (value i in lane i & 31): a stride-2 (x,y) or stride-1 warp scan per row of 32
//     values, followed directly by the coalesced global store — the values cross shared memory once.
#pragma once
#include "covt_device.cuh"

namespace covt {

constexpr int LEAN_FRONT = 16;                          // front slack of the stage (leading fake values)
constexpr int LEAN_STAGE_WORDS = LEAN_FRONT + 512 + 16;  // + back slack

// bit (8 * b + q) = MSB of byte b of word q: byte j of the lane window (j = 4q + b) is a continuation byte
__device__ __forceinline__ uint32_t cont_mask_scattered(const uint32_t w[4])
{
    return ((w[0] >> 7) & 0x01010101u) | ((w[1] >> 6) & 0x02020202u) | ((w[2] >> 5) & 0x04040404u) | ((w[3] >> 4) & 0x08080808u);
}
__host__ __device__ constexpr uint32_t cont_bit_of_byte(int j) { return 1u << (8 * (j & 3) + (j >> 2)); }

// zero the bytes of the lane window outside [lo16, hi16)
__device__ __forceinline__ void lean_mask_window(uint32_t w[4], uint32_t lo16, uint32_t hi16)
{
    const uint32_t valid16 = ((1u << hi16) - 1u) & ~((1u << lo16) - 1u);
#pragma unroll
    for (int q = 0; q < 4; q++) w[q] &= nibble_to_bytemask((valid16 >> (4 * q)) & 0xfu);
}

// Partial value carried into a lane from the 4 bytes before its window: k trailing continuation bytes (k <= 4).
// ov collects every `mul` seen: bit 28 set <=> some value had 4 continuation bytes (the Java reader's cap, :157-186).
__device__ __forceinline__ void lean_carry_in(uint32_t halo, uint32_t& acc, uint32_t& mul, uint32_t& ov)
{
    const uint32_t hterm = ~halo & 0x80808080u;
    const uint32_t k = hterm ? (uint32_t)(__clz(hterm) >> 3) : 4u;
    const uint32_t hv = __funnelshift_rc(halo, 0u, 32u - 8u * k);  // the k trailing bytes, oldest in the low byte
    mul = 1u << (7u * k);
    acc = ((hv & 0x7fu) | ((hv >> 1) & 0x3f80u) | ((hv >> 2) & 0x1fc000u) | ((hv >> 3) & 0x0fe00000u)) & (mul - 1u);
    ov |= mul;
}

// halo of every lane = last word of the previous lane; lane 0 takes the chunk's carry; returns the next chunk's carry
__device__ __forceinline__ uint32_t lean_halo(uint32_t last_word, uint32_t& carry_halo)
{
    uint32_t halo = __shfl_up_sync(FULL, last_word, 1);
    if (lane_id() == 0) halo = carry_halo;
    carry_halo = __shfl_sync(FULL, last_word, 31);
    return halo;
}

// The byte loop, staging flavour: stores every value whose terminator lies in this lane at *sp++ (raw, not zigzag-decoded).
__device__ __forceinline__ void lean_stage_lane(const uint32_t w[4], uint32_t cm, uint32_t acc, uint32_t mul, uint32_t* sp, uint32_t& ov)
{
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t p = (w[j >> 2] >> (8 * (j & 3))) & 0x7fu;
        const uint32_t v = p * mul + acc;
        const bool term = (cm & cont_bit_of_byte(j)) == 0u;
        if (term) *sp = v;
        sp += term ? 1 : 0;
        acc = term ? 0u : v;
        mul = term ? 1u : mul << 7;
        ov |= mul;
    }
}

// The byte loop, aggregate flavour: two accumulators that swap at every terminator, so that no parity bookkeeping and no
// division is needed (everything stays exact mod 2^32 like the Java int sums). On return `cur` is the sum of the class of the
// lane's NEXT value and `oth` the sum of the other class: after an even number of values cur = the sum at even lane-local
// positions.
template <bool ZZ>
__device__ __forceinline__ void lean_sum_lane(const uint32_t w[4], uint32_t cm, uint32_t acc, uint32_t mul, int32_t& cur_out, int32_t& oth_out, uint32_t& ov)
{
    int32_t cur = 0, oth = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t p = (w[j >> 2] >> (8 * (j & 3))) & 0x7fu;
        const uint32_t v = p * mul + acc;
        const bool term = (cm & cont_bit_of_byte(j)) == 0u;
        const int32_t t = cur + (ZZ ? zigzag_decode32(v) : (int32_t)v);
        cur = term ? oth : cur;
        oth = term ? t : oth;
        acc = term ? 0u : v;
        mul = term ? 1u : mul << 7;
        ov |= mul;
    }
    cur_out = cur;
    oth_out = oth;
}

struct LeanLane {
    uint32_t cm;     // scattered continuation mask of the lane window (after masking)
    uint32_t cnt;    // terminators in the lane window, fakes included
    uint32_t excl;   // warp-exclusive prefix of cnt
    uint32_t total;  // warp total of cnt
};

// Common front end: mask, continuation bits, counts, carry-in. `win` = the lane's 16 bytes as loaded.
__device__ __forceinline__ LeanLane lean_front(uint4 win, bool partial, uint32_t lo16, uint32_t hi16, uint32_t& carry_halo,
                                               uint32_t w[4], uint32_t& acc, uint32_t& mul, uint32_t& ov)
{
    w[0] = win.x; w[1] = win.y; w[2] = win.z; w[3] = win.w;
    if (partial) lean_mask_window(w, lo16, hi16);  // warp-uniform branch
    LeanLane L;
    L.cm = cont_mask_scattered(w);
    L.cnt = 16u - (uint32_t)__popc(L.cm);
    const uint32_t halo = lean_halo(w[3], carry_halo);
    lean_carry_in(halo, acc, mul, ov);
    return L;
}

// Position (1-based byte offset inside the lane window) of the lane's m-th terminator, m >= 1. Rare path.
__device__ __forceinline__ uint32_t lean_nth_terminator(uint32_t cm, uint32_t m)
{
    uint32_t seen = 0;
    for (int j = 0; j < 16; j++) {
        if ((cm & cont_bit_of_byte(j)) == 0u && ++seen == m) return (uint32_t)j + 1u;
    }
    return 0;
}

// ---- striped zigzag / delta / store ------------------------------------------------------------------------------
// stage[i], i < n: raw values of one chunk in stream order. `index0` = stream index of stage[0]; (cx, cy) = running sums of
// the values before stage[0] at even / odd STREAM positions (cy unused for single-accumulator posts). Updated on return.
template <int POST, bool WIDEN>
__device__ __forceinline__ void lean_rows(const uint32_t* stage, uint32_t n, void* dst, uint64_t index0, int32_t& cx, int32_t& cy,
                                          uint32_t num_bits, bool no_shift)
{
    const unsigned lane = lane_id();
    if (POST == POST_PLAIN || POST == POST_ZZ) {
        for (uint32_t r = 0; r < n; r += 32) {
            const uint32_t i = r + lane;
            if (i < n) {
                const uint32_t raw = stage[i];
                const int32_t v = POST == POST_ZZ ? zigzag_decode32(raw) : (int32_t)raw;
                if (WIDEN) reinterpret_cast<int64_t*>(dst)[index0 + i] = (int64_t)v;
                else reinterpret_cast<int32_t*>(dst)[index0 + i] = v;
            }
        }
        return;
    }
    if (POST == POST_ZZ_DELTA_XY) {
        // lane parity == chunk-local index parity (rows are 32 wide); stream parity adds index0
        const bool odd_first = (index0 & 1ull) != 0ull;
        const bool is_y = ((lane & 1u) != 0u) != odd_first;
        int32_t carry = is_y ? cy : cx;
        for (uint32_t r = 0; r < n; r += 32) {
            const uint32_t i = r + lane;
            int32_t v = i < n ? zigzag_decode32(stage[i]) : 0;
#pragma unroll
            for (int d = 2; d < 32; d <<= 1) {
                const int32_t t = __shfl_up_sync(FULL, v, d);
                if (lane >= (unsigned)d) v += t;
            }
            v += carry;
            if (i < n) reinterpret_cast<int32_t*>(dst)[index0 + i] = v;
            carry = __shfl_sync(FULL, v, 30 + (lane & 1u));
        }
        const int32_t c_even_lane = __shfl_sync(FULL, carry, 0), c_odd_lane = __shfl_sync(FULL, carry, 1);
        cx = odd_first ? c_odd_lane : c_even_lane;
        cy = odd_first ? c_even_lane : c_odd_lane;
        return;
    }
    // single accumulator: POST_ZZ_DELTA, POST_DELTA_MORTON
    int32_t carry = cx;
    for (uint32_t r = 0; r < n; r += 32) {
        const uint32_t i = r + lane;
        const uint32_t raw = i < n ? stage[i] : 0u;
        int32_t v = POST == POST_ZZ_DELTA ? zigzag_decode32(raw) : (int32_t)raw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t t = __shfl_up_sync(FULL, v, d);
            if (lane >= (unsigned)d) v += t;
        }
        v += carry;
        if (i < n) {
            if (POST == POST_DELTA_MORTON) reinterpret_cast<int2*>(dst)[index0 + i] = morton_decode(v, num_bits, no_shift);
            else if (WIDEN) reinterpret_cast<int64_t*>(dst)[index0 + i] = (int64_t)v;
            else reinterpret_cast<int32_t*>(dst)[index0 + i] = v;
        }
        carry = __shfl_sync(FULL, v, 31);
    }
    cx = carry;
}

// runtime dispatch on the (warp-uniform) post kind
__device__ __forceinline__ void lean_rows_dispatch(int post, bool widen, const uint32_t* stage, uint32_t n, void* dst, uint64_t index0,
                                                   int32_t& cx, int32_t& cy, uint32_t num_bits, bool no_shift)
{
    switch (post) {
    case POST_PLAIN:
        if (widen) lean_rows<POST_PLAIN, true>(stage, n, dst, index0, cx, cy, num_bits, no_shift);
        else lean_rows<POST_PLAIN, false>(stage, n, dst, index0, cx, cy, num_bits, no_shift);
        break;
    case POST_ZZ: lean_rows<POST_ZZ, false>(stage, n, dst, index0, cx, cy, num_bits, no_shift); break;
    case POST_ZZ_DELTA:
        if (widen) lean_rows<POST_ZZ_DELTA, true>(stage, n, dst, index0, cx, cy, num_bits, no_shift);
        else lean_rows<POST_ZZ_DELTA, false>(stage, n, dst, index0, cx, cy, num_bits, no_shift);
        break;
    case POST_ZZ_DELTA_XY: lean_rows<POST_ZZ_DELTA_XY, false>(stage, n, dst, index0, cx, cy, num_bits, no_shift); break;
    default: lean_rows<POST_DELTA_MORTON, false>(stage, n, dst, index0, cx, cy, num_bits, no_shift); break;
    }
}

}  // namespace covt
