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
//   * zigzag + delta run on four consecutive values per lane (one LDS.128, one warp scan per 128 values, one STG.128): the
//     values cross shared memory once and leave with 16-byte coalesced stores (lean_rows4).
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
    uint32_t so = (uint32_t)__cvta_generic_to_shared(sp);  // 32-bit shared address: one predicated add per terminator
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t p = (w[j >> 2] >> (8 * (j & 3))) & 0x7fu;
        acc = p * mul + acc;
        ov |= mul;
        if ((cm & cont_bit_of_byte(j)) == 0u) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(so), "r"(acc) : "memory");
            so += 4;
            acc = 0;
            mul = 1;
        } else {
            mul <<= 7;
        }
    }
    ov |= mul;
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

// The aggregate pass without per-byte values: zigzag decode is linear in the payload bytes once the sign is known —
//   zz(v) = s * (v >> 1) - b,  b = bit 0 of the value's FIRST byte, s = 1 - 2b,  v >> 1 = (p0 >> 1) + sum_{k>=1} p_k << (7k - 1)
// — so every byte adds its own term q * m to the running sum of its value's class (first byte: q = p0 >> 1, m = s, and -b;
// byte k >= 1: q = p_k, m = s << (7k - 1)), all exact mod 2^32 like the Java int sums. No accumulator, no per-byte zigzag: the
// multiply-adds run on the FMA pipe, which the byte loops otherwise leave half idle (the ALU pipe is what bounds them).
// A value that started in an earlier lane continues with the multiplier derived from the halo (its earlier bytes were added by the
// lanes that hold them). ZZ = false: plain sums (m = 1 << 7k).
struct LinCarry { uint32_t m; bool first; };  // multiplier of the lane's first byte / "the lane's first byte starts a value"
template <bool ZZ>
__device__ __forceinline__ LinCarry lean_lin_carry_in(uint32_t halo)
{
    const uint32_t hterm = ~halo & 0x80808080u;
    const uint32_t k = hterm ? (uint32_t)(__clz(hterm) >> 3) : 4u;  // trailing continuation bytes before the lane
    LinCarry c;
    c.first = k == 0u;
    if (ZZ) {
        const uint32_t b = (halo >> ((32u - 8u * k) & 31u)) & 1u;   // bit 0 of the value's first byte (k >= 1)
        c.m = (1u - 2u * b) << ((7u * k - 1u) & 31u);
    } else {
        c.m = 1u << (7u * k);
    }
    return c;
}
template <bool ZZ>
__device__ __forceinline__ void lean_sum_lane_lin(const uint32_t w[4], uint32_t cm, LinCarry c, int32_t& cur, int32_t& oth)
{
    uint32_t wm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) wm[q] = w[q] & 0x7f7f7f7fu;
    uint32_t sm = c.m;
    bool first = c.first;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t p = __byte_perm(wm[j >> 2], 0u, 0x4440u + (j & 3));
        const bool term = (cm & cont_bit_of_byte(j)) == 0u;
        uint32_t q = p, m = sm;
        if (ZZ) {
            if (first) {
                const uint32_t b = p & 1u;
                q = p >> 1;
                m = 1u - 2u * b;
                cur -= (int32_t)b;
            }
            cur += (int32_t)(q * m);
            sm = first ? m << 6 : m << 7;
        } else {
            if (first) m = 1u;
            cur += (int32_t)(q * m);
            sm = m << 7;
        }
        if (term) { const int32_t t = cur; cur = oth; oth = t; }
        first = term;
    }
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

// VariableByte tail of FastPFOR (bytes already flipped to LEB128 convention): a value may span 5 bytes. JavaFastPFOR does not
// cap it — a 6th byte is added at shift 35 mod 32 = 3 — so a stream with FIVE consecutive non-final bytes decodes to garbage in
// the reference; the decoder (and the oracle) flag it as COVT_ERR_VARINT_OVERLONG instead of imitating the wrap-around.
// carry = continuation bits of the 5 bytes before the window (in: previous window, out: this one).
__device__ __forceinline__ bool vb_run_of_five(const uint32_t w[4], uint32_t& carry)
{
    uint32_t m = 0;  // bit j = continuation bit of byte j of the lane's 16 bytes
#pragma unroll
    for (int k = 0; k < 4; k++) m |= (((((w[k] & 0x80808080u) >> 7) * 0x00204081u) >> 21) & 0xfu) << (4 * k);
    uint32_t prev = __shfl_up_sync(FULL, m >> 11, 1);
    if (lane_id() == 0) prev = carry;
    carry = __shfl_sync(FULL, m >> 11, 31);
    const uint32_t M = (m << 5) | (prev & 31u);
    return (M & (M >> 1) & (M >> 2) & (M >> 3) & (M >> 4)) != 0u;
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

// ---- zigzag / delta / store, four values per lane ---------------------------------------------------------------
// A: 16-byte aligned shared array. A[s4 + i], i < n, are the raw values of one chunk in stream order, where
// s4 = index0 & 3 (index0 = stream index of the chunk's first value) and A[0 .. s4) are zero: A[j] then belongs to stream index
// base + j with base = index0 - s4 a multiple of 4, so that a lane's four consecutive values are (x, y, x, y), its 16 output
// bytes are 16-byte aligned in dst, and one LDS.128 / STG.128 moves them. One warp scan per 128 values.
// (cx, cy) = running sums of the values before the chunk at even / odd stream positions; updated on return.
template <int POST, bool WIDEN>
__device__ __forceinline__ void lean_rows4(uint32_t* A, uint32_t s4, uint32_t n, void* dst, uint64_t index0, int32_t& cx, int32_t& cy,
                                           uint32_t num_bits, bool no_shift)
{
    const unsigned lane = lane_id();
    const uint64_t base = index0 - s4;
    const uint32_t end = s4 + n;
    // zero the tail of the last 4-vector so that no delta has to be masked below (A[0 .. s4) is zero by contract)
    if (lane < 3u && ((end + lane) >> 2) == (end >> 2) && (end & 3u)) A[end + lane] = 0;
    __syncwarp();
    for (uint32_t r = 0; r < end; r += 128) {
        const uint32_t j = r + 4u * lane;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (j < end) raw = *reinterpret_cast<const uint4*>(A + j);
        int32_t d0, d1, d2, d3;
        if (POST == POST_PLAIN || POST == POST_DELTA_MORTON) { d0 = (int32_t)raw.x; d1 = (int32_t)raw.y; d2 = (int32_t)raw.z; d3 = (int32_t)raw.w; }
        else { d0 = zigzag_decode32(raw.x); d1 = zigzag_decode32(raw.y); d2 = zigzag_decode32(raw.z); d3 = zigzag_decode32(raw.w); }
        int32_t o0, o1, o2, o3;
        if (POST == POST_ZZ_DELTA_XY) {
            const int32_t sx = d0 + d2, sy = d1 + d3;
            int32_t ix = sx, iy = sy;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t tx = __shfl_up_sync(FULL, ix, d), ty = __shfl_up_sync(FULL, iy, d);
                if (lane >= (unsigned)d) { ix += tx; iy += ty; }
            }
            o0 = cx + (ix - sx) + d0;
            o1 = cy + (iy - sy) + d1;
            o2 = o0 + d2;
            o3 = o1 + d3;
            cx += __shfl_sync(FULL, ix, 31);
            cy += __shfl_sync(FULL, iy, 31);
        } else if (POST == POST_ZZ_DELTA || POST == POST_DELTA_MORTON) {
            const int32_t sm = d0 + d1 + d2 + d3;
            int32_t is = sm;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t t = __shfl_up_sync(FULL, is, d);
                if (lane >= (unsigned)d) is += t;
            }
            o0 = cx + (is - sm) + d0;
            o1 = o0 + d1;
            o2 = o1 + d2;
            o3 = o2 + d3;
            cx += __shfl_sync(FULL, is, 31);
        } else {
            o0 = d0; o1 = d1; o2 = d2; o3 = d3;
        }
        if (j >= end) continue;
        const uint64_t e = base + j;  // stream index of element 0 of this lane (multiple of 4)
        const bool full = j >= s4 && j + 4u <= end;
        const bool v0 = j + 0 >= s4, v1 = j + 1 >= s4 && j + 1 < end, v2 = j + 2 >= s4 && j + 2 < end, v3 = j + 3 >= s4 && j + 3 < end;
        if (POST == POST_DELTA_MORTON) {
            int2* out = reinterpret_cast<int2*>(dst) + e;
            const int2 m0 = morton_decode(o0, num_bits, no_shift), m1 = morton_decode(o1, num_bits, no_shift);
            const int2 m2 = morton_decode(o2, num_bits, no_shift), m3 = morton_decode(o3, num_bits, no_shift);
            if (full) {
                *reinterpret_cast<int4*>(out) = make_int4(m0.x, m0.y, m1.x, m1.y);
                *reinterpret_cast<int4*>(out + 2) = make_int4(m2.x, m2.y, m3.x, m3.y);
            } else { if (v0) out[0] = m0; if (v1) out[1] = m1; if (v2) out[2] = m2; if (v3) out[3] = m3; }
        } else if (WIDEN) {
            int64_t* out = reinterpret_cast<int64_t*>(dst) + e;
            if (full) {
                *reinterpret_cast<longlong2*>(out) = make_longlong2((long long)o0, (long long)o1);
                *reinterpret_cast<longlong2*>(out + 2) = make_longlong2((long long)o2, (long long)o3);
            } else { if (v0) out[0] = o0; if (v1) out[1] = o1; if (v2) out[2] = o2; if (v3) out[3] = o3; }
        } else {
            int32_t* out = reinterpret_cast<int32_t*>(dst) + e;
            if (full) *reinterpret_cast<int4*>(out) = make_int4(o0, o1, o2, o3);
            else { if (v0) out[0] = o0; if (v1) out[1] = o1; if (v2) out[2] = o2; if (v3) out[3] = o3; }
        }
    }
}

// Two full rows at once: A[0 .. 256) are 256 consecutive raw values whose first stream index e0 is a multiple of 4 — no bounds, no
// partial vectors (the large-stream decode pass emits nothing else between the first and the last row of a superchunk). Eight
// values per lane (two LDS.128 with a 2-way bank conflict, which costs load-store cycles the kernel has to spare): ONE warp scan
// per 256 values instead of two, and 32 contiguous output bytes per lane.
template <int POST>
__device__ __forceinline__ void lean_row256(const uint32_t* A, void* dst, uint64_t e0, int32_t& cx, int32_t& cy, uint32_t num_bits, bool no_shift)
{
    const unsigned lane = lane_id();
    const uint4 r0 = *reinterpret_cast<const uint4*>(A + 8u * lane), r1 = *reinterpret_cast<const uint4*>(A + 8u * lane + 4u);
    int32_t d[8] = {(int32_t)r0.x, (int32_t)r0.y, (int32_t)r0.z, (int32_t)r0.w, (int32_t)r1.x, (int32_t)r1.y, (int32_t)r1.z, (int32_t)r1.w};
    if (POST != POST_PLAIN && POST != POST_DELTA_MORTON) {
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = zigzag_decode32((uint32_t)d[i]);
    }
    int32_t o[8];
    if (POST == POST_ZZ_DELTA_XY) {
        const int32_t sx = (d[0] + d[2]) + (d[4] + d[6]), sy = (d[1] + d[3]) + (d[5] + d[7]);
        int32_t ix = sx, iy = sy;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int32_t tx = __shfl_up_sync(FULL, ix, k), ty = __shfl_up_sync(FULL, iy, k);
            if (lane >= (unsigned)k) { ix += tx; iy += ty; }
        }
        o[0] = cx + (ix - sx) + d[0];
        o[1] = cy + (iy - sy) + d[1];
#pragma unroll
        for (int i = 2; i < 8; i++) o[i] = o[i - 2] + d[i];
        cx += __shfl_sync(FULL, ix, 31);
        cy += __shfl_sync(FULL, iy, 31);
    } else if (POST == POST_ZZ_DELTA || POST == POST_DELTA_MORTON) {
        const int32_t sm = ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
        int32_t is = sm;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int32_t t = __shfl_up_sync(FULL, is, k);
            if (lane >= (unsigned)k) is += t;
        }
        o[0] = cx + (is - sm) + d[0];
#pragma unroll
        for (int i = 1; i < 8; i++) o[i] = o[i - 1] + d[i];
        cx += __shfl_sync(FULL, is, 31);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = d[i];
    }
    const uint64_t e = e0 + 8u * lane;
    if (POST == POST_DELTA_MORTON) {
        int4* out = reinterpret_cast<int4*>(reinterpret_cast<int2*>(dst) + e);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int2 m0 = morton_decode(o[2 * i], num_bits, no_shift), m1 = morton_decode(o[2 * i + 1], num_bits, no_shift);
            out[i] = make_int4(m0.x, m0.y, m1.x, m1.y);
        }
    } else {
        int4* out = reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(dst) + e);
        out[0] = make_int4(o[0], o[1], o[2], o[3]);
        out[1] = make_int4(o[4], o[5], o[6], o[7]);
    }
}

// runtime dispatch on the (warp-uniform) post kind
__device__ __forceinline__ void lean_rows4_dispatch(int post, bool widen, uint32_t* A, uint32_t s4, uint32_t n, void* dst, uint64_t index0,
                                                    int32_t& cx, int32_t& cy, uint32_t num_bits, bool no_shift)
{
    switch (post) {
    case POST_PLAIN:
        if (widen) lean_rows4<POST_PLAIN, true>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        else lean_rows4<POST_PLAIN, false>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        break;
    case POST_ZZ:
        if (widen) lean_rows4<POST_ZZ, true>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        else lean_rows4<POST_ZZ, false>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        break;
    case POST_ZZ_DELTA:
        if (widen) lean_rows4<POST_ZZ_DELTA, true>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        else lean_rows4<POST_ZZ_DELTA, false>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift);
        break;
    case POST_ZZ_DELTA_XY: lean_rows4<POST_ZZ_DELTA_XY, false>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift); break;
    default: lean_rows4<POST_DELTA_MORTON, false>(A, s4, n, dst, index0, cx, cy, num_bits, no_shift); break;
    }
}

}  // namespace covt
