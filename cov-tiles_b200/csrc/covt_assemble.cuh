// covt_assemble.cuh — geometry-column assembler (one warp per layer).
//
// Replaces CovtParser.convertGeometryColumn + getLineString / getICELineString / getLinearRing
// (J/decoder/CovtParser.java:135-274, 513-550) with GeoArrow-style buffers instead of JTS objects:
//   a_geom[F+1] -> parts, a_part[P+1] -> rings, a_ring[R+1] -> vertices, a_coords[2V'] (x,y).
// Semantics = what the encoder wrote (CovtConverter.java:580-639,689-758): topology
// streams hold COUNTS; which streams a feature consumes depends on its type. The sequential cursor
// walk of the reference becomes a cascade of warp scans: features -> parts -> rings -> vertices, each
// level expanded 32 children at a time with a 5-step search over the parents' prefix in shared memory.
#pragma once
#include "covt_device.cuh"

namespace covt {

struct LayerIO {
    const uint8_t* types; uint32_t F;
    const int32_t *geom, *part, *ring, *voff, *vbuf;
    uint32_t n_geom, n_part, n_ring, n_voff;
    uint64_t vbuf_ints;
    int32_t *a_geom, *a_part, *a_ring, *a_coords;
    uint32_t cap_parts, cap_rings;
    uint64_t cap_coords;  // vertices
    bool close_rings;
};
struct AsmResult { uint32_t status, n_parts, n_rings, n_vertices, n_coords; };

constexpr int ASM_SMEM_WORDS = 10 * 33;

// last index i in [0,32) with arr[i] <= key (arr nondecreasing, arr[0] == 0)
__device__ __forceinline__ uint32_t search32(const uint32_t* arr, uint32_t key)
{
    uint32_t lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
        if (arr[lo + step] <= key) lo += step;
    return lo;
}

__device__ __forceinline__ void warp_assemble(const LayerIO& io, uint32_t* sm, AsmResult& res)
{
    const unsigned lane = lane_id();
    uint32_t* f_start = sm;            // [33] exclusive prefix of parts per feature
    uint32_t* f_type = sm + 33;        // [32]
    uint32_t* f_pe = sm + 66;          // [32] first part-stream entry of the feature
    uint32_t* p_start = sm + 99;       // [33] exclusive prefix of rings per part
    uint32_t* p_poly = sm + 132;       // [32]
    uint32_t* p_re = sm + 165;         // [32] first ring-stream entry of the part
    uint32_t* p_line_n = sm + 198;     // [32] vertex count of a non-polygon part
    uint32_t* r_start = sm + 231;      // [33] exclusive prefix of output vertices per ring
    uint32_t* r_src = sm + 264;        // [32] first source vertex of the ring
    uint32_t* r_n = sm + 297;          // [32] source vertex count of the ring

    const bool ice = io.voff != nullptr;
    const uint64_t src_total = ice ? io.n_voff : io.vbuf_ints / 2;
    const uint64_t dict = io.vbuf_ints / 2;
    uint32_t gc = 0, pc = 0, rc = 0;  // stream cursors
    uint32_t p = 0, r = 0;            // assembled parts / rings
    uint64_t v = 0, s = 0;            // assembled (output) vertices / consumed source vertices
    uint32_t status = COVT_OK;
    if (lane == 0) { io.a_geom[0] = 0; io.a_part[0] = 0; io.a_ring[0] = 0; }

#define ASM_CHECK(cond, code)                                  \
    if (__any_sync(FULL, (cond))) { status = (code); goto done; }

    for (uint32_t f0 = 0; f0 < io.F; f0 += 32) {
        const uint32_t f = f0 + lane;
        const bool fvalid = f < io.F;
        const uint32_t t = fvalid ? __ldg(io.types + f) : 0xffu;
        ASM_CHECK(fvalid && (t == COVT_GT_MULTIPOINT || t > COVT_GT_MULTIPOLYGON), COVT_ERR_UNSUPPORTED_GEOMETRY);
        const bool uses_g = fvalid && (t == COVT_GT_MULTILINESTRING || t == COVT_GT_MULTIPOLYGON);
        uint32_t tot_g;
        const uint32_t gidx = gc + warp_exclusive_scan(uses_g ? 1u : 0u, tot_g);
        ASM_CHECK(uses_g && gidx >= io.n_geom, COVT_ERR_TOPOLOGY);
        int32_t nparts_s = fvalid ? 1 : 0;
        if (uses_g) nparts_s = __ldg(io.geom + gidx);
        ASM_CHECK(nparts_s < 0, COVT_ERR_TOPOLOGY);
        const uint32_t nparts = (uint32_t)nparts_s;
        const uint32_t part_entries = (fvalid && t != COVT_GT_POINT) ? nparts : 0u;
        uint32_t tot_np, tot_pe;
        const uint32_t np_excl = warp_exclusive_scan(nparts, tot_np);
        const uint32_t pe_excl = warp_exclusive_scan(part_entries, tot_pe);
        ASM_CHECK((uint64_t)p + tot_np > io.cap_parts || (uint64_t)pc + tot_pe > io.n_part, COVT_ERR_TOPOLOGY);
        if (fvalid) io.a_geom[f + 1] = (int32_t)(p + np_excl + nparts);
        f_start[lane] = np_excl;
        f_type[lane] = t;
        f_pe[lane] = pc + pe_excl;
        __syncwarp();
        for (uint32_t k0 = 0; k0 < tot_np; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool pvalid = k < tot_np;
            uint32_t nrings = 0, ring_entries = 0, line_n = 0;
            bool poly = false;
            if (pvalid) {
                const uint32_t fi = search32(f_start, k);
                const uint32_t tt = f_type[fi];
                poly = (tt == COVT_GT_POLYGON || tt == COVT_GT_MULTIPOLYGON);
                int32_t cnt = 1;
                if (tt != COVT_GT_POINT) cnt = __ldg(io.part + f_pe[fi] + (k - f_start[fi]));
                if (cnt < 0) nrings = 0xffffffffu;  // flagged below
                else if (poly) { nrings = (uint32_t)cnt; ring_entries = (uint32_t)cnt; }
                else { nrings = 1; line_n = (uint32_t)cnt; }
            }
            ASM_CHECK(nrings == 0xffffffffu, COVT_ERR_TOPOLOGY);
            uint32_t tot_nr, tot_re;
            const uint32_t nr_excl = warp_exclusive_scan(nrings, tot_nr);
            const uint32_t re_excl = warp_exclusive_scan(ring_entries, tot_re);
            ASM_CHECK((uint64_t)r + tot_nr > io.cap_rings || (uint64_t)rc + tot_re > io.n_ring, COVT_ERR_TOPOLOGY);
            if (pvalid) io.a_part[p + k + 1] = (int32_t)(r + nr_excl + nrings);
            p_start[lane] = nr_excl;
            p_poly[lane] = poly ? 1u : 0u;
            p_re[lane] = rc + re_excl;
            p_line_n[lane] = line_n;
            __syncwarp();
            for (uint32_t q0 = 0; q0 < tot_nr; q0 += 32) {
                const uint32_t q = q0 + lane;
                const bool rvalid = q < tot_nr;
                uint32_t nv = 0, outn = 0;
                bool bad = false;
                if (rvalid) {
                    const uint32_t pi = search32(p_start, q);
                    if (p_poly[pi]) {
                        const int32_t c = __ldg(io.ring + p_re[pi] + (q - p_start[pi]));
                        if (c < 0) bad = true;
                        else { nv = (uint32_t)c; outn = nv + ((io.close_rings && nv > 0) ? 1u : 0u); }
                    } else {
                        nv = p_line_n[pi];
                        outn = nv;
                    }
                }
                ASM_CHECK(bad, COVT_ERR_TOPOLOGY);
                uint64_t tot_sv, tot_ov;
                const uint64_t sv_excl = warp_exclusive_scan_u64(nv, tot_sv);
                const uint64_t ov_excl = warp_exclusive_scan_u64(outn, tot_ov);
                ASM_CHECK(s + tot_sv > src_total || v + tot_ov > io.cap_coords || tot_ov > 0xffffffffull, COVT_ERR_TOPOLOGY);
                if (rvalid) io.a_ring[r + q + 1] = (int32_t)(v + ov_excl + outn);
                r_start[lane] = (uint32_t)ov_excl;
                r_src[lane] = (uint32_t)(s + sv_excl);
                r_n[lane] = nv;
                __syncwarp();
                const uint32_t n_out = (uint32_t)tot_ov;
                bool oob = false;
                // 4 batches of 32 output vertices per trip: the index loads, then the coordinate gathers, are issued
                // back to back so that their latencies overlap (the decoded streams were written by earlier kernels,
                // so the read-only path is safe here)
                for (uint32_t u0 = 0; u0 < n_out; u0 += 128) {
                    uint64_t si[4];
                    bool ok[4];
#pragma unroll
                    for (int b4 = 0; b4 < 4; b4++) {
                        const uint32_t u = u0 + 32u * b4 + lane;
                        ok[b4] = u < n_out;
                        si[b4] = 0;
                        if (ok[b4]) {
                            const uint32_t ri = search32(r_start, u);
                            const uint32_t i = u - r_start[ri];
                            si[b4] = (uint64_t)r_src[ri] + (i == r_n[ri] ? 0u : i);
                        }
                    }
                    if (ice) {
                        int32_t o[4];
#pragma unroll
                        for (int b4 = 0; b4 < 4; b4++) o[b4] = ok[b4] ? __ldg(io.voff + si[b4]) : 0;
#pragma unroll
                        for (int b4 = 0; b4 < 4; b4++) {
                            if (ok[b4] && (o[b4] < 0 || (uint64_t)o[b4] >= dict)) { oob = true; ok[b4] = false; }
                            si[b4] = (uint64_t)(uint32_t)o[b4];
                        }
                    }
                    int2 xy[4];
#pragma unroll
                    for (int b4 = 0; b4 < 4; b4++) xy[b4] = ok[b4] ? __ldg(reinterpret_cast<const int2*>(io.vbuf) + si[b4]) : make_int2(0, 0);
#pragma unroll
                    for (int b4 = 0; b4 < 4; b4++)
                        if (ok[b4]) reinterpret_cast<int2*>(io.a_coords)[v + u0 + 32u * b4 + lane] = xy[b4];
                }
                ASM_CHECK(oob, COVT_ERR_TOPOLOGY);
                v += tot_ov;
                s += tot_sv;
                __syncwarp();
            }
            r += tot_nr;
            rc += tot_re;
            __syncwarp();
        }
        p += tot_np;
        pc += tot_pe;
        gc += tot_g;
        __syncwarp();
    }
done:
#undef ASM_CHECK
    res.status = status;
    res.n_parts = p;
    res.n_rings = r;
    res.n_vertices = (uint32_t)s;
    res.n_coords = (uint32_t)v;
}

}  // namespace covt
