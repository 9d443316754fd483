// covt_assemble.cuh — geometry-column assembler (one warp per layer).
//
// Replaces CovtParser.convertGeometryColumn + getLineString / getICELineString / getLinearRing
// (J/decoder/CovtParser.java:135-274, 513-550) with GeoArrow-style buffers instead of JTS objects:
//   a_geom[F+1] -> parts, a_part[P+1] -> rings, a_ring[R+1] -> vertices, a_coords[2V'] (x,y).
// Semantics = what the encoder wrote (CovtConverter.java:580-639,689-758): topology
// streams hold COUNTS; which streams a feature consumes depends on its type. The sequential cursor
// walk of the reference becomes three flat passes of warp scans: features -> parts, parts -> rings,
// rings -> vertices (see warp_assemble).
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

constexpr int ASM_SMEM_WORDS = 3 * 33;

// last index i in [0,32) with arr[i] <= key (arr nondecreasing, arr[0] == 0)
__device__ __forceinline__ uint32_t search32(const uint32_t* arr, uint32_t key)
{
    uint32_t lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
        if (arr[lo + step] <= key) lo += step;
    return lo;
}

// exclusive warp scan of small counts: when every lane holds 0 or 1 (the common case: one part per feature, one ring per
// part) a ballot + popc replaces the 5-step shuffle scan
__device__ __forceinline__ uint32_t warp_exclusive_scan_small(uint32_t v, uint32_t& total)
{
    const unsigned nz = __ballot_sync(FULL, v != 0u), big = __ballot_sync(FULL, v > 1u);
    if (!big) {
        total = (uint32_t)__popc(nz);
        return (uint32_t)__popc(nz & ((1u << lane_id()) - 1u));
    }
    return warp_exclusive_scan(v, total);
}

// The reference walks features with three running cursors (CovtParser.java:135-274). Here the walk is three flat passes, each
// 32 items per trip with one or two warp scans — features -> parts, parts -> rings, rings -> vertices — instead of a nested
// cascade that paid ~550 warp instructions per 32 features however simple they were (ncu: 2420 per layer of 49 features):
//   level 1 writes a_geom and, as scratch in a_part[p + 1], the geometry type of every part's feature;
//   level 2 reads that, writes the final a_part and, as scratch in a_ring[r + 1], the vertex count of every non-polygon ring
//           (>= 0) or -1 for a polygon ring (its count comes from the ring stream);
//   level 3 reads that, writes the final a_ring and moves the vertices (gathering through vertex_offsets for ICE layers).
__device__ __forceinline__ void warp_assemble(const LayerIO& io, uint32_t* sm, AsmResult& res)
{
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t* r_start = sm;        // [33] first output vertex of the ring, relative to the block of 32 rings
    uint32_t* r_src = sm + 33;     // [32] first source vertex of the ring
    uint32_t* r_n = sm + 66;       // [32] source vertex count of the ring

    const bool ice = io.voff != nullptr;
    const uint64_t src_total = ice ? io.n_voff : io.vbuf_ints / 2;
    const uint64_t dict = io.vbuf_ints / 2;
    uint32_t p = 0, r = 0;  // assembled parts / rings
    uint64_t v = 0, s = 0;  // assembled (output) vertices / consumed source vertices
    uint32_t status = COVT_OK;
    if (lane == 0) { io.a_geom[0] = 0; io.a_part[0] = 0; io.a_ring[0] = 0; }

#define ASM_CHECK(cond, code)                                  \
    if (__any_sync(FULL, (cond))) { status = (code); goto done; }

    // ---- level 1: features -> parts ------------------------------------------------------------
    {
        uint32_t gc = 0;  // geometry_offsets cursor
        for (uint32_t f0 = 0; f0 < io.F; f0 += 32) {
            const uint32_t f = f0 + lane;
            const bool fvalid = f < io.F;
            const uint32_t t = fvalid ? (uint32_t)__ldg(io.types + f) : (uint32_t)COVT_GT_POINT;
            ASM_CHECK(fvalid && (t == COVT_GT_MULTIPOINT || t > COVT_GT_MULTIPOLYGON), COVT_ERR_UNSUPPORTED_GEOMETRY);
            const bool uses_g = fvalid && (t == COVT_GT_MULTILINESTRING || t == COVT_GT_MULTIPOLYGON);
            const unsigned gmask = __ballot_sync(FULL, uses_g);
            const uint32_t gidx = gc + (uint32_t)__popc(gmask & lt);
            ASM_CHECK(uses_g && gidx >= io.n_geom, COVT_ERR_TOPOLOGY);
            int32_t nparts_s = fvalid ? 1 : 0;
            if (uses_g) nparts_s = __ldg(io.geom + gidx);
            ASM_CHECK(nparts_s < 0 || (uint32_t)nparts_s > io.cap_parts, COVT_ERR_TOPOLOGY);
            const uint32_t nparts = (uint32_t)nparts_s;
            uint32_t tot;
            const uint32_t excl = warp_exclusive_scan_small(nparts, tot);
            ASM_CHECK((uint64_t)p + tot > io.cap_parts, COVT_ERR_TOPOLOGY);
            if (fvalid) io.a_geom[f + 1] = (int32_t)(p + excl + nparts);
            for (uint32_t k = 0; k < nparts; k++) io.a_part[p + excl + k + 1] = (int32_t)t;  // scratch: the part's geometry type
            p += tot;
            gc += (uint32_t)__popc(gmask);
        }
    }
    __syncwarp();
    // ---- level 2: parts -> rings ----------------------------------------------------------------
    {
        uint32_t pc = 0;  // part_offsets cursor
        for (uint32_t k0 = 0; k0 < p; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool pvalid = k < p;
            const uint32_t tt = pvalid ? (uint32_t)io.a_part[k + 1] : (uint32_t)COVT_GT_POINT;
            const bool uses_p = pvalid && tt != COVT_GT_POINT;
            const unsigned pmask = __ballot_sync(FULL, uses_p);
            const uint32_t pe = pc + (uint32_t)__popc(pmask & lt);
            ASM_CHECK(uses_p && pe >= io.n_part, COVT_ERR_TOPOLOGY);
            const int32_t cnt = uses_p ? __ldg(io.part + pe) : 1;
            const bool poly = pvalid && (tt == COVT_GT_POLYGON || tt == COVT_GT_MULTIPOLYGON);
            ASM_CHECK(pvalid && (cnt < 0 || (poly && (uint32_t)cnt > io.cap_rings)), COVT_ERR_TOPOLOGY);
            const uint32_t nrings = pvalid ? (poly ? (uint32_t)cnt : 1u) : 0u;
            uint32_t tot;
            const uint32_t excl = warp_exclusive_scan_small(nrings, tot);
            ASM_CHECK((uint64_t)r + tot > io.cap_rings, COVT_ERR_TOPOLOGY);
            // scratch: vertex count of a line / point ring, -1 for a polygon ring
            for (uint32_t j = 0; j < nrings; j++) io.a_ring[r + excl + j + 1] = poly ? -1 : cnt;
            if (pvalid) io.a_part[k + 1] = (int32_t)(r + excl + nrings);
            r += tot;
            pc += (uint32_t)__popc(pmask);
        }
    }
    __syncwarp();
    // ---- level 3: rings -> vertices ---------------------------------------------------------------
    {
        uint32_t rc = 0;  // ring_offsets cursor
        for (uint32_t q0 = 0; q0 < r; q0 += 32) {
            const uint32_t q = q0 + lane;
            const bool rvalid = q < r;
            const int32_t info = rvalid ? io.a_ring[q + 1] : 0;
            const bool is_poly = rvalid && info < 0;
            const unsigned rmask = __ballot_sync(FULL, is_poly);
            const uint32_t re = rc + (uint32_t)__popc(rmask & lt);
            ASM_CHECK(is_poly && re >= io.n_ring, COVT_ERR_TOPOLOGY);
            const int32_t nv_s = is_poly ? __ldg(io.ring + re) : info;
            ASM_CHECK(rvalid && nv_s < 0, COVT_ERR_TOPOLOGY);
            const uint32_t nv = rvalid ? (uint32_t)nv_s : 0u;
            const uint32_t outn = nv + ((is_poly && io.close_rings && nv > 0) ? 1u : 0u);
            uint64_t tot_sv, tot_ov, sv_excl, ov_excl;
            if (__any_sync(FULL, nv >= (1u << 26))) {
                sv_excl = warp_exclusive_scan_u64(nv, tot_sv);
                ov_excl = warp_exclusive_scan_u64(outn, tot_ov);
            } else {  // 32 counts below 2^26 cannot overflow 32 bits; closing vertices: one per closed ring
                uint32_t t1;
                const unsigned closed = __ballot_sync(FULL, outn != nv);
                sv_excl = warp_exclusive_scan(nv, t1);
                ov_excl = sv_excl + (uint32_t)__popc(closed & lt);
                tot_sv = t1;
                tot_ov = (uint64_t)t1 + (uint32_t)__popc(closed);
            }
            ASM_CHECK(s + tot_sv > src_total || v + tot_ov > io.cap_coords || tot_ov > 0xffffffffull, COVT_ERR_TOPOLOGY);
            if (rvalid) io.a_ring[q + 1] = (int32_t)(v + ov_excl + outn);
            __syncwarp();
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
                        si[b4] = (uint64_t)r_src[ri] + (i == r_n[ri] ? 0u : i);  // i == n: the closing vertex = vertex 0 of the ring
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
            rc += (uint32_t)__popc(rmask);
        }
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
