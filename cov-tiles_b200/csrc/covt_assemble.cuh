// covt_assemble.cuh — geometry-column assembler (one warp per layer).
//
// Replaces CovtParser.convertGeometryColumn + getLineString / getICELineString / getLinearRing
// (J/decoder/CovtParser.java:135-274, 513-550) with GeoArrow-style buffers instead of JTS objects:
//   a_geom[F+1] -> parts, a_part[P+1] -> rings, a_ring[R+1] -> vertices, a_coords[2V'] (x,y).
// Semantics = what the encoder wrote (CovtConverter.java:580-639,689-758): topology
// streams hold COUNTS; which streams a feature consumes depends on its type. The sequential cursor
// walk of the reference becomes three flat passes of warp scans: features -> parts, parts -> rings,
// rings -> vertices, 128 items per trip (see warp_assemble).
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

constexpr int ASM_IPL = 4;                          // items (features / parts / rings) per lane and trip
constexpr int ASM_TRIP = 32 * ASM_IPL;              // items per trip
constexpr int ASM_WINDOW = 1024;                    // output vertices per ring-mark window (32 mask words)
constexpr int ASM_SMEM_WORDS = 3 * ASM_TRIP + 32;   // per warp: three ring tables + the mask words

// The reference walks features with three running cursors (CovtParser.java:135-274). Here the walk is three flat passes —
// features -> parts, parts -> rings, rings -> vertices — each covering ASM_TRIP = 128 items per trip, FOUR CONSECUTIVE items per
// lane: the four count loads of a lane are independent (memory-level parallelism), a lane-local prefix plus ONE warp scan per
// quantity replaces per-32 scans, and a layer of F features needs F/128 dependent trips per level instead of F/32 (the
// assembler waits on its dependent loads, not on issue slots).
//   level 1 writes a_geom and, as scratch in a_part[p + 1], the geometry type of every part's feature;
//   level 2 reads that, writes the final a_part and, as scratch in a_ring[r + 1], the vertex count of every non-polygon ring
//           (>= 0) or -1 for a polygon ring (its count comes from the ring stream);
//   level 3 reads that, writes the final a_ring and moves the vertices (gathering through vertex_offsets for ICE layers).
__device__ __forceinline__ void warp_assemble(const LayerIO& io, uint32_t* sm, AsmResult& res)
{
    const unsigned lane = lane_id();
    uint32_t* r_start = sm;                // [ASM_TRIP] first output vertex of the k-th vertex-owning ring, relative to the trip
    uint32_t* r_src = sm + ASM_TRIP;       // [ASM_TRIP] its first source vertex
    uint32_t* r_n = sm + 2 * ASM_TRIP;     // [ASM_TRIP] its source vertex count
    uint32_t* marks = sm + 3 * ASM_TRIP;   // [32] ring-start bits of the current window of output vertices

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
        for (uint32_t f0 = 0; f0 < io.F; f0 += ASM_TRIP) {
            const uint32_t fb = f0 + ASM_IPL * lane;
            // four types in one word (the slice is 16-byte aligned and padded)
            const uint32_t tw = fb < io.F ? __ldg(reinterpret_cast<const uint32_t*>(io.types + fb)) : 0u;
            uint32_t t[ASM_IPL], nparts[ASM_IPL];
            bool valid[ASM_IPL], uses_g[ASM_IPL];
            bool bad_type = false;
            uint32_t n_g = 0;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                valid[i] = fb + i < io.F;
                t[i] = (tw >> (8 * i)) & 0xffu;
                bad_type = bad_type || (valid[i] && (t[i] == COVT_GT_MULTIPOINT || t[i] > COVT_GT_MULTIPOLYGON));
                uses_g[i] = valid[i] && (t[i] == COVT_GT_MULTILINESTRING || t[i] == COVT_GT_MULTIPOLYGON);
                n_g += uses_g[i] ? 1u : 0u;
            }
            ASM_CHECK(bad_type, COVT_ERR_UNSUPPORTED_GEOMETRY);
            uint32_t tot_g = 0, gidx = gc;
            bool err = false;
            if (__any_sync(FULL, n_g != 0u)) {  // multi-geometries in this trip: their part counts come from geometry_offsets
                gidx = gc + warp_exclusive_scan(n_g, tot_g);
            }
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                int32_t np = valid[i] ? 1 : 0;
                if (uses_g[i]) {
                    if (gidx >= io.n_geom) err = true;
                    else np = __ldg(io.geom + gidx);
                    gidx++;
                }
                if (np < 0 || (uint32_t)np > io.cap_parts) { err = true; np = 0; }
                nparts[i] = (uint32_t)np;
            }
            ASM_CHECK(err, COVT_ERR_TOPOLOGY);
            const uint32_t lane_sum = nparts[0] + nparts[1] + nparts[2] + nparts[3];
            uint32_t tot, excl;
            if (tot_g == 0) {  // one part per feature: the prefix is the feature index
                excl = min(fb, io.F) - f0;
                tot = min((uint32_t)ASM_TRIP, io.F - f0);
            } else if (__any_sync(FULL, (nparts[0] | nparts[1] | nparts[2] | nparts[3]) >= (1u << 24))) {  // 32-bit sums could wrap
                uint64_t t64;
                const uint64_t e64 = warp_exclusive_scan_u64((uint64_t)nparts[0] + nparts[1] + nparts[2] + nparts[3], t64);
                ASM_CHECK((uint64_t)p + t64 > io.cap_parts, COVT_ERR_TOPOLOGY);
                excl = (uint32_t)e64;
                tot = (uint32_t)t64;
            } else {
                excl = warp_exclusive_scan(lane_sum, tot);
            }
            ASM_CHECK((uint64_t)p + tot > io.cap_parts, COVT_ERR_TOPOLOGY);
            uint32_t pw = p + excl;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                for (uint32_t k = 0; k < nparts[i]; k++) io.a_part[pw + k + 1] = (int32_t)t[i];  // scratch: the part's geometry type
                pw += nparts[i];
                if (valid[i]) io.a_geom[fb + i + 1] = (int32_t)pw;
            }
            p += tot;
            gc += tot_g;
        }
    }
    __syncwarp();
    // ---- level 2: parts -> rings ----------------------------------------------------------------
    {
        uint32_t pc = 0;  // part_offsets cursor
        for (uint32_t k0 = 0; k0 < p; k0 += ASM_TRIP) {
            const uint32_t kb = k0 + ASM_IPL * lane;
            uint32_t tt[ASM_IPL], nrings[ASM_IPL];
            int32_t cnt[ASM_IPL];
            bool valid[ASM_IPL], poly[ASM_IPL];
            uint32_t n_p = 0;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                valid[i] = kb + i < p;
                tt[i] = valid[i] ? (uint32_t)io.a_part[kb + i + 1] : (uint32_t)COVT_GT_POINT;
                n_p += (valid[i] && tt[i] != COVT_GT_POINT) ? 1u : 0u;
            }
            uint32_t tot_p;
            uint32_t pe = pc + warp_exclusive_scan(n_p, tot_p);
            bool err = false;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                cnt[i] = 1;
                if (valid[i] && tt[i] != COVT_GT_POINT) {
                    if (pe >= io.n_part) err = true;
                    else cnt[i] = __ldg(io.part + pe);
                    pe++;
                }
                poly[i] = valid[i] && (tt[i] == COVT_GT_POLYGON || tt[i] == COVT_GT_MULTIPOLYGON);
                if (valid[i] && (cnt[i] < 0 || (poly[i] && (uint32_t)cnt[i] > io.cap_rings))) { err = true; cnt[i] = 0; }
                nrings[i] = valid[i] ? (poly[i] ? (uint32_t)cnt[i] : 1u) : 0u;
            }
            ASM_CHECK(err, COVT_ERR_TOPOLOGY);
            const uint32_t lane_sum = nrings[0] + nrings[1] + nrings[2] + nrings[3];
            uint32_t tot, excl;
            if (__any_sync(FULL, (nrings[0] | nrings[1] | nrings[2] | nrings[3]) >= (1u << 24))) {  // 32-bit sums could wrap
                uint64_t t64;
                const uint64_t e64 = warp_exclusive_scan_u64((uint64_t)nrings[0] + nrings[1] + nrings[2] + nrings[3], t64);
                ASM_CHECK((uint64_t)r + t64 > io.cap_rings, COVT_ERR_TOPOLOGY);
                excl = (uint32_t)e64;
                tot = (uint32_t)t64;
            } else {
                excl = warp_exclusive_scan(lane_sum, tot);
            }
            ASM_CHECK((uint64_t)r + tot > io.cap_rings, COVT_ERR_TOPOLOGY);
            uint32_t rw = r + excl;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                // scratch: vertex count of a line / point ring, -1 for a polygon ring
                for (uint32_t j = 0; j < nrings[i]; j++) io.a_ring[rw + j + 1] = poly[i] ? -1 : cnt[i];
                rw += nrings[i];
                if (valid[i]) io.a_part[kb + i + 1] = (int32_t)rw;
            }
            r += tot;
            pc += tot_p;
        }
    }
    __syncwarp();
    // ---- level 3: rings -> vertices ---------------------------------------------------------------
    {
        uint32_t rc = 0;  // ring_offsets cursor
        for (uint32_t q0 = 0; q0 < r; q0 += ASM_TRIP) {
            const uint32_t qb = q0 + ASM_IPL * lane;
            int32_t info[ASM_IPL];
            uint32_t nv[ASM_IPL], outn[ASM_IPL];
            bool valid[ASM_IPL], is_poly[ASM_IPL];
            uint32_t n_r = 0;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                valid[i] = qb + i < r;
                info[i] = valid[i] ? io.a_ring[qb + i + 1] : 0;
                is_poly[i] = valid[i] && info[i] < 0;
                n_r += is_poly[i] ? 1u : 0u;
            }
            uint32_t tot_r = 0, re = rc;
            if (__any_sync(FULL, n_r != 0u)) re = rc + warp_exclusive_scan(n_r, tot_r);
            bool err = false;
            bool huge = false;
            uint32_t lane_sv = 0, lane_ov = 0;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) {
                int32_t n = info[i];
                if (is_poly[i]) {
                    n = 0;
                    if (re >= io.n_ring) err = true;
                    else n = __ldg(io.ring + re);
                    re++;
                }
                if (valid[i] && n < 0) { err = true; n = 0; }
                nv[i] = valid[i] ? (uint32_t)n : 0u;
                outn[i] = nv[i] + ((is_poly[i] && io.close_rings && nv[i] > 0) ? 1u : 0u);
                huge = huge || nv[i] >= (1u << 24);
                lane_sv += nv[i];
                lane_ov += outn[i];
            }
            ASM_CHECK(err, COVT_ERR_TOPOLOGY);
            uint64_t tot_sv, tot_ov, sv_excl, ov_excl;
            if (__any_sync(FULL, huge)) {
                sv_excl = warp_exclusive_scan_u64((uint64_t)nv[0] + nv[1] + nv[2] + nv[3], tot_sv);
                ov_excl = warp_exclusive_scan_u64((uint64_t)outn[0] + outn[1] + outn[2] + outn[3], tot_ov);
            } else {  // 128 counts below 2^24 cannot overflow 32 bits
                uint32_t t1, t2;
                sv_excl = warp_exclusive_scan(lane_sv, t1);
                ov_excl = warp_exclusive_scan(lane_ov, t2);
                tot_sv = t1;
                tot_ov = t2;
            }
            ASM_CHECK(s + tot_sv > src_total || v + tot_ov > io.cap_coords || tot_ov > 0xffffffffull, COVT_ERR_TOPOLOGY);
            // ring tables of the trip, compacted over the rings that own output vertices (an empty ring owns none)
            bool own[ASM_IPL];
            uint32_t n_own = 0;
#pragma unroll
            for (int i = 0; i < ASM_IPL; i++) { own[i] = outn[i] != 0u; n_own += own[i] ? 1u : 0u; }
            uint32_t tot_own;
            uint32_t kw = warp_exclusive_scan(n_own, tot_own);
            uint32_t first_out[ASM_IPL];  // first output vertex of the lane's rings, relative to the trip
            __syncwarp();
            {
                uint64_t so = s + sv_excl, oo = ov_excl;
#pragma unroll
                for (int i = 0; i < ASM_IPL; i++) {
                    first_out[i] = (uint32_t)oo;
                    if (own[i]) {
                        r_start[kw] = (uint32_t)oo;
                        r_src[kw] = (uint32_t)so;
                        r_n[kw] = nv[i];
                        kw++;
                    }
                    so += nv[i];
                    oo += outn[i];
                    if (valid[i]) io.a_ring[qb + i + 1] = (int32_t)(v + oo);
                }
            }
            const uint32_t n_out = (uint32_t)tot_ov;
            bool oob = false;
            // The ring of an output vertex: windows of ASM_WINDOW = 1024 output vertices. Every owning ring that starts inside
            // the window sets one bit of a 32-word mask (its positions are distinct); the ring of vertex u is then
            //   (#starts before the window) + (#bits at or below u) - 1
            // = one popc and two shuffles per vertex instead of a binary search over the trip's 128 rings.
            uint32_t k_before = 0;
            for (uint32_t w0 = 0; w0 < n_out; w0 += ASM_WINDOW) {
                marks[lane] = 0u;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < ASM_IPL; i++) {
                    const uint32_t d = first_out[i] - w0;
                    if (own[i] && d < (uint32_t)ASM_WINDOW) atomicOr(&marks[d >> 5], 1u << (d & 31u));
                }
                __syncwarp();
                const uint32_t my_mask = marks[lane];
                uint32_t tot_starts;
                const uint32_t my_before = k_before + warp_exclusive_scan((uint32_t)__popc(my_mask), tot_starts);
                const uint32_t w_end = min(n_out, w0 + (uint32_t)ASM_WINDOW);
                // 4 batches of 32 output vertices per trip of the copy loop: the index loads, then the coordinate gathers, are
                // issued back to back so that their latencies overlap (the decoded streams were written by earlier kernels, so
                // the read-only path is safe here)
                for (uint32_t u0 = w0; u0 < w_end; u0 += 128) {
                    uint64_t si[4];
                    bool ok[4];
#pragma unroll
                    for (int b4 = 0; b4 < 4; b4++) {
                        const uint32_t u = u0 + 32u * b4 + lane;
                        const int word = (int)((u0 - w0) >> 5) + b4;  // uniform; < 32 because u0 - w0 <= 896
                        const uint32_t m = __shfl_sync(FULL, my_mask, word);
                        const uint32_t kb = __shfl_sync(FULL, my_before, word);
                        ok[b4] = u < w_end;
                        si[b4] = 0;
                        if (ok[b4]) {
                            const uint32_t k = kb + (uint32_t)__popc(m & lanemask_le()) - 1u;
                            const uint32_t i = u - r_start[k];
                            si[b4] = (uint64_t)r_src[k] + (i == r_n[k] ? 0u : i);  // i == n: the closing vertex = vertex 0 of the ring
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
                k_before += tot_starts;
                __syncwarp();
            }
            ASM_CHECK(oob, COVT_ERR_TOPOLOGY);
            v += tot_ov;
            s += tot_sv;
            rc += tot_r;
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
