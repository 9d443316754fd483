// covt_kernels.cu — the sm_100a kernels of the COVT batch decoder and their launchers.
//
//   k0_scan_tiles / k0_fill_layers : device-side container walk (gen-2b, gen-3) -> covt_layer table + one dense decode-task
//                                    list per codec class (CovtParser.decodeLayerMetadata, J/decoder/CovtParser.java:574-652;
//                                    SURVEY §A.1)
//   scan_*                         : result layout = exclusive prefix sums of 16-byte-rounded slice sizes
//   k_seg_begin / k_seg_end        : running totals of a batch decoded in segments (upload of segment i+1 overlaps decode of i)
//   k_decode_class<C>              : one kernel per codec class (Byte-RLE, RLE, 32-bit varint, 64-bit varint, FastPFOR):
//                                    the static codecs of DecodingUtils.java, a thread or a warp per stream
//   k_assemble_layers              : one warp per layer (CovtParser.convertGeometryColumn :135-274)
//   k1a_aggregate / k1b_decode     : large varint streams of the stream API in 512-byte chunks over the whole GPU, with a
//                                    segmented scan over (count, sumEven, sumOdd) in between (DecodingUtils.java:55-112,394-409)
//   k_tile_status / k_alg_bytes    : per-tile status; algorithmic bytes per kernel (profiling only)
#include "covt_assemble.cuh"
#include "covt_internal.h"
#include "covt_streams.cuh"
#include "covt_walk.cuh"
#include "covt_props.cuh"

namespace covt {

// =================================================================================================
// result layout (DESIGN.md "result layout"): slice sizes of one layer in every result buffer
// =================================================================================================
// word index (uint32) of covt_layer[l].streams[s].status in the layer table = l * LAYER_WORDS + STREAM0_STATUS_WORD + s * STREAM_REF_WORDS
constexpr uint32_t LAYER_WORDS = sizeof(covt_layer) / 4, STREAM_REF_WORDS = sizeof(covt_stream_ref) / 4;
constexpr uint32_t STREAM0_STATUS_WORD = (offsetof(covt_layer, streams) + offsetof(covt_stream_ref, status)) / 4;
__host__ __device__ inline uint32_t kBufElemSizeDev(int b)
{
    return (b == COVT_BUF_S_GEOMETRY_TYPES || b == COVT_BUF_STREAM_ARENA) ? 1u : (b == COVT_BUF_S_IDS ? 8u : 4u);
}
__host__ __device__ inline uint64_t align_elems(uint64_t n, uint32_t elem_size)
{
    uint64_t per = 16 / elem_size;
    return (n + per - 1) / per * per;
}

__device__ __forceinline__ uint64_t slot_nv(const covt_layer& L, int slot)
{
    return L.streams[slot].encoding == COVT_ENC_ABSENT ? 0ull : (uint64_t)L.streams[slot].num_values;
}
__device__ __forceinline__ uint64_t vbuf_ints_of(const covt_layer& L, uint32_t flags)
{
    uint64_t n = slot_nv(L, COVT_SLOT_VBUF);
    if (L.geom_column_type == COVT_CT_ICE_MORTON_CODE) n *= 2;
    else if (L.geom_column_type == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) n *= 2;
    return n;
}

// dispatch table of CovtParser.decodeGeometryColumn (:405-510) and decodedIds (:552-572), SURVEY §8a
__host__ __device__ inline uint32_t resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags)
{
    switch (stream_type) {
    case COVT_ST_GEOMETRY_TYPES: return COVT_OP_BYTE_RLE;  // always Byte-RLE whatever the label (:405-406)
    case COVT_ST_GEOMETRY_OFFSETS: case COVT_ST_PART_OFFSETS: case COVT_ST_RING_OFFSETS:
        if (encoding == COVT_ENC_RLE) return COVT_OP_RLE_U32;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA;
        return COVT_OP_NONE;
    case COVT_ST_VERTEX_OFFSETS: case COVT_ST_INDEX_BUFFER:
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_ZZ_DELTA;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA;
        return COVT_OP_NONE;
    case COVT_ST_VERTEX_BUFFER:
        if (column_type == COVT_CT_ICE_MORTON_CODE) {
            if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_DELTA_MORTON;
            if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_DELTA_MORTON;
            return COVT_OP_NONE;
        }
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) return COVT_OP_VARINT_ZZ_DELTA_XY;
        if (encoding == COVT_ENC_FAST_PFOR_DELTA_ZIG_ZAG) return COVT_OP_PFOR_ZZ_DELTA_XY;
        return COVT_OP_NONE;
    case COVT_ST_DATA:
        if (encoding == COVT_ENC_RLE) return COVT_OP_RLE_U64;
        if (encoding == COVT_ENC_VARINT) return (flags & COVT_FLAG_ID_WIDTH_32) ? COVT_OP_VARINT_U32_AS_I64 : COVT_OP_VARINT_U64;
        if (encoding == COVT_ENC_VARINT_DELTA_ZIG_ZAG) {
            if (flags & COVT_FLAG_ID_DVZZ_IS_RLE) return COVT_OP_RLE_U64;
            return (flags & COVT_FLAG_ID_WIDTH_32) ? COVT_OP_VARINT_ZZ_DELTA_AS_I64 : COVT_OP_VARINT_ZZ_DELTA_64;
        }
        return COVT_OP_NONE;
    default: return COVT_OP_NONE;
    }
}
uint32_t host_resolve_op(uint32_t st, uint32_t enc, uint32_t ct, uint32_t flags) { return resolve_op(st, enc, ct, flags); }

__host__ __device__ inline int op_class_of(uint32_t op)
{
    switch (op) {
    case COVT_OP_BYTE_RLE: return CLASS_BYTE_RLE;
    case COVT_OP_RLE_U32: case COVT_OP_RLE_U64: case COVT_OP_RLE_S64: return CLASS_RLE;
    case COVT_OP_VARINT_U32: case COVT_OP_VARINT_ZZ: case COVT_OP_VARINT_ZZ_DELTA: case COVT_OP_VARINT_ZZ_DELTA_XY:
    case COVT_OP_VARINT_DELTA_MORTON: case COVT_OP_VARINT_U32_AS_I64: case COVT_OP_VARINT_ZZ_DELTA_AS_I64: case COVT_OP_VARINT_ZZ_AS_I64:
        return CLASS_VARINT32;
    case COVT_OP_VARINT_U64: case COVT_OP_VARINT_ZZ_DELTA_64: return CLASS_VARINT64;
    case COVT_OP_PFOR_ZZ_DELTA: case COVT_OP_PFOR_ZZ_DELTA_XY: case COVT_OP_PFOR_DELTA_MORTON: return CLASS_PFOR;
    default: return -1;
    }
}

int host_op_class_of(uint32_t op) { return op_class_of(op); }

// =================================================================================================
// K0: container walk, one thread per tile (covt_walk.cuh). Two passes over the tile headers: sizes first (the result buffers
// are allocated from their column sums), then the layer table and the dense per-class task lists.
// =================================================================================================
__host__ __device__ constexpr uint32_t slot_stream_type(int s)
{
    return s == COVT_SLOT_ID ? COVT_ST_DATA : s == COVT_SLOT_INDEX ? COVT_ST_INDEX_BUFFER : (uint32_t)(COVT_ST_GEOMETRY_TYPES + (s - COVT_SLOT_TYPES));
}
__host__ __device__ constexpr int slot_buf(int s)
{
    return s == COVT_SLOT_ID ? COVT_BUF_S_IDS : s == COVT_SLOT_TYPES ? COVT_BUF_S_GEOMETRY_TYPES : COVT_BUF_S_GEOMETRY_OFFSETS + (s - COVT_SLOT_GEOM);
}
static_assert(slot_buf(COVT_SLOT_VBUF) == COVT_BUF_S_VERTEX_BUFFER && slot_buf(COVT_SLOT_INDEX) == COVT_BUF_S_INDEX_BUFFER, "slot -> buffer");
static_assert(slot_stream_type(COVT_SLOT_VBUF) == COVT_ST_VERTEX_BUFFER && slot_stream_type(COVT_SLOT_TYPES) == COVT_ST_GEOMETRY_TYPES, "slot -> stream type");

__device__ __forceinline__ uint64_t lite_nv(const Lite& lite, int slot) { return lite.has(slot) ? (uint64_t)lite.nv(slot) : 0ull; }
__device__ __forceinline__ uint64_t lite_vbuf_ints(const Lite& lite, uint32_t geom_ct, uint32_t flags)
{
    uint64_t n = lite_nv(lite, COVT_SLOT_VBUF);
    if (geom_ct == COVT_CT_ICE_MORTON_CODE) n *= 2;
    else if (geom_ct == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) n *= 2;
    return n;
}
// slice size (elements, before the 16-byte rounding) of the layer in result buffer b
__device__ __forceinline__ uint64_t lite_slice_size(const Lite& lite, uint32_t geom_ct, uint32_t flags, int b)
{
    const uint64_t F = lite_nv(lite, COVT_SLOT_TYPES);
    switch (b) {
    case COVT_BUF_S_GEOMETRY_TYPES: return F;
    case COVT_BUF_S_IDS: return lite_nv(lite, COVT_SLOT_ID);
    case COVT_BUF_S_GEOMETRY_OFFSETS: return lite_nv(lite, COVT_SLOT_GEOM);
    case COVT_BUF_S_PART_OFFSETS: return lite_nv(lite, COVT_SLOT_PART);
    case COVT_BUF_S_RING_OFFSETS: return lite_nv(lite, COVT_SLOT_RING);
    case COVT_BUF_S_VERTEX_OFFSETS: return lite_nv(lite, COVT_SLOT_VOFF);
    case COVT_BUF_S_VERTEX_BUFFER: return lite_vbuf_ints(lite, geom_ct, flags);
    case COVT_BUF_S_INDEX_BUFFER: return lite_nv(lite, COVT_SLOT_INDEX);
    default: break;
    }
    if (flags & COVT_FLAG_SKIP_ASSEMBLY) return 0;
    const uint64_t cap_parts = F + lite_nv(lite, COVT_SLOT_PART);
    switch (b) {
    case COVT_BUF_A_GEOM_OFFSETS: return F + 1;
    case COVT_BUF_A_PART_OFFSETS: return cap_parts + 1;
    case COVT_BUF_A_RING_OFFSETS: return cap_parts + lite_nv(lite, COVT_SLOT_RING) + 1;
    case COVT_BUF_A_COORDS: {
        const uint64_t V = lite.has(COVT_SLOT_VOFF) ? lite_nv(lite, COVT_SLOT_VOFF) : lite_vbuf_ints(lite, geom_ct, flags) / 2;
        return 2 * (V + ((flags & COVT_FLAG_CLOSE_RINGS) ? lite_nv(lite, COVT_SLOT_RING) : 0));
    }
    default: return 0;
    }
}

// pass 1: layers per tile + slice sizes per result buffer + tasks per codec class (column-major: col * n_tiles + tile).
// The per-tile sums live in REGISTERS (every index below is a compile-time constant after unrolling; the codec class, the one
// run-time index, is matched by an unrolled compare): as shared-memory arrays (19 x 8 B x 128 threads on top of the layer scratch)
// they took 31 KB per block, i.e. 217 KB of the SM's 256 KB L1/shared array at 7 resident blocks, which left no L1 for the header
// bytes — every byte load of the walk went to L2 (ncu: lts sectors == l1tex sectors, k0_scan_tiles 0.75 -> 2.16 ms per 1 M tiles).
#ifndef K0_SCAN_MINB
#define K0_SCAN_MINB 1
#endif
#ifndef K0_FILL_MINB
#define K0_FILL_MINB 1
#endif
__global__ void __launch_bounds__(K0_BLOCK, K0_SCAN_MINB)
k0_scan_tiles(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
              const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, uint64_t* tile_cols, uint32_t* tile_status)
{
    __shared__ uint32_t s_lite[LITE_WORDS * K0_BLOCK];
    const uint32_t t = blockIdx.x * K0_BLOCK + threadIdx.x;
    if (t >= n_tiles) return;
    const Lite lite = {s_lite + threadIdx.x};
    uint64_t acc[TILE_COLS];
#pragma unroll
    for (int i = 0; i < TILE_COLS; i++) acc[i] = 0;
    NoProps props;
    const uint32_t st = walk_tile(blob, tile_offsets[t], tile_offsets[t + 1], container, tj_fields, tj_layers, lite, props,
                                  [&](uint32_t, const LayerHead& H) {
                                      acc[0] += 1;
#pragma unroll
                                      for (int b = 0; b < COVT_NUM_BUFFERS; b++)
                                          acc[1 + b] += align_elems(lite_slice_size(lite, H.geom_ct, flags, b), kBufElemSizeDev(b));
#pragma unroll
                                      for (int s = 0; s < COVT_NUM_SLOTS; s++) {
                                          const int c = lite.has(s) ? op_class_of(resolve_op(slot_stream_type(s), lite.enc(s), H.geom_ct, flags)) : -1;
#pragma unroll
                                          for (int k = 0; k < NUM_OP_CLASSES; k++) acc[COL_CLASS0 + k] += (c == k) ? 1u : 0u;
                                      }
                                  });
    tile_status[t] = st;
#pragma unroll
    for (int i = 0; i < TILE_COLS; i++) tile_cols[(uint64_t)i * n_tiles + t] = acc[i];
}

// pass 2: the covt_layer table with result offsets (tile_cols now holds exclusive prefixes) and one decode task per present stream,
// appended to the dense list of its codec class
__global__ void __launch_bounds__(K0_BLOCK, K0_FILL_MINB)
k0_fill_layers(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
               const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, const uint64_t* tile_cols, ResultBuffers bufs, covt_layer* layers,
               DeviceTask* tasks, ClassOffsets class_off, uint32_t* first_layer, const SegState* seg, uint64_t* totals)
{
    __shared__ uint32_t s_lite[LITE_WORDS * K0_BLOCK];
    const uint32_t t = blockIdx.x * K0_BLOCK + threadIdx.x;
    if (seg->overflow) return;
    uint64_t payload_bytes = 0, stream_out_bytes = 0;  // -> totals[1], totals[2] (covt_timing)
    if (t < n_tiles) {
    const Lite lite = {s_lite + threadIdx.x};
    // running positions of the tile, in registers (see k0_scan_tiles). tile_cols holds exclusive prefixes inside this segment;
    // seg->base = totals of the segments before it (the class columns stay segment-local: the task lists are rebuilt for every segment)
    uint64_t run[TILE_COLS];
#pragma unroll
    for (int i = 0; i < TILE_COLS; i++)
        run[i] = tile_cols[(uint64_t)i * n_tiles + t] + (i < COL_CLASS0 ? seg->base[i] : class_off.off[i - COL_CLASS0]);
    first_layer[t] = (uint32_t)run[0];
    const uint32_t tile = t + tile_base;
    NoProps props;
    walk_tile(blob, tile_offsets[t], tile_offsets[t + 1], container, tj_fields, tj_layers, lite, props, [&](uint32_t li, const LayerHead& H) {
        const uint32_t layer = (uint32_t)run[0];
        const uint32_t num_bits = 32u - (uint32_t)__clz(H.extent);  // CovtParser.java:77
        // ops + payload offsets (streams of the geometry column follow each other in slot order)
        uint32_t layer_status = COVT_OK;
        uint64_t pay = H.geom_offset;
        const uint64_t F = lite_nv(lite, COVT_SLOT_TYPES);
        const uint64_t cap_parts = (flags & COVT_FLAG_SKIP_ASSEMBLY) ? 0 : F + lite_nv(lite, COVT_SLOT_PART);
        const uint64_t cap_rings = (flags & COVT_FLAG_SKIP_ASSEMBLY) ? 0 : cap_parts + lite_nv(lite, COVT_SLOT_RING);
        uint4* dst = reinterpret_cast<uint4*>(&layers[layer]);
        dst[0] = make_uint4(tile, li, H.extent, H.num_features);
        // (status is patched below once every stream's op is known)
        uint32_t wq[4];
        int q = 2;  // 16-byte chunk being assembled, words 8 ..
        int k = 0;
        auto push = [&](uint32_t w) {
            wq[k++] = w;
            if (k == 4) { dst[q++] = make_uint4(wq[0], wq[1], wq[2], wq[3]); k = 0; }
        };
        push((uint32_t)H.name_offset);
        push((uint32_t)(H.name_offset >> 32));
#pragma unroll
        for (int s = 0; s < COVT_NUM_SLOTS; s++) {
            const bool have = lite.has(s);
            uint64_t off = 0;
            uint32_t op = COVT_OP_NONE, status = COVT_OK;
            if (have) {
                if (s == COVT_SLOT_ID) off = H.id_offset;
                else { off = pay; pay += lite.bl(s); }
                op = resolve_op(slot_stream_type(s), lite.enc(s), H.geom_ct, flags);
                if (op == COVT_OP_NONE) {
                    status = COVT_ERR_UNSUPPORTED_ENCODING;
                    if (!layer_status) layer_status = COVT_ERR_UNSUPPORTED_ENCODING;
                }
            }
            if (have) {
                payload_bytes += lite.bl(s);
                stream_out_bytes += (s == COVT_SLOT_VBUF ? lite_vbuf_ints(lite, H.geom_ct, flags) : (uint64_t)lite.nv(s)) * kBufElemSizeDev(slot_buf(s));
            }
            push((uint32_t)off);
            push((uint32_t)(off >> 32));
            push(have ? lite.bl(s) : 0u);
            push(have ? lite.nv(s) : 0u);
            push((have ? lite.enc(s) : (uint32_t)COVT_ENC_ABSENT) | (op << 8));
            push(status);
            // the decode task of the stream
            const int c = have ? op_class_of(op) : -1;
            if (c >= 0) {
                const int b = slot_buf(s);
                uint32_t nv = lite.nv(s);
                if (s == COVT_SLOT_VBUF && H.geom_ct == COVT_CT_ICE && !(flags & COVT_FLAG_ICE_VB_COUNT_IS_INTS)) nv *= 2;
                const uint64_t dptr = reinterpret_cast<uint64_t>(bufs.ptr[b]) + run[1 + b] * kBufElemSizeDev(b);
                uint64_t task_slot = 0;
#pragma unroll
                for (int k = 0; k < NUM_OP_CLASSES; k++)
                    if (c == k) { task_slot = run[COL_CLASS0 + k]; run[COL_CLASS0 + k] += 1; }
                uint2* tp = reinterpret_cast<uint2*>(&tasks[task_slot]);
                tp[0] = make_uint2((uint32_t)off, (uint32_t)(off >> 32));
                tp[1] = make_uint2((uint32_t)dptr, (uint32_t)(dptr >> 32));
                tp[2] = make_uint2(lite.bl(s), nv);
                tp[3] = make_uint2(op | (num_bits << 8) | ((flags & COVT_FLAG_MORTON_NO_SHIFT) ? 1u << 16 : 0u) | (1u << 24), COVT_OK);
                tp[4] = make_uint2(0u, layer * LAYER_WORDS + STREAM0_STATUS_WORD + (uint32_t)s * STREAM_REF_WORDS);
            }
        }
        // out[13]: element offset of the layer's slice in every result buffer
#pragma unroll
        for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
            const uint64_t o = run[1 + b];
            push((uint32_t)o);
            push((uint32_t)(o >> 32));
            run[1 + b] = o + align_elems(lite_slice_size(lite, H.geom_ct, flags, b), kBufElemSizeDev(b));
        }
        push(0u); push(0u); push(0u); push(0u);  // n_parts, n_rings, n_vertices, n_coords: written by the assembler
        push((uint32_t)cap_parts);
        push((uint32_t)cap_rings);
        push((uint32_t)H.layer_start); push((uint32_t)(H.layer_start >> 32));  // header_offset
        dst[1] = make_uint4(H.num_columns, layer_status, H.geom_ct | (num_bits << 8) | (lite.has(COVT_SLOT_ID) ? 1u << 16 : 0u), H.name_length);
        run[0] += 1;
    });
    }
    // payload / decoded-stream bytes of the batch: one pair of atomics per warp
    for (int d = 16; d >= 1; d >>= 1) {
        payload_bytes += __shfl_down_sync(FULL, payload_bytes, d);
        stream_out_bytes += __shfl_down_sync(FULL, stream_out_bytes, d);
    }
    if ((threadIdx.x & 31u) == 0 && payload_bytes) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[1]), (unsigned long long)payload_bytes);
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[2]), (unsigned long long)stream_out_bytes);
    }
}

// =================================================================================================
// exclusive scan of the TILE_COLS columns over tiles (three small kernels)
// =================================================================================================
constexpr int SCAN_BLOCK = 256;

__device__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total, uint64_t* sm /*[SCAN_BLOCK/32]*/)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t wt;
    uint64_t e = warp_exclusive_scan_u64(v, wt);
    if (lane == 31) sm[warp] = wt;
    __syncthreads();
    uint64_t base = 0, tot = 0;
    for (unsigned w = 0; w < SCAN_BLOCK / 32; w++) {
        if (w < warp) base += sm[w];
        tot += sm[w];
    }
    __syncthreads();
    *total = tot;
    return e + base;
}

__global__ void scan_block_sums(const uint64_t* cols, uint32_t n, uint64_t* block_sums)
{
    __shared__ uint64_t sm[SCAN_BLOCK / 32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t col = blockIdx.y;
    uint64_t v = i < n ? cols[(uint64_t)col * n + i] : 0;
    uint64_t tot;
    block_exclusive_scan_u64(v, &tot, sm);
    if (threadIdx.x == 0) block_sums[(uint64_t)col * gridDim.x + blockIdx.x] = tot;
}
// one block per column walks the block sums
__global__ void scan_block_prefix(uint64_t* block_sums, uint32_t n_blocks, uint64_t* totals)
{
    __shared__ uint64_t sm[SCAN_BLOCK / 32];
    const uint32_t col = blockIdx.x;
    uint64_t running = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += SCAN_BLOCK) {
        const uint32_t b = b0 + threadIdx.x;
        uint64_t v = b < n_blocks ? block_sums[(uint64_t)col * n_blocks + b] : 0;
        uint64_t tot;
        uint64_t e = block_exclusive_scan_u64(v, &tot, sm);
        if (b < n_blocks) block_sums[(uint64_t)col * n_blocks + b] = running + e;
        running += tot;
    }
    if (threadIdx.x == 0) totals[col] = running;
}
__global__ void scan_apply(uint64_t* cols, uint32_t n, const uint64_t* block_sums)
{
    __shared__ uint64_t sm[SCAN_BLOCK / 32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t col = blockIdx.y;
    uint64_t v = i < n ? cols[(uint64_t)col * n + i] : 0;
    uint64_t tot;
    uint64_t e = block_exclusive_scan_u64(v, &tot, sm);
    if (i < n) cols[(uint64_t)col * n + i] = e + block_sums[(uint64_t)col * gridDim.x + blockIdx.x];
}

// =================================================================================================
// Stream decode: ONE KERNEL PER CODEC CLASS over a task table (one warp per stream).
// Each kernel carries a single codec, so its instructions stay cache-resident and every warp of an SM runs the same
// code; an earlier all-codecs-in-one-kernel version stalled ~85 % of its issue slots on instruction fetch
// (profiles/r01_layers_fused_ncu.txt). A warp scans 32 task slots at a time, ballots the ones of its class and
// decodes them one after the other; 32-slot groups are handed out dynamically for load balance.
// =================================================================================================
constexpr int DEC_WARPS = 4;
constexpr int DEC_WARP_SMEM = WARP_SMEM_BYTES;
// the FastPFOR kernel stages whole streams: payload window + value stage per warp
constexpr int PFOR_WARP_SMEM = (PFOR_SMEM_WORDS + 4 + LEAN_STAGE_WORDS) * 4;
constexpr int PFOR_BIG_WARP_SMEM = (404 + LEAN_STAGE_WORDS) * 4;  // pass 2: block window + container window + value stage (warp_pfor_stream)
template <int CLASS> __host__ __device__ constexpr uint32_t class_group() { return (CLASS == CLASS_VARINT32 || CLASS == CLASS_PFOR) ? 1u : 32u; }
constexpr int RLE_BIG_WARP_SMEM = RLE_WIN_WORDS * 4;  // pass 2 of the RLE class: the literal-group window (warp_rle_literals)
template <int CLASS> __host__ __device__ constexpr int class_warp_smem() { return CLASS == CLASS_PFOR ? PFOR_WARP_SMEM : (CLASS == CLASS_RLE ? RLE_BIG_WARP_SMEM : DEC_WARP_SMEM); }

// Work tickets. Every warp of a grid draws its items from ONE counter; at 2.4 M items per launch the same-address atomics alone are a
// millisecond of serialised L2 time (one atomic unit per address), and each ticket is a ~300-cycle round trip at the head of the
// warp's dependent chain (ticket -> task record -> stream bytes). So a warp takes TICKET_BATCH consecutive items per atomic.
// (The lane-per-stream classes take ONE group of 32 streams per ticket: with 4, byte_rle went 0.71 -> 0.78 ms and varint64 1.01 -> 1.11 ms
// per 1 M tiles — a group already is 32 streams, and the longest group of a batch of four decides when the warp is done.)
template <uint32_t TICKET_BATCH>  // a power of two; the counters start at 0, so every batch starts at a multiple of it
struct WarpTickets {
    uint32_t next = 0;  // ONE register of state: a batch is used up when next is a multiple of TICKET_BATCH again
    __device__ __forceinline__ uint32_t take(uint32_t* counter)
    {
        if ((next & (TICKET_BATCH - 1u)) == 0u) {  // warp-uniform
            uint32_t v = 0;
            if (lane_id() == 0) v = atomicAdd(counter, TICKET_BATCH);
            next = __shfl_sync(FULL, v, 0);
        }
        return next++;
    }
};

template <int CLASS, bool BIG = false>
__device__ __forceinline__ void decode_one(const StreamTask& t, void* wsm, StreamOutcome& o)
{
    uint32_t* stage = reinterpret_cast<uint32_t*>(wsm);
    o.status = COVT_OK;
    o.consumed = 0;
    if (CLASS == CLASS_BYTE_RLE) warp_byte_rle_stream(t, o);
    else if (CLASS == CLASS_RLE) {
        if (t.op == COVT_OP_RLE_U32) warp_rle_stream<int32_t, false>(t, o, stage);
        else if (t.op == COVT_OP_RLE_U64) warp_rle_stream<int64_t, false>(t, o, stage);
        else warp_rle_stream<int64_t, true>(t, o, stage);
    } else if (CLASS == CLASS_VARINT32) {
        if (t.op == COVT_OP_VARINT_ZZ_DELTA_XY && (t.num_values & 1u)) { o.status = COVT_ERR_COUNT_MISMATCH; return; }
        const bool widen = t.op == COVT_OP_VARINT_U32_AS_I64 || t.op == COVT_OP_VARINT_ZZ_DELTA_AS_I64 || t.op == COVT_OP_VARINT_ZZ_AS_I64;
        const int post = t.op == COVT_OP_VARINT_U32_AS_I64 ? POST_PLAIN : (t.op == COVT_OP_VARINT_ZZ_DELTA_AS_I64 ? POST_ZZ_DELTA :
                         (t.op == COVT_OP_VARINT_ZZ_AS_I64 ? POST_ZZ : post_kind_of_op(t.op)));
        warp_varint32_stream(t, stage, o, post, widen);
    } else if (CLASS == CLASS_VARINT64) {
        if (t.op == COVT_OP_VARINT_U64) warp_varint64_stream<false>(t, reinterpret_cast<uint64_t*>(wsm), o);
        else warp_varint64_stream<true>(t, reinterpret_cast<uint64_t*>(wsm), o);
    } else {
        if (t.op == COVT_OP_PFOR_ZZ_DELTA_XY && (t.num_values & 1u)) { o.status = COVT_ERR_COUNT_MISMATCH; return; }
        if (!BIG) warp_pfor_stream_smem(t, stage, stage + PFOR_SMEM_WORDS + 4, o, post_kind_of_op(t.op));
        else warp_pfor_stream(t, stage, o, post_kind_of_op(t.op));  // larger than the window: read the page through global memory
    }
}

// Where a stream's outcome goes. Batch path (status_words != nullptr): word d.ref of the status table — the status field of the
// stream's slot in the layer table, or of a property column's record — and only when it is NOT ok: the container walk initialised
// every status to COVT_OK, so the common case costs no scattered store at all. Stream path: status and bytes consumed go back to
// the task, which the host reads. exact_length == 2 (property streams): the stream must end exactly at its byteLength.
__device__ __forceinline__ void report_outcome(DeviceTask* tasks, uint32_t i, const DeviceTask& d, const StreamOutcome& o, uint32_t* status_words)
{
    uint32_t st = o.status;
    if (d.exact_length == 2 && (st == COVT_OK || st == COVT_ERR_VARINT_OVERLONG) && o.consumed != d.byte_length) st = COVT_ERR_COUNT_MISMATCH;
    if (status_words) {
        if (st != COVT_OK) status_words[d.ref] = st;
    } else {
        tasks[i].status = st;
        tasks[i].consumed = o.consumed;
    }
}

// which streams of a class a single thread decodes on its own
template <int CLASS>
__device__ __forceinline__ bool is_small_task(uint32_t num_values, uint32_t byte_length)
{
    if (CLASS == CLASS_BYTE_RLE || CLASS == CLASS_RLE || CLASS == CLASS_VARINT64)
        return num_values <= SMALL_STREAM_VALUES && byte_length <= SMALL_STREAM_BYTES;
    return false;
}

__device__ __forceinline__ StreamTask make_stream_task(const uint8_t* blob, const DeviceTask& d)
{
    StreamTask t;
    t.src = blob + d.src_offset;
    t.dst = d.dst;
    t.byte_length = d.byte_length;
    t.num_values = d.num_values;
    t.op = d.op;
    t.num_bits = d.num_bits;
    t.no_shift = d.no_shift;
    t.exact_length = d.exact_length;
    return t;
}

// Pass 1 of a codec class. A work group = GROUP consecutive tasks of the class, handed out dynamically. Small sequential
// streams are decoded right here by one thread each (32 at a time). Streams that need a whole warp are decoded here only by
// the classes that always take a warp per stream (GROUP == 1); the other classes push them onto big_queue for pass 2 — a warp
// that decoded the large streams of its own group one after the other made the whole kernel wait for the unluckiest group
// (fixture sweep: k_decode_rle 7.2 ms for 2.3 GB of output, the 60 000-value id streams of one tile land in one group).
template <int CLASS, int MINB>
__global__ void __launch_bounds__(DEC_WARPS * 32, MINB)
k_decode_class(const uint8_t* blob, DeviceTask* tasks, uint32_t n_tasks_arg, uint32_t* work_counter, const SegState* seg, uint32_t* status_words,
               uint32_t* big_queue, uint32_t* big_count)
{
    // batch path: the task count of the current segment lives on the device (the host never learns it in the pipelined mode)
    if (seg && seg->overflow) return;
    const uint32_t n_tasks = seg ? (uint32_t)seg->seg_total[COL_CLASS0 + CLASS] : n_tasks_arg;
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint8_t* wsm = smem + warp * class_warp_smem<CLASS>();
    auto report = [&](uint32_t i, const DeviceTask& d, const StreamOutcome& o) { report_outcome(tasks, i, d, o, status_words); };
    constexpr uint32_t GROUP = class_group<CLASS>();
    const uint32_t n_groups = (n_tasks + GROUP - 1u) / GROUP;
#ifndef WARP_CLASS_TICKETS
#define WARP_CLASS_TICKETS 4u
#endif
    WarpTickets<(GROUP == 1u ? WARP_CLASS_TICKETS : 1u)> tickets;
    for (;;) {
        const uint32_t g = tickets.take(work_counter);
        if (g >= n_groups) break;
        const uint32_t mine = g * GROUP + lane;
        const bool have = lane < GROUP && mine < n_tasks;
        DeviceTask d;
        if (have) d = tasks[mine];
        const bool small = have && is_small_task<CLASS>(d.num_values, d.byte_length);
        if (small) {
            const StreamTask t = make_stream_task(blob, d);
            StreamOutcome o = {COVT_OK, 0};
            if (CLASS == CLASS_BYTE_RLE) thread_byte_rle_stream(t, o);
            else if (CLASS == CLASS_RLE) {
                if (t.op == COVT_OP_RLE_U32) thread_rle_stream<int32_t>(t, false, o);
                else thread_rle_stream<int64_t>(t, t.op == COVT_OP_RLE_S64, o);
            } else if (CLASS == CLASS_VARINT64) thread_varint64_stream(t, t.op == COVT_OP_VARINT_ZZ_DELTA_64, o);
            report(mine, d, o);
        }
        __syncwarp();
        unsigned todo = __ballot_sync(FULL, have && !small);
        if (GROUP > 1) {
            // hand the large streams of this group to pass 2
            if (todo) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(big_count, (uint32_t)__popc(todo));
                base = __shfl_sync(FULL, base, 0);
                if ((todo >> lane) & 1u) big_queue[base + __popc(todo & ((1u << lane) - 1u))] = mine;
            }
        } else if (todo) {
            // GROUP == 1: lane 0 holds the stream
            const DeviceTask dw = tasks[g];
            if (CLASS == CLASS_PFOR && dw.byte_length / 4u > PFOR_SMEM_WORDS) {
                // too large for the shared-memory window: pass 2 walks it through global memory, and needs so little shared
                // memory that twice as many warps fit an SM
                if (lane == 0) big_queue[atomicAdd(big_count, 1u)] = g;
            } else {
                const StreamTask t = make_stream_task(blob, dw);
                StreamOutcome o;
                decode_one<CLASS>(t, wsm, o);
                __syncwarp();
                if (lane == 0) report(g, dw, o);
            }
        }
    }
}

// Pass 2: one warp per queued large stream
template <int CLASS>
__global__ void __launch_bounds__(DEC_WARPS * 32)
k_decode_class_big(const uint8_t* blob, DeviceTask* tasks, uint32_t* work_counter, const SegState* seg, uint32_t* status_words,
                   const uint32_t* big_queue, const uint32_t* big_count)
{
    if (seg && seg->overflow) return;
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint8_t* wsm = smem + warp * (CLASS == CLASS_PFOR ? PFOR_BIG_WARP_SMEM : class_warp_smem<CLASS>());
    const uint32_t n = *big_count;
    if (n == 0u) return;  // nothing queued (batches of small tiles): skip the 7 000 same-address ticket atomics of an idle grid
    for (;;) {
        // (one item per ticket here: the queued streams are few and large, balance matters more than the atomic)
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(work_counter, 1u);
        q = __shfl_sync(FULL, q, 0);
        if (q >= n) break;
        const uint32_t i = big_queue[q];
        const DeviceTask dw = tasks[i];
        const StreamTask t = make_stream_task(blob, dw);
        StreamOutcome o;
        decode_one<CLASS, true>(t, wsm, o);
        __syncwarp();
        if (lane == 0) report_outcome(tasks, i, dw, o, status_words);
    }
}

// =================================================================================================
// geometry assembly: one warp per layer (after every stream of the batch has been decoded)
// =================================================================================================
template <int MINB>
__global__ void __launch_bounds__(DEC_WARPS * 32, MINB)
k_assemble_layers(covt_layer* all_layers, ResultBuffers bufs, uint32_t flags, uint32_t* work_counter, const SegState* seg, uint64_t* totals,
                  uint32_t* tile_err)
{
    if (seg->overflow) return;
    covt_layer* layers = all_layers + seg->seg_layer_base;
    const uint32_t n_layers = seg->seg_layers;
    __shared__ uint32_t s_asm[DEC_WARPS][ASM_SMEM_WORDS + 6];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    // One layer per ticket, and no running sums in this kernel (k_layer_totals adds up the layer table afterwards). Measured per 1 M
    // tiles: sums of vertices / output bytes kept per warp (registers or shared memory) 4.96 -> 5.61 ms; 4 layers per ticket 5.14 ->
    // 5.61 ms (neighbouring warps no longer work on neighbouring layers, whose output slices are adjacent).
    WarpTickets<1> tickets;
    for (;;) {
        const uint32_t l = tickets.take(work_counter);
        if (l >= n_layers) break;
        covt_layer* L = &layers[l];
        uint32_t layer_status = L->status;
        if (layer_status == COVT_ERR_BAD_METADATA) continue;
        // the decoders wrote every stream's outcome into the layer record: first error in slot order is the layer's status
        uint32_t st = 0;
        if (lane < COVT_NUM_SLOTS && L->streams[lane].encoding != COVT_ENC_ABSENT) st = L->streams[lane].status;
        const unsigned bad = __ballot_sync(FULL, st != 0);
        if (bad && !layer_status) layer_status = __shfl_sync(FULL, st, __ffs(bad) - 1);
        AsmResult ar = {COVT_OK, 0, 0, 0, 0};
        if (!(flags & COVT_FLAG_SKIP_ASSEMBLY) && layer_status == COVT_OK) {
            LayerIO io;
            auto slice = [&](int slot, int b) -> const void* {
                if (L->streams[slot].encoding == COVT_ENC_ABSENT) return nullptr;
                return reinterpret_cast<const uint8_t*>(bufs.ptr[b]) + L->out[b] * kBufElemSizeDev(b);
            };
            io.types = reinterpret_cast<const uint8_t*>(slice(COVT_SLOT_TYPES, COVT_BUF_S_GEOMETRY_TYPES));
            io.F = L->streams[COVT_SLOT_TYPES].num_values;
            io.geom = reinterpret_cast<const int32_t*>(slice(COVT_SLOT_GEOM, COVT_BUF_S_GEOMETRY_OFFSETS));
            io.n_geom = io.geom ? L->streams[COVT_SLOT_GEOM].num_values : 0;
            io.part = reinterpret_cast<const int32_t*>(slice(COVT_SLOT_PART, COVT_BUF_S_PART_OFFSETS));
            io.n_part = io.part ? L->streams[COVT_SLOT_PART].num_values : 0;
            io.ring = reinterpret_cast<const int32_t*>(slice(COVT_SLOT_RING, COVT_BUF_S_RING_OFFSETS));
            io.n_ring = io.ring ? L->streams[COVT_SLOT_RING].num_values : 0;
            io.voff = reinterpret_cast<const int32_t*>(slice(COVT_SLOT_VOFF, COVT_BUF_S_VERTEX_OFFSETS));
            io.n_voff = io.voff ? L->streams[COVT_SLOT_VOFF].num_values : 0;
            io.vbuf = reinterpret_cast<const int32_t*>(slice(COVT_SLOT_VBUF, COVT_BUF_S_VERTEX_BUFFER));
            const uint64_t vb_ints = vbuf_ints_of(*L, flags);
            io.vbuf_ints = vb_ints;
            io.a_geom = reinterpret_cast<int32_t*>(bufs.ptr[COVT_BUF_A_GEOM_OFFSETS]) + L->out[COVT_BUF_A_GEOM_OFFSETS];
            io.a_part = reinterpret_cast<int32_t*>(bufs.ptr[COVT_BUF_A_PART_OFFSETS]) + L->out[COVT_BUF_A_PART_OFFSETS];
            io.a_ring = reinterpret_cast<int32_t*>(bufs.ptr[COVT_BUF_A_RING_OFFSETS]) + L->out[COVT_BUF_A_RING_OFFSETS];
            io.a_coords = reinterpret_cast<int32_t*>(bufs.ptr[COVT_BUF_A_COORDS]) + L->out[COVT_BUF_A_COORDS];
            io.cap_parts = L->cap_parts;
            io.cap_rings = L->cap_rings;
            const uint64_t V = io.voff ? io.n_voff : vb_ints / 2;
            io.cap_coords = V + ((flags & COVT_FLAG_CLOSE_RINGS) ? io.n_ring : 0);
            io.close_rings = (flags & COVT_FLAG_CLOSE_RINGS) != 0;
            warp_assemble(io, s_asm[warp], ar);
            if (ar.status != COVT_OK) layer_status = ar.status;
        }
        if (lane == 0) {
            // first failing layer of the tile (lowest layer index wins): k_tile_status turns the key into the tile's status
            if (layer_status) atomicMin(&tile_err[L->tile], (min(L->layer_index, 0xffffffu) << 8) | (layer_status & 0xffu));
            L->status = layer_status;
            L->n_parts = ar.n_parts;
            L->n_rings = ar.n_rings;
            L->n_vertices = ar.n_vertices;
            L->n_coords = ar.n_coords;
        }
        __syncwarp();
    }
}

// totals[0] += assembled vertices, totals[2] += bytes of the assembled buffers: one thread per layer of the finished batch
__global__ void k_layer_totals(const covt_layer* layers, uint32_t flags, uint64_t* totals, const SegState* seg)
{
    if (seg->overflow) return;
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_layers = (uint32_t)seg->base[0];
    uint64_t v = 0, b = 0;
    if (l < n_layers) {
        const covt_layer& L = layers[l];
        v = L.n_vertices;
        if (!(flags & COVT_FLAG_SKIP_ASSEMBLY) && L.status == COVT_OK)
            b = 4ull * ((L.streams[COVT_SLOT_TYPES].num_values + 1ull) + (L.n_parts + 1ull) + (L.n_rings + 1ull)) + 8ull * L.n_coords;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        v += __shfl_down_sync(FULL, v, d);
        b += __shfl_down_sync(FULL, b, d);
    }
    if ((threadIdx.x & 31u) == 0 && (v | b)) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[0]), (unsigned long long)v);
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[2]), (unsigned long long)b);
    }
}

// =================================================================================================
// K1: large 32-bit varint streams over many CTAs, wait-free.
//   k1a_aggregate : every warp reads one 512-byte chunk and reduces it to (count, sumEven, sumOdd)
//   k1_scan_*     : segmented exclusive scan of those triples (the x/y combine is associative, not commutative)
//   k1b_decode    : every warp decodes its chunk again, now knowing its exclusive prefix, and writes final values
// A single-pass decoupled look-back version (profiles/r01_a_k1_lookback_4k_tiles_ncu.txt) kept 7 of 8 warps at the
// barrier and is capped near 256 GB/s by the look-back round trip; two cheap passes with no waiting beat it.
// =================================================================================================
__device__ __forceinline__ ChunkState cs_combine(const ChunkState& l, const ChunkState& r)
{
    if (r.flags & 1u) return r;  // r starts a new stream
    if ((r.count | (uint32_t)r.a | (uint32_t)r.b) == 0u) return l;  // empty / padding element
    ChunkState o;
    o.count = l.count + r.count;
    if (r.flags & 2u) {
        const bool odd = l.count & 1u;
        o.a = l.a + (odd ? r.b : r.a);
        o.b = l.b + (odd ? r.a : r.b);
    } else {
        o.a = l.a + r.a;
        o.b = 0;
    }
    o.flags = (l.flags & 1u) | (r.flags & 2u);
    return o;
}
__device__ __forceinline__ ChunkState cs_shfl_up(const ChunkState& t, int d)
{
    ChunkState o;
    o.count = __shfl_up_sync(FULL, t.count, d);
    o.a = __shfl_up_sync(FULL, t.a, d);
    o.b = __shfl_up_sync(FULL, t.b, d);
    o.flags = __shfl_up_sync(FULL, t.flags, d);
    return o;
}
__device__ __forceinline__ ChunkState cs_warp_inclusive(ChunkState v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const ChunkState u = cs_shfl_up(v, d);
        if (lane_id() >= (unsigned)d) v = cs_combine(u, v);
    }
    return v;
}
// inclusive scan over the block; returns this thread's inclusive value, *total = the block's aggregate
__device__ ChunkState cs_block_inclusive(ChunkState v, ChunkState* total, ChunkState* sm /*[32]*/)
{
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    v = cs_warp_inclusive(v);
    if (lane == 31) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        ChunkState w = lane < n_warps ? sm[lane] : ChunkState{0, 0, 0, 0};
        w = cs_warp_inclusive(w);
        if (lane < n_warps) sm[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v = cs_combine(sm[warp - 1], v);
    *total = sm[n_warps - 1];
    __syncthreads();
    return v;
}

// which large stream a launch-wide chunk index belongs to (few streams: a short binary search)
__device__ __forceinline__ uint32_t k1_find_stream(const BigStream* streams, uint32_t n_streams, uint32_t g)
{
    uint32_t lo = 0, hi = n_streams;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (streams[mid].first_chunk <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// A superchunk = K1_SC_WINDOWS consecutive 512-byte windows of one stream, walked by ONE warp front to back: what a window costs
// besides its byte loop (finding the stream, the 64-bit address arithmetic, the warp reductions, the state record) is paid once per
// superchunk, and the lane sums run across windows — a lane only has to know the class (even / odd position inside the
// superchunk) of its first value in each window, which is the parity of the values before it: one ballot + popc.
struct K1Super {
    uintptr_t a0;        // 16-byte aligned address at or before the stream's first byte
    uint64_t total;      // head + byte_length: the stream's window-relative end
    uint32_t head;       // bytes between a0 and the first stream byte
    uint32_t w0, w1;     // windows [w0, w1) of the stream
    uint32_t n_windows;  // windows of the whole stream
};
__device__ __forceinline__ K1Super k1_super(const uint8_t* blob, const BigStream& S, uint32_t sc)
{
    K1Super k;
    const uint8_t* src = blob + S.src_offset;
    k.a0 = reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15);
    k.head = (uint32_t)(reinterpret_cast<uintptr_t>(src) - k.a0);
    k.total = (uint64_t)k.head + S.byte_length;
    k.n_windows = (uint32_t)((k.total + WARP_CHUNK_BYTES - 1) / WARP_CHUNK_BYTES);
    k.w0 = sc * K1_SC_WINDOWS;
    k.w1 = min(k.w0 + (uint32_t)K1_SC_WINDOWS, k.n_windows);
    return k;
}
__device__ __forceinline__ uint4 k1_load_window(const K1Super& k, uint32_t w)
{
    const uint64_t off = (uint64_t)w * WARP_CHUNK_BYTES + lane_id() * 16u;
    uint4 win = make_uint4(0, 0, 0, 0);
    if (off < k.total) win = ldg_stream128(reinterpret_cast<const void*>(k.a0 + off));
    return win;
}
// window-relative bounds of the stream inside window w (only the first and the last window of a stream are partial)
__device__ __forceinline__ bool k1_window_bounds(const K1Super& k, uint32_t w, uint32_t& head_f, uint32_t& tail_f, uint32_t& lo16, uint32_t& hi16)
{
    const unsigned lane = lane_id();
    const uint64_t base = (uint64_t)w * WARP_CHUNK_BYTES;
    head_f = w == 0 ? k.head : 0u;
    const uint32_t end_in = (uint32_t)umin64(WARP_CHUNK_BYTES, k.total - base);
    tail_f = WARP_CHUNK_BYTES - end_in;
    lo16 = head_f > lane * 16u ? min(16u, head_f - lane * 16u) : 0u;
    hi16 = end_in > lane * 16u ? min(16u, end_in - lane * 16u) : 0u;
    return head_f != 0u || tail_f != 0u;
}

template <bool ZZ>
__device__ __forceinline__ void k1a_super(const K1Super& k, int32_t& ta, int32_t& tb, uint32_t& tc)
{
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_le() ^ (1u << lane);
    uint32_t carry_halo = 0;
    if (k.w0 > 0) carry_halo = __ldg(reinterpret_cast<const uint32_t*>(k.a0 + (uint64_t)k.w0 * WARP_CHUNK_BYTES - 4));
    const uint32_t halo_first = carry_halo;
    int32_t A0 = 0, A1 = 0;  // lane sums of the values at even / odd positions of the superchunk (fakes count as positions)
    uint32_t cnt = 0, gpar = 0;
    uint4 nxt = k1_load_window(k, k.w0);
    for (uint32_t w = k.w0; w < k.w1; w++) {
        const uint4 win = nxt;
        if (w + 1 < k.w1) nxt = k1_load_window(k, w + 1);
        uint32_t head_f = 0, tail_f = 0, lo16 = 0, hi16 = 16;
        bool partial = false;
        if (w == 0 || w + 1 == k.n_windows) partial = k1_window_bounds(k, w, head_f, tail_f, lo16, hi16);
        uint32_t wv[4];
        wv[0] = win.x; wv[1] = win.y; wv[2] = win.z; wv[3] = win.w;
        if (partial) lean_mask_window(wv, lo16, hi16);  // warp-uniform branch
        LeanLane L;
        L.cm = cont_mask_scattered(wv);
        L.cnt = 16u - (uint32_t)__popc(L.cm);
        const uint32_t halo = lean_halo(wv[3], carry_halo);
        const bool odd = L.cnt & 1u;
        const unsigned ob = __ballot_sync(FULL, odd);
        const bool fp = (gpar ^ (uint32_t)__popc(ob & lt)) & 1u;  // class of the lane's first value in this window
        int32_t cur = fp ? A1 : A0, oth = fp ? A0 : A1;
        lean_sum_lane_lin<ZZ>(wv, L.cm, lean_lin_carry_in<ZZ>(halo), cur, oth);
        const bool ep = fp != odd;  // class of `cur` after L.cnt swaps
        A0 = ep ? oth : cur;
        A1 = ep ? cur : oth;
        gpar ^= (uint32_t)__popc(ob);
        cnt += L.cnt;
    }
    ta = (int32_t)__reduce_add_sync(FULL, (unsigned)A0);
    tb = (int32_t)__reduce_add_sync(FULL, (unsigned)A1);
    tc = __reduce_add_sync(FULL, cnt);
    // Every byte added its term where it lies, but a value belongs to the superchunk that holds its terminator: the bytes of a value
    // that began before this superchunk come in (class 0: it is the superchunk's first value), those of a value that is still open
    // at its end go out (class = parity of the values counted so far).
    uint32_t acc, mul, ov = 0;
    lean_carry_in(halo_first, acc, mul, ov);
    ta += ZZ ? zigzag_decode32(acc) : (int32_t)acc;
    lean_carry_in(carry_halo, acc, mul, ov);
    const int32_t z_out = ZZ ? zigzag_decode32(acc) : (int32_t)acc;
    if (gpar & 1u) tb -= z_out; else ta -= z_out;
}

__global__ void __launch_bounds__(K1_WARPS * 32)
k1a_aggregate(const uint8_t* blob, const BigStream* streams, uint32_t n_streams, uint32_t n_chunks, ChunkState* states)
{
    const unsigned lane = lane_id();
    const uint32_t g = blockIdx.x * K1_WARPS + (threadIdx.x >> 5);
    if (g >= n_chunks) return;
    const uint32_t si = k1_find_stream(streams, n_streams, g);
    const BigStream S = streams[si];
    const uint32_t sc = g - S.first_chunk;
    const K1Super k = k1_super(blob, S, sc);
    int32_t ta, tb;
    uint32_t tc;
    const bool zz = (S.post == POST_ZZ || S.post == POST_ZZ_DELTA || S.post == POST_ZZ_DELTA_XY);
    if (zz) k1a_super<true>(k, ta, tb, tc);
    else k1a_super<false>(k, ta, tb, tc);
    const uint32_t head_f = sc == 0 ? k.head : 0u;
    const uint32_t tail_f = k.w1 == k.n_windows ? (uint32_t)((uint64_t)k.n_windows * WARP_CHUNK_BYTES - k.total) : 0u;
    if (head_f & 1u) { const int32_t t = ta; ta = tb; tb = t; }  // leading fakes shifted every real value by head_f positions
    const bool xy = S.post == POST_ZZ_DELTA_XY;
    if (lane == 0) {
        ChunkState v;
        v.count = tc - head_f - tail_f;
        v.a = xy ? ta : ta + tb;
        v.b = xy ? tb : 0;
        v.flags = (xy ? 2u : 0u) | (sc == 0 ? 1u : 0u);
        states[g] = v;
    }
    // (the overlong flag is raised by k1b_decode, which knows where the stream's numValues-th value ends)
}

__global__ void __launch_bounds__(K1_SCAN_BLOCK) k1_scan_reduce(const ChunkState* states, uint32_t n, ChunkState* block_states)
{
    __shared__ ChunkState sm[32];
    const uint32_t i = blockIdx.x * K1_SCAN_BLOCK + threadIdx.x;
    ChunkState v = i < n ? states[i] : ChunkState{0, 0, 0, 0};
    ChunkState tot;
    cs_block_inclusive(v, &tot, sm);
    if (threadIdx.x == 0) block_states[blockIdx.x] = tot;
}
// one block: exclusive scan of the block aggregates (in place)
__global__ void __launch_bounds__(K1_SCAN_BLOCK) k1_scan_blocks(ChunkState* block_states, uint32_t nb)
{
    __shared__ ChunkState sm[32];
    ChunkState running = {0, 0, 0, 0};
    for (uint32_t b0 = 0; b0 < nb; b0 += K1_SCAN_BLOCK) {
        const uint32_t b = b0 + threadIdx.x;
        ChunkState v = b < nb ? block_states[b] : ChunkState{0, 0, 0, 0};
        ChunkState tot;
        const ChunkState inc = cs_block_inclusive(v, &tot, sm);
        // exclusive = running (+) inclusive of the previous element
        ChunkState prev = cs_shfl_up(inc, 1);
        __shared__ ChunkState warp_last[32];
        if (lane_id() == 31) warp_last[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (lane_id() == 0) prev = (threadIdx.x == 0) ? ChunkState{0, 0, 0, 0} : warp_last[(threadIdx.x >> 5) - 1];
        const ChunkState excl = threadIdx.x == 0 ? running : cs_combine(running, prev);
        if (b < nb) block_states[b] = excl;
        running = cs_combine(running, tot);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(K1_SCAN_BLOCK) k1_scan_apply(ChunkState* states, uint32_t n, const ChunkState* block_states)
{
    __shared__ ChunkState sm[32];
    __shared__ ChunkState warp_last[32];
    const uint32_t i = blockIdx.x * K1_SCAN_BLOCK + threadIdx.x;
    const ChunkState v = i < n ? states[i] : ChunkState{0, 0, 0, 0};
    ChunkState tot;
    const ChunkState inc = cs_block_inclusive(v, &tot, sm);
    ChunkState prev = cs_shfl_up(inc, 1);
    if (lane_id() == 31) warp_last[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (lane_id() == 0) prev = (threadIdx.x == 0) ? ChunkState{0, 0, 0, 0} : warp_last[(threadIdx.x >> 5) - 1];
    ChunkState excl = threadIdx.x == 0 ? block_states[blockIdx.x] : cs_combine(block_states[blockIdx.x], prev);
    if (v.flags & 1u) excl = ChunkState{0, 0, 0, v.flags};  // first chunk of a stream: nothing before it
    if (i < n) states[i] = excl;
}

// The window of a large stream that holds value #num_values: reports the bytes consumed up to its terminator and, when the
// window goes on behind it, redoes the overlong check on the bytes before the cut only; returns the window's overlong bits.
// Rare and out of line; everything by value so that the caller keeps nothing in local memory for it.
__device__ __noinline__ uint32_t k1b_stream_end(BigStream S, K1Super k, uint32_t w, uint4 win, uint32_t cm, uint32_t cnt, uint32_t excl,
                                                uint32_t count_before, uint32_t halo_in, uint32_t ov)
{
    const unsigned lane = lane_id();
    uint32_t head_f, tail_f, lo16, hi16;
    k1_window_bounds(k, w, head_f, tail_f, lo16, hi16);
    const int64_t first = (int64_t)count_before + excl - head_f;  // stream index of the lane's first terminator (fakes negative)
    const bool mine = (int64_t)S.num_values > first && (int64_t)S.num_values <= first + (int64_t)cnt;
    const unsigned bm = __ballot_sync(FULL, mine);
    uint32_t cut = 0;
    if (mine) cut = lane * 16u + lean_nth_terminator(cm, (uint32_t)((int64_t)S.num_values - first));
    cut = __shfl_sync(FULL, cut, bm ? __ffs(bm) - 1 : 0);
    const uint32_t end_in_chunk = WARP_CHUNK_BYTES - tail_f;
    if (!bm || cut > end_in_chunk) return ov;  // (a terminator behind the stream's bytes is a fake zero)
    if (S.consumed_out && lane == 0) *S.consumed_out = (uint32_t)((uint64_t)w * WARP_CHUNK_BYTES + cut - k.head);
    if (cut < end_in_chunk) {
        const uint32_t hi_cut = cut > lane * 16u ? min(16u, cut - lane * 16u) : 0u;
        uint32_t w2[4], acc2, mul2, halo2 = halo_in;
        int32_t c2, o2;
        ov = 0;
        const LeanLane L2 = lean_front(win, true, lo16, hi_cut, halo2, w2, acc2, mul2, ov);
        lean_sum_lane<false>(w2, L2.cm, acc2, mul2, c2, o2, ov);
    }
    return ov;
}

// Decode pass over one superchunk. Raw values are staged window after window behind whatever the previous windows left over; full
// rows of 256 values (1 KiB of output) leave through lean_row256 — no bounds, two LDS.128 / one scan / two STG.128 per lane — and
// the remainder (< 256 values) moves to the front of the stage. Only the first row of a superchunk (it may start in the middle
// of a 16-byte output vector) and the last partial row take the predicated lean_rows4.
constexpr int K1B_ROW = 256;
constexpr int K1B_STAGE_WORDS = LEAN_FRONT + K1B_ROW + 512 + 16;
template <int POST>
__device__ __forceinline__ void k1b_super(const BigStream& S, const K1Super& k, const ChunkState& P, bool last_super, uint32_t* stage)
{
    const unsigned lane = lane_id();
    uint32_t* A = stage + LEAN_FRONT;
    const uint32_t s4 = P.count & 3u;  // see lean_rows4: A[j] belongs to stream index base + j, base a multiple of 4
    if (lane < 4u) A[lane] = 0;
    uint32_t fill = s4;
    uint64_t base = (uint64_t)P.count - s4;
    bool first_row = s4 != 0u;
    uint32_t count_before = P.count;  // terminators (real values) of the stream before the current window
    int32_t cx = P.a, cy = P.b;
    uint32_t ov_total = 0;
    uint32_t carry_halo = 0;
    if (k.w0 > 0) carry_halo = __ldg(reinterpret_cast<const uint32_t*>(k.a0 + (uint64_t)k.w0 * WARP_CHUNK_BYTES - 4));
    uint4 nxt = k1_load_window(k, k.w0);
    __syncwarp();
    for (uint32_t w = k.w0; w < k.w1; w++) {
        const uint4 win = nxt;
        if (w + 1 < k.w1) nxt = k1_load_window(k, w + 1);
        uint32_t head_f = 0, tail_f = 0, lo16 = 0, hi16 = 16;
        bool partial = false;
        if (w == 0 || w + 1 == k.n_windows) partial = k1_window_bounds(k, w, head_f, tail_f, lo16, hi16);
        uint32_t wv[4], acc, mul, ov = 0;
        const uint32_t halo_in = carry_halo;
        LeanLane L = lean_front(win, partial, lo16, hi16, carry_halo, wv, acc, mul, ov);
        L.excl = warp_exclusive_scan(L.cnt, L.total);
        lean_stage_lane(wv, L.cm, acc, mul, A + fill + L.excl - head_f, ov);
        __syncwarp();
        const uint32_t n_here = L.total - head_f - tail_f;
        if (count_before < S.num_values) {
            // bytes the reference reader consumes = position right after the terminator of value #num_values; what follows it is
            // not read (and must not raise the overlong flag). Only the window that holds that value looks for it.
            if (count_before + n_here >= S.num_values) ov = k1b_stream_end(S, k, w, win, L.cm, L.cnt, L.excl, count_before, halo_in, ov);
            ov_total |= ov;
            fill += min(n_here, S.num_values - count_before);
        }
        count_before += n_here;
        if (fill >= (uint32_t)K1B_ROW) {
            const uint32_t rows = fill / K1B_ROW;  // 1 or 2
            uint32_t r = 0;
            if (first_row) {
                lean_rows4<POST, false>(A, s4, K1B_ROW - s4, S.dst, base + s4, cx, cy, S.num_bits, S.no_shift != 0);
                first_row = false;
                r = 1;
            }
            for (; r < rows; r++) lean_row256<POST>(A + K1B_ROW * r, S.dst, base + K1B_ROW * r, cx, cy, S.num_bits, S.no_shift != 0);
            const uint32_t left = fill - rows * K1B_ROW;
            uint4 mv0 = make_uint4(0, 0, 0, 0), mv1 = mv0;
            if (4u * lane < left) mv0 = *reinterpret_cast<const uint4*>(A + K1B_ROW * rows + 4u * lane);
            if (128u + 4u * lane < left) mv1 = *reinterpret_cast<const uint4*>(A + K1B_ROW * rows + 128u + 4u * lane);
            __syncwarp();
            if (4u * lane < left) *reinterpret_cast<uint4*>(A + 4u * lane) = mv0;
            if (128u + 4u * lane < left) *reinterpret_cast<uint4*>(A + 128u + 4u * lane) = mv1;
            base += (uint64_t)K1B_ROW * rows;
            fill = left;
            __syncwarp();
        }
    }
    if (fill) {
        if (first_row) lean_rows4<POST, false>(A, s4, fill - s4, S.dst, base + s4, cx, cy, S.num_bits, S.no_shift != 0);
        else lean_rows4<POST, false>(A, 0u, fill, S.dst, base, cx, cy, S.num_bits, S.no_shift != 0);
    }
    if (P.count < S.num_values && __any_sync(FULL, (ov_total >> 28) & 1u) && lane == 0) atomicMax(S.status_out, (uint32_t)COVT_ERR_VARINT_OVERLONG);
    if (last_super && lane == 0 && count_before < S.num_values) atomicMax(S.status_out, (uint32_t)COVT_ERR_TRUNCATED);
}

#ifndef K1B_MIN_BLOCKS
#define K1B_MIN_BLOCKS 4  // 64 registers: 32 warps per SM (measured: 1.83 ms vs 1.95 ms at 87 registers / 16 warps on the 1 GiB stream)
#endif
__global__ void __launch_bounds__(K1_WARPS * 32, K1B_MIN_BLOCKS)
k1b_decode(const uint8_t* blob, const BigStream* streams, uint32_t n_streams, uint32_t n_chunks, const ChunkState* states)
{
    __shared__ __align__(16) uint32_t s_stage[K1_WARPS][K1B_STAGE_WORDS];
    const uint32_t g = blockIdx.x * K1_WARPS + (threadIdx.x >> 5);
    if (g >= n_chunks) return;
    const uint32_t si = k1_find_stream(streams, n_streams, g);
    const BigStream S = streams[si];
    const uint32_t sc = g - S.first_chunk;
    const ChunkState P = states[g];
    const K1Super k = k1_super(blob, S, sc);
    uint32_t* stage = s_stage[threadIdx.x >> 5];
    const bool last = sc == S.n_chunks - 1;
    switch (S.post) {
    case POST_PLAIN: k1b_super<POST_PLAIN>(S, k, P, last, stage); break;
    case POST_ZZ: k1b_super<POST_ZZ>(S, k, P, last, stage); break;
    case POST_ZZ_DELTA: k1b_super<POST_ZZ_DELTA>(S, k, P, last, stage); break;
    case POST_ZZ_DELTA_XY: k1b_super<POST_ZZ_DELTA_XY>(S, k, P, last, stage); break;
    default: k1b_super<POST_DELTA_MORTON>(S, k, P, last, stage); break;
    }
}

// Error path of the large streams. k1b_decode reports COVT_ERR_VARINT_OVERLONG when a value with four continuation bytes lies
// before the stream's numValues-th terminator. The Java reader ends such a value after its 4th byte (DecodingUtils.java:157-186),
// so it counts more values than there are terminators: the stream may reach numValues earlier than the terminator count says
// (other bytes consumed) or where the terminator count falls short — or run out of bytes (COVT_ERR_TRUNCATED). One block per
// stream redoes the count the way Java reads. A run (the bytes after a terminator up to and including the next terminator)
// parses the same wherever the parse started before it, so every thread owns the runs that START in its slice of the stream.
constexpr int K1_RESOLVE_THREADS = 256;
struct JavaRunWalk { uint64_t p; uint32_t count; bool truncated; };
// walks the runs that start in [p, hi); stops early once `limit` values are complete (p = right after that value)
__device__ __forceinline__ JavaRunWalk java_walk_runs(const uint8_t* src, uint64_t p, uint64_t hi, uint64_t len, uint64_t limit)
{
    JavaRunWalk w{p, 0u, false};
    uint64_t count = 0;
    while (w.p < hi && w.p < len && count < limit) {
        for (;;) {  // one Java value per trip; leaves at the run's terminator
            int k = 0;
            bool term = false;
            while (k < 4 && w.p < len) { k++; if (!(src[w.p++] & 0x80u)) { term = true; break; } }
            if (!term && k < 4) { w.truncated = true; w.count = (uint32_t)count; return w; }  // the bytes end inside a value
            count++;
            if (term || count >= limit) break;
            if (w.p >= len) break;  // ended on a 4-byte value
        }
    }
    w.count = (uint32_t)umin64(count, 0xffffffffull);
    return w;
}
__global__ void __launch_bounds__(K1_RESOLVE_THREADS) k1_resolve_overlong(const uint8_t* blob, const BigStream* streams)
{
    const BigStream S = streams[blockIdx.x];
    if (*S.status_out != (uint32_t)COVT_ERR_VARINT_OVERLONG) return;  // (block-uniform)
    __shared__ uint64_t s_count[K1_RESOLVE_THREADS];
    __shared__ uint64_t s_start[K1_RESOLVE_THREADS];
    const uint8_t* src = blob + S.src_offset;
    const uint64_t len = S.byte_length;
    const uint64_t per = (len + K1_RESOLVE_THREADS - 1) / K1_RESOLVE_THREADS;
    const uint64_t lo = umin64(len, per * threadIdx.x), hi = umin64(len, lo + per);
    uint64_t p = lo;  // first run start in the slice
    if (p > 0) while (p < hi && (src[p - 1] & 0x80u)) p++;
    if (p > 0 && p == hi && hi > lo && (src[p - 1] & 0x80u)) p = len;  // no run starts here
    if (lo == hi) p = len;
    const JavaRunWalk w = java_walk_runs(src, p, hi, len, ~0ull);
    s_count[threadIdx.x] = w.count;
    s_start[threadIdx.x] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t before = 0;
        uint32_t status = COVT_ERR_TRUNCATED, consumed = (uint32_t)len;
        for (int t = 0; t < K1_RESOLVE_THREADS; t++) {
            if (before + s_count[t] >= S.num_values) {  // value #num_values ends in a run of slice t
                const uint64_t per_t = umin64(len, per * (uint64_t)t + per);
                const JavaRunWalk f = java_walk_runs(src, s_start[t], per_t, len, S.num_values - before);
                status = COVT_ERR_VARINT_OVERLONG;
                consumed = (uint32_t)f.p;
                break;
            }
            before += s_count[t];
        }
        *S.status_out = status;
        if (S.consumed_out) *S.consumed_out = consumed;
    }
}

// =================================================================================================
// finalize: tile status = first layer error (unless the container walk already failed), totals
// =================================================================================================
// per segment, before its fill/decode kernels: publish the segment's layer range, check the capacities, clear the work counters
__global__ void k_seg_begin(SegState* seg, uint32_t* work_counters)
{
    const unsigned i = threadIdx.x;
    if (i < WORK_COUNTERS) work_counters[i] = 0;
    bool over = false;
    if (i < TILE_COLS) over = (i < COL_CLASS0 ? seg->base[i] : 0ull) + seg->seg_total[i] > seg->cap[i];  // task lists are per segment
    over = __any_sync(FULL, over);
    if (i == 0) {
        if (over) seg->overflow = 1;
        seg->seg_layer_base = (uint32_t)seg->base[0];
        seg->seg_layers = (uint32_t)seg->seg_total[0];
    }
}
// per segment, after its kernels: advance the running totals
__global__ void k_seg_end(SegState* seg, uint32_t* first_layer_end)
{
    const unsigned i = threadIdx.x;
    if (seg->overflow) return;
    if (i < TILE_COLS) seg->base[i] += seg->seg_total[i];
    __syncwarp();
    if (i == 0 && first_layer_end) *first_layer_end = (uint32_t)seg->base[0];
}

// tile status = first layer error (the key the assembler left in tile_err), unless the container walk already failed.
// (A pass over the layer table instead cost 0.45 ms per 1 M tiles: 2.1 M scattered 4-byte reads of 368-byte records.)
__global__ void k_tile_status(const uint32_t* tile_err, uint32_t n_tiles, uint32_t* tile_status, const SegState* seg)
{
    if (seg->overflow) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t e = tile_err[t];
    if (e != 0xffffffffu && tile_status[t] == COVT_OK) tile_status[t] = e & 0xffu;
}

// COVT_FLAG_PROFILE_KERNELS only: algorithmic bytes per kernel (payload read + decoded stream written, no padding, no
// intermediates) -> totals[3 .. 7] per codec class, totals[8] assembler
__global__ void k_alg_bytes(const covt_layer* layers, uint32_t n_layers_bound, uint32_t flags, uint64_t* totals, const SegState* seg)
{
    if (seg->overflow) return;
    const uint8_t slot_es[COVT_NUM_SLOTS] = {8, 1, 4, 4, 4, 4, 4, 4};
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_layers = (uint32_t)seg->base[0];
    uint64_t acc[NUM_OP_CLASSES + 1];
    for (int i = 0; i <= NUM_OP_CLASSES; i++) acc[i] = 0;
    if (l < n_layers) {
        const covt_layer& L = layers[l];
        for (int s = 0; s < COVT_NUM_SLOTS; s++) {
            if (L.streams[s].encoding == COVT_ENC_ABSENT) continue;
            const uint64_t nv = s == COVT_SLOT_VBUF ? vbuf_ints_of(L, flags) : (uint64_t)L.streams[s].num_values;
            const int c = op_class_of(L.streams[s].op);
            if (c >= 0) acc[c] += L.streams[s].byte_length + nv * slot_es[s];
        }
        if (!(flags & COVT_FLAG_SKIP_ASSEMBLY) && L.status == COVT_OK) {
            const uint64_t F = slot_nv(L, COVT_SLOT_TYPES);
            const uint64_t wr = 4ull * ((F + 1) + (L.n_parts + 1ull) + (L.n_rings + 1ull)) + 8ull * L.n_coords;
            // SURVEY §8(d) unit cost of the assembler: 4 B offset read + 8 B coordinate written per vertex (+ 8 B gathered through L2
            // for ICE layers); the PLAIN vertices it re-reads from S_VERTEX_BUFFER are an intermediate round trip, not algorithmic bytes
            const uint64_t rd = F + 4ull * (slot_nv(L, COVT_SLOT_GEOM) + slot_nv(L, COVT_SLOT_PART) + slot_nv(L, COVT_SLOT_RING)) +
                                (L.streams[COVT_SLOT_VOFF].encoding != COVT_ENC_ABSENT ? (4ull + 8ull) * L.n_vertices : 0ull);
            acc[NUM_OP_CLASSES] += wr + rd;
        }
    }
    for (int i = 0; i <= NUM_OP_CLASSES; i++) {
        uint64_t v = acc[i];
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(FULL, v, d);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(reinterpret_cast<unsigned long long*>(&totals[3 + i]), (unsigned long long)v);
    }
}

// =================================================================================================
// launchers
// =================================================================================================
cudaError_t launch_k0_scan_tiles(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
                                 const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, uint64_t* tile_cols,
                                 uint32_t* tile_status, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    k0_scan_tiles<<<(n_tiles + K0_BLOCK - 1) / K0_BLOCK, K0_BLOCK, 0, st>>>(blob, tile_offsets, n_tiles, tile_base, container, tj_fields, tj_layers, flags, tile_cols, tile_status);
    return cudaGetLastError();
}

cudaError_t launch_scan_tile_cols(uint64_t* tile_cols, uint32_t n_tiles, uint64_t* block_sums, uint64_t* totals, cudaStream_t st, uint32_t n_cols)
{
    if (!n_tiles) return cudaSuccess;
    const uint32_t nb = (n_tiles + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_block_sums<<<dim3(nb, n_cols), SCAN_BLOCK, 0, st>>>(tile_cols, n_tiles, block_sums);
    scan_block_prefix<<<n_cols, SCAN_BLOCK, 0, st>>>(block_sums, nb, totals);
    scan_apply<<<dim3(nb, n_cols), SCAN_BLOCK, 0, st>>>(tile_cols, n_tiles, block_sums);
    return cudaGetLastError();
}

// ---- property columns (covt_props.cuh) ----
cudaError_t launch_k0_props(bool fill, const uint8_t* blob, const uint64_t* tile_offsets, const covt_layer* layers, uint32_t n_layers, uint32_t container,
                            const uint32_t* tj_fields, uint32_t tj_layers, uint64_t* pcols, const PropOut& out, uint64_t* totals, cudaStream_t st)
{
    if (!n_layers) return cudaSuccess;
    const uint32_t grid = (n_layers + K0_BLOCK - 1) / K0_BLOCK;
    if (fill) k0_props<true><<<grid, K0_BLOCK, 0, st>>>(blob, tile_offsets, layers, n_layers, container, tj_fields, tj_layers, pcols, out, totals);
    else k0_props<false><<<grid, K0_BLOCK, 0, st>>>(blob, tile_offsets, layers, n_layers, container, tj_fields, tj_layers, pcols, out, totals);
    return cudaGetLastError();
}
cudaError_t launch_prop_finish(const uint8_t* blob, uint32_t n_cols, uint32_t n_dicts, const PropOut& out, cudaStream_t st)
{
    if (n_dicts) k_prop_finish_dicts<<<(n_dicts + 3) / 4, 128, 0, st>>>(out.dicts, n_dicts, out.aux, out.aux_dict_base, static_cast<int32_t*>(out.buf[COVT_PBUF_DICT_OFFSETS]));
    if (n_cols) k_prop_finish_columns<<<(n_cols + 3) / 4, 128, 0, st>>>(blob, out.cols, n_cols, out.aux, out.dicts, out);
    return cudaGetLastError();
}

cudaError_t launch_k0_fill_layers(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
                                  const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, const uint64_t* tile_cols,
                                  ResultBuffers bufs, covt_layer* layers, DeviceTask* tasks, ClassOffsets class_off, uint32_t* first_layer,
                                  const SegState* seg, uint64_t* totals, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    k0_fill_layers<<<(n_tiles + K0_BLOCK - 1) / K0_BLOCK, K0_BLOCK, 0, st>>>(blob, tile_offsets, n_tiles, tile_base, container, tj_fields, tj_layers, flags, tile_cols, bufs, layers, tasks, class_off, first_layer, seg, totals);
    return cudaGetLastError();
}

static int grid_for(int sm_count, int per_sm, uint64_t n_items, int items_per_block)
{
    int64_t want = ((int64_t)n_items + items_per_block - 1) / items_per_block;
    int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

const char* op_class_name(int c)
{
    static const char* names[NUM_OP_CLASSES] = {"k_decode_byte_rle", "k_decode_rle", "k_decode_varint32", "k_decode_varint64", "k_decode_pfor"};
    return c >= 0 && c < NUM_OP_CLASSES ? names[c] : "?";
}

cudaError_t launch_seg_begin(SegState* seg, uint32_t* work_counters, cudaStream_t st)
{
    k_seg_begin<<<1, 32, 0, st>>>(seg, work_counters);
    return cudaGetLastError();
}
cudaError_t launch_seg_end(SegState* seg, uint32_t* first_layer_end, cudaStream_t st)
{
    k_seg_end<<<1, 32, 0, st>>>(seg, first_layer_end);
    return cudaGetLastError();
}

// n_tasks: the exact task count (stream path, seg == nullptr) or an upper bound used for the grid size only (batch path).
// counters: [0] group ticket of pass 1, [1] queue length, [2] ticket of pass 2 (all zero before the launch).
// big_queue: n_tasks words (unused by the classes that decode every stream with a warp in pass 1).
cudaError_t launch_decode_class(int op_class, const uint8_t* blob, DeviceTask* tasks, uint32_t n_tasks, uint32_t* counters, uint32_t* big_queue,
                                const SegState* seg, uint32_t* layers, int sm_count, int blocks_per_sm, cudaStream_t st, cudaStream_t big_st,
                                cudaEvent_t pass1_done)
{
    if (!n_tasks) return cudaSuccess;
    // The second pass (queued large streams, a warp each) is a latency-bound tail with few warps: on its own stream it runs beside
    // the first pass of the NEXT codec class instead of holding the whole GPU for itself.
    cudaStream_t st2 = st;
    const bool side = big_st != nullptr && op_class != CLASS_VARINT32;
    // Byte-RLE and RLE never touch the warp stage
    const int smem = (op_class == CLASS_BYTE_RLE || op_class == CLASS_RLE) ? 0 : DEC_WARPS * (op_class == CLASS_PFOR ? PFOR_WARP_SMEM : DEC_WARP_SMEM);
    const int per_sm = blocks_per_sm > 0 ? blocks_per_sm : (op_class == CLASS_PFOR ? 8 : 12);
    const uint32_t group = (op_class == CLASS_VARINT32 || op_class == CLASS_PFOR) ? 1u : 32u;
    const int grid = grid_for(sm_count, per_sm, ((uint64_t)n_tasks + group - 1) / group, DEC_WARPS);
    const int grid_big = grid_for(sm_count, per_sm, n_tasks, DEC_WARPS);
    uint32_t *c0 = counters, *c1 = counters + 1, *c2 = counters + 2;
    // Minimum resident blocks per SM (profiles/r01_experiments.md): the 32-bit varint and FastPFOR kernels are capped at 64
    // registers (8 blocks = 32 warps per SM; below that the chunk decoder spills), the thread-per-stream kernels need no cap.
#define COVT_PASS1(C, MINB, q, qc)                                                                             \
    k_decode_class<C, MINB><<<grid, DEC_WARPS * 32, smem, st>>>(blob, tasks, n_tasks, c0, seg, layers, q, qc); \
    if (side) {                                                                                                \
        cudaEventRecord(pass1_done, st);                                                                       \
        cudaStreamWaitEvent(big_st, pass1_done, 0);                                                            \
        st2 = big_st;                                                                                          \
    }
    switch (op_class) {
    case CLASS_BYTE_RLE:
        COVT_PASS1(CLASS_BYTE_RLE, 1, big_queue, c1);
        k_decode_class_big<CLASS_BYTE_RLE><<<grid_big, DEC_WARPS * 32, smem, st2>>>(blob, tasks, c2, seg, layers, big_queue, c1);
        break;
    case CLASS_RLE:
        COVT_PASS1(CLASS_RLE, 1, big_queue, c1);
        k_decode_class_big<CLASS_RLE><<<grid_big, DEC_WARPS * 32, DEC_WARPS * RLE_BIG_WARP_SMEM, st2>>>(blob, tasks, c2, seg, layers, big_queue, c1);
        break;
    case CLASS_VARINT32: COVT_PASS1(CLASS_VARINT32, 8, nullptr, nullptr); break;
    case CLASS_VARINT64:
        COVT_PASS1(CLASS_VARINT64, 1, big_queue, c1);
        k_decode_class_big<CLASS_VARINT64><<<grid_big, DEC_WARPS * 32, smem, st2>>>(blob, tasks, c2, seg, layers, big_queue, c1);
        break;
    case CLASS_PFOR:
        COVT_PASS1(CLASS_PFOR, 8, big_queue, c1);
        k_decode_class_big<CLASS_PFOR><<<grid_for(sm_count, 12, n_tasks, DEC_WARPS), DEC_WARPS * 32, DEC_WARPS * PFOR_BIG_WARP_SMEM, st2>>>(blob, tasks, c2, seg, layers,
                                                                                                                                       big_queue, c1);
        break;
    default: return cudaErrorInvalidValue;
    }
#undef COVT_PASS1
    return cudaGetLastError();
}

// n_layers_bound: upper bound of the segment's layer count (grid size only)
cudaError_t launch_assemble_layers(covt_layer* layers, uint32_t n_layers_bound, ResultBuffers bufs, uint32_t flags,
                                   uint32_t* work_counter, const SegState* seg, uint64_t* totals, uint32_t* tile_err, int sm_count, cudaStream_t st)
{
    if (!n_layers_bound) return cudaSuccess;
    // 32 registers, 16 blocks = 64 warps per SM: the assembler waits on dependent loads (profiles/r01_experiments.md)
    k_assemble_layers<16><<<grid_for(sm_count, 16, n_layers_bound, DEC_WARPS), DEC_WARPS * 32, 0, st>>>(layers, bufs, flags, work_counter, seg, totals, tile_err);
    return cudaGetLastError();
}

cudaError_t launch_k1_varint_stream(const uint8_t* blob, const BigStream* streams, uint32_t n_streams, uint32_t n_chunks,
                                    ChunkState* states, ChunkState* block_states, cudaStream_t st)
{
    if (!n_chunks) return cudaSuccess;
    const uint32_t grid = (n_chunks + K1_WARPS - 1) / K1_WARPS;
    const uint32_t nb = (n_chunks + K1_SCAN_BLOCK - 1) / K1_SCAN_BLOCK;
    k1a_aggregate<<<grid, K1_WARPS * 32, 0, st>>>(blob, streams, n_streams, n_chunks, states);
    k1_scan_reduce<<<nb, K1_SCAN_BLOCK, 0, st>>>(states, n_chunks, block_states);
    k1_scan_blocks<<<1, K1_SCAN_BLOCK, 0, st>>>(block_states, nb);
    k1_scan_apply<<<nb, K1_SCAN_BLOCK, 0, st>>>(states, n_chunks, block_states);
    k1b_decode<<<grid, K1_WARPS * 32, 0, st>>>(blob, streams, n_streams, n_chunks, states);
    k1_resolve_overlong<<<n_streams, K1_RESOLVE_THREADS, 0, st>>>(blob, streams);  // error path only: returns at once otherwise
    return cudaGetLastError();
}

cudaError_t launch_finalize(const covt_layer* layers, const uint32_t* tile_err, uint32_t n_tiles, uint32_t n_layers_bound, uint32_t flags,
                            uint32_t* tile_status, uint64_t* totals, const SegState* seg, cudaStream_t st)
{
    if (!n_tiles) return cudaSuccess;
    k_tile_status<<<(n_tiles + 255) / 256, 256, 0, st>>>(tile_err, n_tiles, tile_status, seg);
    if (n_layers_bound) k_layer_totals<<<(n_layers_bound + 255) / 256, 256, 0, st>>>(layers, flags, totals, seg);
    if ((flags & COVT_FLAG_PROFILE_KERNELS) && n_layers_bound) k_alg_bytes<<<(n_layers_bound + 255) / 256, 256, 0, st>>>(layers, n_layers_bound, flags, totals, seg);
    return cudaGetLastError();
}

}  // namespace covt
