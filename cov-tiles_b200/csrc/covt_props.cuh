// covt_props.cuh — property columns on the device.
//
// Replaces CovtParser.decodePropertyColumn (J/decoder/CovtParser.java:276-390) and getStringDictionary (:379-390) with a columnar
// result (include/covt_b200.h: covt_prop_column / covt_prop_dictionary + seven value buffers): a validity bitmap per column (the
// PRESENT stream, Byte-RLE of a java.util.BitSet), the DENSE values of the present features, and per string dictionary an offsets
// array into the tile's own UTF-8 bytes. The stream codecs are the ones of the geometry path (covt_streams.cuh): property streams
// become decode tasks of the same codec-class kernels.
//
//   k0_props<false>        : walk of every layer's metadata (covt_walk.cuh, Props hooks; one thread per layer of the layer table the
//                            geometry pass wrote) counting columns, dictionaries, buffer slices and tasks per layer
//   (column scan)          : exclusive prefixes -> every layer knows where its records, slices and tasks go
//   k0_props<true>         : the same walk writing the column / dictionary records and one decode task per stream
//   k_decode_class<...>    : Byte-RLE (present bitsets, BOOLEAN data), RLE (INT_64 data, dictionary indices, dictionary lengths),
//                            32-bit varints widened to i64 (INT_64 data, CovtParser.java:303-311)
//   k_prop_finish_dicts    : lengths -> offsets (exclusive scan), checked against the dictionary's byteLength
//   k_prop_finish_columns  : validity popcount vs the data stream's numValues, FLOAT/DOUBLE copy (DecodingUtils.decodeFloatsLE
//                            :446-453), dictionary index range check, column status, null expansion (dense values -> one slot per
//                            feature: validity bitmap + values = the Arrow layout)
//
// Column status = the first failure in this order (oracle/covt_oracle.c follows the same order): (1) metadata-only checks (missing
// stream, unsupported type / encoding, a stream that leaves its tile, counts no codec can reach, FLOAT size mismatch), (2) the
// dictionary's status, (3) the PRESENT stream (decode + exact consumption), (4) set bits of the validity bitmap != numValues of the data
// stream, (5) the DATA stream, (6) a dictionary index outside the dictionary.
#pragma once
#include "covt_device.cuh"
#include "covt_walk.cuh"

namespace covt {

constexpr uint32_t PROP_AUX_FILL_ONES = 1u, PROP_AUX_COPY = 2u;  // flags word of a column's aux record (PropOut etc.: covt_internal.h)

__host__ __device__ inline uint32_t kPropBufElemSize(int b)
{
    return (b == COVT_PBUF_VALIDITY || b == COVT_PBUF_BOOL) ? 1u : ((b == COVT_PBUF_I64 || b == COVT_PBUF_F64) ? 8u : 4u);
}

struct PStream { uint64_t off; uint32_t nv, bl, enc; bool have; };

__device__ __forceinline__ uint64_t prop_align(uint64_t n, uint32_t es) { const uint64_t per = 16 / es; return (n + per - 1) / per * per; }

template <bool FILL>
struct PropWalk {
    // context
    const uint8_t* blob;
    uint64_t tile_end;
    uint32_t tile;
    bool gen3;
    PropOut out;
    uint64_t cnt[PROP_COLS];  // pass 1: sums of the tile; pass 2: running positions (start = exclusive prefix)
    uint64_t payload_bytes = 0, output_bytes = 0;  // of the streams handed to the decoders / the slices reserved (covt_timing)
    // layer / column being walked
    uint32_t layer = 0, F = 0, dt = 0, ct = 0, nlen = 0, dict_index = 0;
    uint64_t noff = 0;
    bool has_dict = false, localized = false;
    PStream P, D, L, Y, pend;
    uint64_t pend_sub_off = 0;
    uint32_t pend_sub_len = 0;

    __device__ __forceinline__ bool want() const { return true; }
    __device__ __forceinline__ void set_layer(uint32_t li) { layer = li; }
    __device__ __forceinline__ bool in_tile(const PStream& s) const { return s.off <= tile_end && (uint64_t)s.bl <= tile_end - s.off; }

    __device__ void begin_column(uint64_t name_off, uint32_t name_len, uint32_t data_type, uint32_t column_type, uint32_t num_features)
    {
        noff = name_off; nlen = name_len; dt = data_type; ct = column_type; F = num_features;
        P.have = D.have = L.have = Y.have = pend.have = false;
        has_dict = dt == COVT_DT_STRING && (ct == COVT_CT_DICTIONARY || ct == COVT_CT_LOCALIZED_DICTIONARY);
        localized = dt == COVT_DT_STRING && ct == COVT_CT_LOCALIZED_DICTIONARY;
        if (has_dict) {
            // the record exists from here on (sub-columns refer to it); it stays BAD_METADATA if the walk never ends the column
            dict_index = (uint32_t)cnt[PROP_COL_DICTS]++;
            if (FILL) {
                covt_prop_dictionary d;
                d.tile = tile; d.layer = layer; d.n_entries = 0; d.status = COVT_ERR_BAD_METADATA;
                d.offsets_offset = 0; d.bytes_offset = 0; d.n_bytes = 0;
                out.dicts[dict_index] = d;
                out.aux[out.aux_dict_base + dict_index] = COVT_OK;
            }
        }
    }

    __device__ void stream(uint32_t st, uint64_t sub_off, uint32_t sub_len, uint32_t nv, uint32_t bl, uint32_t enc, uint64_t off)
    {
        PStream s = {off, nv, bl, enc, true};
        if (sub_len == 0) {  // present / data / length / dictionary: the first of each counts
            if (st == COVT_ST_PRESENT) { if (!P.have) P = s; }
            else if (st == COVT_ST_DATA) { if (!D.have) D = s; }
            else if (st == COVT_ST_LENGTH) { if (!L.have) L = s; }
            else if (st == COVT_ST_DICTIONARY) { if (!Y.have) Y = s; }
            return;
        }
        if (!localized) return;  // a stray named stream of a plain column: hopped over
        // localized dictionary (gen-2b fixtures): pairs (present_<s>, <s>), adjacent, sharing the column's dictionary
        if (st == COVT_ST_PRESENT) {
            if (pend.have) emit(pend, pend, pend_sub_off, pend_sub_len, true);  // no partner: BAD_METADATA
            pend = s;
            pend_sub_off = sub_off;
            pend_sub_len = sub_len;
            return;
        }
        if (!pend.have || pend_sub_len != sub_len) return;
        for (uint32_t i = 0; i < sub_len; i++)
            if (__ldg(blob + pend_sub_off + i) != __ldg(blob + sub_off + i)) return;
        emit(pend, s, pend_sub_off, pend_sub_len, false);
        pend.have = false;
    }

    __device__ void end_column()
    {
        if (localized) {
            if (pend.have) emit(pend, pend, pend_sub_off, pend_sub_len, true);
        } else {
            emit(P, D, 0, 0, false);
        }
        if (has_dict) emit_dictionary();
    }

    __device__ void push_task(uint32_t op, const PStream& s, uint32_t num_values, void* dst, uint64_t ref)
    {
        const int c = op_class_of_prop(op);
        payload_bytes += s.bl;
        if (FILL) {
            DeviceTask t;
            t.src_offset = s.off;
            t.dst = reinterpret_cast<uint8_t*>(dst);
            t.byte_length = s.bl;
            t.num_values = num_values;
            t.op = (uint8_t)op;
            t.num_bits = 0;
            t.no_shift = 0;
            t.exact_length = 2;  // the stream must end exactly at its byteLength (COVT_ERR_COUNT_MISMATCH otherwise)
            t.status = COVT_OK;
            t.consumed = 0;
            t.ref = (uint32_t)ref;
            out.tasks[cnt[PROP_COL_CLASS0 + c]] = t;
        }
        cnt[PROP_COL_CLASS0 + c]++;
    }
    __device__ __forceinline__ static int op_class_of_prop(uint32_t op)
    {
        return op == COVT_OP_BYTE_RLE ? CLASS_BYTE_RLE : ((op == COVT_OP_RLE_U32 || op == COVT_OP_RLE_U64 || op == COVT_OP_RLE_S64) ? CLASS_RLE : CLASS_VARINT32);
    }

    // one (sub-)column: plan, slices, record, tasks. orphan = a present_<s> stream without its partner.
    __device__ void emit(const PStream& Ps, const PStream& Ds, uint64_t sub_off, uint32_t sub_len, bool orphan)
    {
        uint32_t st = COVT_OK, kind = COVT_PV_NONE, opD = COVT_OP_NONE, flags = 0;
        int vbuf = -1;
        uint64_t nvals = 0;   // elements reserved in vbuf (bytes for BOOL)
        uint32_t d_count = 0; // values the DATA task produces
        bool useP = false;
        const uint32_t VB = (F + 7u) / 8u;
        if (orphan || dt == 0xFFu || ct > COVT_CT_ICE_MORTON_CODE || !Ds.have) st = COVT_ERR_BAD_METADATA;
        else if (dt == COVT_DT_BOOLEAN) {
            kind = COVT_PV_BOOL;
            if (!in_tile(Ds) || (Ps.have && !in_tile(Ps))) st = COVT_ERR_TRUNCATED;
            else if (!plausible_count((Ds.nv + 7u) / 8u, Ds.bl)) st = COVT_ERR_TRUNCATED;
            else if (Ps.have) { if (!plausible_count(VB, Ps.bl)) st = COVT_ERR_TRUNCATED; else if (Ds.nv > F) st = COVT_ERR_COUNT_MISMATCH; else useP = true; }
            else if (Ds.nv != F) st = COVT_ERR_COUNT_MISMATCH;  // no present stream: every feature has a value (CovtParser.java:280-290)
            else flags |= PROP_AUX_FILL_ONES;
            vbuf = COVT_PBUF_BOOL; nvals = VB; d_count = (Ds.nv + 7u) / 8u; opD = COVT_OP_BYTE_RLE;
        } else if (!Ps.have) st = COVT_ERR_BAD_METADATA;
        else if (dt == COVT_DT_STRING) {
            if (!has_dict) st = COVT_ERR_UNSUPPORTED_ENCODING;  // CovtParser.java:345-347
            else if (!localized && (!L.have || !Y.have)) st = COVT_ERR_BAD_METADATA;
            else if (!in_tile(Ps) || !in_tile(Ds) || !plausible_count(VB, Ps.bl) || !plausible_count(Ds.nv, Ds.bl)) st = COVT_ERR_TRUNCATED;
            else if (Ds.nv > F) st = COVT_ERR_COUNT_MISMATCH;
            kind = COVT_PV_DICT_INDEX; vbuf = COVT_PBUF_DICT_INDEX; nvals = F; d_count = Ds.nv; opD = COVT_OP_RLE_U32; useP = true;
        } else if (dt == COVT_DT_INT_64 || dt == COVT_DT_UINT_64) {
            kind = COVT_PV_I64;
            if (Ds.enc == COVT_ENC_RLE) opD = dt == COVT_DT_INT_64 ? COVT_OP_RLE_S64 : COVT_OP_RLE_U64;  // :299-301
            else if (Ds.enc == COVT_ENC_VARINT_ZIG_ZAG) opD = COVT_OP_VARINT_ZZ_AS_I64;                  // :303-306
            else if (Ds.enc == COVT_ENC_VARINT_DELTA_ZIG_ZAG) opD = COVT_OP_VARINT_ZZ_DELTA_AS_I64;      // :308-311
            else if (Ds.enc == COVT_ENC_VARINT) opD = COVT_OP_VARINT_U32_AS_I64;
            else st = COVT_ERR_UNSUPPORTED_ENCODING;                                                      // :313-315
            if (!st && (!in_tile(Ps) || !in_tile(Ds) || !plausible_count(VB, Ps.bl) || !plausible_count(Ds.nv, Ds.bl))) st = COVT_ERR_TRUNCATED;
            if (!st && Ds.nv > F) st = COVT_ERR_COUNT_MISMATCH;
            vbuf = COVT_PBUF_I64; nvals = F; d_count = Ds.nv; useP = true;
        } else if (dt == COVT_DT_FLOAT || dt == COVT_DT_DOUBLE) {
            const uint32_t es = dt == COVT_DT_FLOAT ? 4u : 8u;
            kind = dt == COVT_DT_FLOAT ? COVT_PV_F32 : COVT_PV_F64;
            if (!in_tile(Ps) || !in_tile(Ds) || !plausible_count(VB, Ps.bl)) st = COVT_ERR_TRUNCATED;
            else if ((uint64_t)Ds.nv * es != Ds.bl || Ds.nv > F) st = COVT_ERR_COUNT_MISMATCH;
            vbuf = dt == COVT_DT_FLOAT ? COVT_PBUF_F32 : COVT_PBUF_F64; nvals = F; flags |= PROP_AUX_COPY; useP = true;
        } else st = COVT_ERR_UNSUPPORTED_ENCODING;  // "Data type not supported", :368-370
        if (st != COVT_OK) { nvals = 0; opD = COVT_OP_NONE; flags = 0; useP = false; }
        const uint64_t vbytes = st == COVT_OK ? VB : 0u;
        const uint64_t col = cnt[PROP_COL_COLUMNS]++;
        const uint64_t v_off = cnt[PROP_COL_BUF0 + COVT_PBUF_VALIDITY];
        cnt[PROP_COL_BUF0 + COVT_PBUF_VALIDITY] += prop_align(vbytes, 1);
        uint64_t d_off = 0;
        if (vbuf >= 0) {
            d_off = cnt[PROP_COL_BUF0 + vbuf];
            cnt[PROP_COL_BUF0 + vbuf] += prop_align(nvals, kPropBufElemSize(vbuf));
        }
        output_bytes += vbytes + nvals * (vbuf >= 0 ? kPropBufElemSize(vbuf) : 0u);
        if (flags & PROP_AUX_COPY) payload_bytes += Ds.bl;  // FLOAT / DOUBLE data is copied by the finish kernel, not by a decode task
        if (FILL) {
            covt_prop_column r;
            r.tile = tile; r.layer = layer;
            r.name_offset = noff; r.sub_offset = sub_len ? sub_off : 0;
            r.name_length = nlen; r.sub_length = sub_len;
            r.data_type = (uint8_t)dt; r.column_type = (uint8_t)ct; r.value_kind = (uint8_t)kind; r.reserved = 0;
            r.status = st;
            r.num_features = F;
            r.num_values = 0;
            r.validity_offset = v_off;
            r.values_offset = d_off;
            r.dictionary = (kind == COVT_PV_DICT_INDEX && has_dict) ? dict_index : 0u;  // (a STRING column of another column type has none: status UNSUPPORTED_ENCODING)
            r.data_num_values = st == COVT_OK ? Ds.nv : 0u;
            out.cols[col] = r;
            uint32_t* a = out.aux + col * PROP_AUX_WORDS;
            a[0] = COVT_OK; a[1] = COVT_OK; a[2] = flags; a[3] = st == COVT_OK ? Ds.bl : 0u;
            a[4] = (uint32_t)Ds.off; a[5] = (uint32_t)(Ds.off >> 32); a[6] = 0; a[7] = 0;
        }
        if (useP) push_task(COVT_OP_BYTE_RLE, Ps, (uint32_t)vbytes, FILL ? static_cast<uint8_t*>(out.buf[COVT_PBUF_VALIDITY]) + v_off : nullptr, col * PROP_AUX_WORDS + 0);
        if (opD != COVT_OP_NONE)
            push_task(opD, Ds, d_count, FILL ? static_cast<uint8_t*>(out.buf[vbuf]) + d_off * kPropBufElemSize(vbuf) : nullptr, col * PROP_AUX_WORDS + 1);
    }

    __device__ void emit_dictionary()
    {
        uint32_t st = COVT_OK, n = 0;
        if (!L.have || !Y.have) st = COVT_ERR_BAD_METADATA;
        else if (!in_tile(L) || !in_tile(Y)) st = COVT_ERR_TRUNCATED;
        else {
            n = gen3 ? Y.nv : L.nv;  // CovtParser.java:352: the DICTIONARY stream's numValues counts the entries
            if (!plausible_count(n, L.bl)) { st = COVT_ERR_TRUNCATED; n = 0; }
        }
        const uint64_t o_off = cnt[PROP_COL_BUF0 + COVT_PBUF_DICT_OFFSETS];
        cnt[PROP_COL_BUF0 + COVT_PBUF_DICT_OFFSETS] += prop_align(st == COVT_OK ? (uint64_t)n + 1u : 0u, 4);
        if (st == COVT_OK) { output_bytes += 4ull * ((uint64_t)n + 1u); payload_bytes += Y.bl; }  // (the dictionary bytes stay in the blob: read by the consumer)
        if (FILL) {
            covt_prop_dictionary d;
            d.tile = tile; d.layer = layer; d.n_entries = n; d.status = st;
            d.offsets_offset = o_off;
            d.bytes_offset = st == COVT_OK ? Y.off : 0u;
            d.n_bytes = st == COVT_OK ? Y.bl : 0u;
            out.dicts[dict_index] = d;
            out.aux[out.aux_dict_base + dict_index] = COVT_OK;
        }
        if (st == COVT_OK)
            push_task(COVT_OP_RLE_U32, L, n, FILL ? static_cast<int32_t*>(out.buf[COVT_PBUF_DICT_OFFSETS]) + o_off + 1 : nullptr, out.aux_dict_base + dict_index);
    }
};

// ONE THREAD PER LAYER of the finished geometry pass (the layer table holds where every complete layer's metadata starts): real
// tiles are few and large — the 91 OMT fixture tiles carry ~12 layers and ~100 property columns each — so a thread per tile left
// the GPU with 5 warps per SM walking 14 KB of metadata each (k0_props 4.5 + 5.1 ms per 23 296 tiles). Property columns exist for
// complete layers only (a tile whose walk fails keeps the columns of the layers before the failure, like its geometry).
// pass 1 (FILL = false): pcols[col * n_layers + layer] = the layer's sums; pass 2 (FILL = true): pcols holds exclusive prefixes
template <bool FILL>
__global__ void __launch_bounds__(K0_BLOCK)
k0_props(const uint8_t* blob, const uint64_t* tile_offsets, const covt_layer* layers, uint32_t n_layers, uint32_t container, const uint32_t* tj_fields,
         uint32_t tj_layers, uint64_t* pcols, PropOut out, uint64_t* totals /* FILL: [0] += payload bytes, [1] += output bytes */)
{
    __shared__ uint32_t s_lite[LITE_WORDS * K0_BLOCK];
    const uint32_t l = blockIdx.x * K0_BLOCK + threadIdx.x;
    if (l >= n_layers) return;
    const Lite lite = {s_lite + threadIdx.x};
    const uint32_t tile = layers[l].tile;
    const uint64_t start = layers[l].header_offset;
    PropWalk<FILL> pw;
    pw.blob = blob;
    pw.tile_end = tile_offsets[tile + 1];
    pw.tile = tile;
    pw.gen3 = container == COVT_CONTAINER_GEN3;
    pw.out = out;
#pragma unroll
    for (int i = 0; i < PROP_COLS; i++) pw.cnt[i] = FILL ? pcols[(uint64_t)i * n_layers + l] + (i >= PROP_COL_CLASS0 ? out.class_off[i - PROP_COL_CLASS0] : 0ull) : 0ull;
    if (layers[l].num_columns > (layers[l].has_id ? 2u : 1u)) {  // (a layer without property columns: nothing to walk)
        Cursor c = {blob, start, pw.tile_end, false};
        LayerHead H;
        pw.set_layer(layers[l].layer_index);
        (void)walk_layer(c, container, tj_fields, tj_layers, lite, H, pw);  // parsed fine a moment ago: the layer is in the table
    }
    if (!FILL) {
#pragma unroll
        for (int i = 0; i < PROP_COLS; i++) pcols[(uint64_t)i * n_layers + l] = pw.cnt[i];
    } else if (pw.payload_bytes | pw.output_bytes) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[0]), (unsigned long long)pw.payload_bytes);
        atomicAdd(reinterpret_cast<unsigned long long*>(&totals[1]), (unsigned long long)pw.output_bytes);
    }
}

// lengths (decoded into offsets[1 .. n]) -> offsets; one warp per dictionary (CovtParser.getStringDictionary :379-390)
__global__ void __launch_bounds__(128) k_prop_finish_dicts(covt_prop_dictionary* dicts, uint32_t n_dicts, const uint32_t* aux, uint64_t aux_dict_base, int32_t* dict_offsets)
{
    const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = lane_id();
    if (d >= n_dicts) return;
    covt_prop_dictionary D = dicts[d];
    if (D.status != COVT_OK) return;
    uint32_t st = aux[aux_dict_base + d];
    if (st == COVT_OK) {
        int32_t* off = dict_offsets + D.offsets_offset;
        uint64_t running = 0;
        bool bad = false;
        for (uint32_t i0 = 0; i0 < D.n_entries; i0 += 32) {
            const uint32_t i = i0 + lane;
            const int32_t len = i < D.n_entries ? off[i + 1] : 0;  // (int)lengthStream[i], :383
            bad = bad || len < 0;
            uint64_t tot;
            const uint64_t incl = warp_exclusive_scan_u64((uint64_t)(len < 0 ? 0 : len), tot) + (uint64_t)(len < 0 ? 0 : len) + running;
            bad = bad || incl > D.n_bytes;
            if (i < D.n_entries) off[i + 1] = (int32_t)incl;
            running += tot;
        }
        if (lane == 0) off[0] = 0;
        if (__any_sync(FULL, bad)) st = COVT_ERR_TRUNCATED;      // a string would run past the dictionary bytes
        else if (running != D.n_bytes) st = COVT_ERR_COUNT_MISMATCH;
    }
    if (lane == 0 && st != COVT_OK) dicts[d].status = st;
}

// one warp per column
__global__ void __launch_bounds__(128) k_prop_finish_columns(const uint8_t* blob, covt_prop_column* cols, uint32_t n_cols, const uint32_t* aux,
                                                              const covt_prop_dictionary* dicts, PropOut out)
{
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = lane_id();
    if (c >= n_cols) return;
    const covt_prop_column C = cols[c];
    if (C.status != COVT_OK) return;
    const uint32_t* a = aux + (uint64_t)c * PROP_AUX_WORDS;
    const uint32_t st_present = a[0], st_data = a[1], flags = a[2], data_bl = a[3];
    const uint64_t data_off = (uint64_t)a[4] | ((uint64_t)a[5] << 32);
    const uint32_t F = C.num_features, VB = (F + 7u) / 8u;
    uint8_t* validity = static_cast<uint8_t*>(out.buf[COVT_PBUF_VALIDITY]) + C.validity_offset;
    uint32_t st = COVT_OK, n_valid = 0;
    uint32_t n_entries = 0;
    if (C.value_kind == COVT_PV_DICT_INDEX) {
        const covt_prop_dictionary D = dicts[C.dictionary];
        st = D.status;
        n_entries = D.n_entries;
    }
    bool present_ok = true;
    if (flags & PROP_AUX_FILL_ONES) {
        for (uint32_t i = lane; i < VB; i += 32) validity[i] = (i + 1u < VB || (F & 7u) == 0u) ? 0xffu : (uint8_t)((1u << (F & 7u)) - 1u);
        n_valid = F;
    } else if (st_present != COVT_OK) {
        present_ok = false;
        if (st == COVT_OK) st = st_present;
    } else {
        // (whole 32-bit words: the slice is 16-byte aligned and padded; bits behind the last feature are masked off)
        const uint32_t* vw = reinterpret_cast<const uint32_t*>(validity);
        const uint32_t nw = (F + 31u) / 32u;
        uint32_t cnt = 0;
        for (uint32_t w = lane; w < nw; w += 32) {
            uint32_t m = vw[w];
            if (w + 1u == nw && (F & 31u)) m &= (1u << (F & 31u)) - 1u;
            cnt += __popc(m);
        }
        n_valid = __reduce_add_sync(FULL, cnt);
    }
    if (st == COVT_OK && n_valid != C.data_num_values) st = COVT_ERR_COUNT_MISMATCH;
    if (st == COVT_OK && (flags & PROP_AUX_COPY)) {
        // DecodingUtils.decodeFloatsLE (:446-453): the little-endian IEEE values of the present features, copied as they are
        uint32_t* dst = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(out.buf[C.value_kind == COVT_PV_F32 ? COVT_PBUF_F32 : COVT_PBUF_F64]) +
                                                    C.values_offset * (C.value_kind == COVT_PV_F32 ? 4u : 8u));
        for (uint32_t w = lane; w < data_bl / 4u; w += 32) dst[w] = ld_u32_unaligned(blob + data_off + 4ull * w);
        __syncwarp();
    }
    if (st == COVT_OK && st_data != COVT_OK) st = st_data;
    if (st == COVT_OK && C.value_kind == COVT_PV_DICT_INDEX) {
        const int32_t* idx = static_cast<const int32_t*>(out.buf[COVT_PBUF_DICT_INDEX]) + C.values_offset;
        bool oob = false;
        for (uint32_t i = lane; i < C.data_num_values; i += 32) { const int32_t v = idx[i]; oob = oob || v < 0 || (uint32_t)v >= n_entries; }
        if (__any_sync(FULL, oob)) st = COVT_ERR_TOPOLOGY;  // Java: ArrayIndexOutOfBounds on dictionaryData[index], :357-358
    }
    // Null expansion (CovtParser.java:317-326, 331-340, 354-364: one Optional per feature): the dense values of the present features
    // sit at the front of the column's F-element slice; spread them to their features' slots, zero the others — the Arrow layout
    // (validity bitmap + one value slot per row). In place and back to front: feature i takes dense value rank(i) <= i, so a chunk of
    // 32 features only reads slots that no later-processed (lower) chunk still needs, and never one a higher chunk has overwritten.
    if (st == COVT_OK && n_valid != F && F != 0u) {
        __syncwarp();
        const uint32_t* vwords = reinterpret_cast<const uint32_t*>(validity);  // the slice is 16-byte aligned and padded
        uint32_t remaining = n_valid;
        void* vals = static_cast<uint8_t*>(out.buf[C.value_kind]) + C.values_offset * kPropBufElemSize(C.value_kind);  // COVT_PV_x == COVT_PBUF_x for x = 1..5
        const int32_t last_chunk = (int32_t)((F - 1u) / 32u);
        constexpr int U = 8;  // chunks of 32 features per trip: their loads are in flight together (a 45 000-feature column is 1 400 chunks)
        for (int32_t top = last_chunk; top >= 0; top -= U) {
            uint32_t m[U], before[U];
#pragma unroll
            for (int k = 0; k < U; k++) {
                const int32_t ch = top - k;
                m[k] = ch >= 0 ? vwords[ch] : 0u;
                if (ch == last_chunk && (F & 31u)) m[k] &= (1u << (F & 31u)) - 1u;
            }
#pragma unroll
            for (int k = 0; k < U; k++) { remaining -= (uint32_t)__popc(m[k]); before[k] = remaining; }
            const uint32_t lt = (1u << lane) - 1u;
            if (C.value_kind == COVT_PV_BOOL) {
                const uint8_t* bits = static_cast<const uint8_t*>(vals);
                uint32_t word[U];
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const uint32_t rank = before[k] + (uint32_t)__popc(m[k] & lt);
                    const uint32_t bit = ((m[k] >> lane) & 1u) ? (bits[rank >> 3] >> (rank & 7u)) & 1u : 0u;
                    word[k] = __ballot_sync(FULL, bit != 0u);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < U; k++)
                    if (lane == (unsigned)k && top - k >= 0) static_cast<uint32_t*>(vals)[top - k] = word[k];
            } else if (C.value_kind == COVT_PV_I64 || C.value_kind == COVT_PV_F64) {
                uint64_t* a = static_cast<uint64_t*>(vals);
                uint64_t v[U];
#pragma unroll
                for (int k = 0; k < U; k++) v[k] = ((m[k] >> lane) & 1u) ? a[before[k] + (uint32_t)__popc(m[k] & lt)] : 0ull;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const int64_t i = ((int64_t)(top - k) << 5) + lane;
                    if (top - k >= 0 && i < (int64_t)F) a[i] = v[k];
                }
            } else {
                uint32_t* a = static_cast<uint32_t*>(vals);
                uint32_t v[U];
#pragma unroll
                for (int k = 0; k < U; k++) v[k] = ((m[k] >> lane) & 1u) ? a[before[k] + (uint32_t)__popc(m[k] & lt)] : 0u;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const int64_t i = ((int64_t)(top - k) << 5) + lane;
                    if (top - k >= 0 && i < (int64_t)F) a[i] = v[k];
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        cols[c].status = st;
        cols[c].num_values = present_ok ? n_valid : 0u;
    }
}

}  // namespace covt
