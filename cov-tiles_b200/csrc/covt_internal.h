// covt_internal.h — host<->kernel interface inside libcovt_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/covt_b200.h"

namespace covt {

// value post-processing of the varint / FastPFOR pipelines
enum PostKind { POST_PLAIN = 0, POST_ZZ = 1, POST_ZZ_DELTA = 2, POST_ZZ_DELTA_XY = 3, POST_DELTA_MORTON = 4 };

constexpr int NUM_OP_CLASSES_C = 5;
// column 0 = layers per tile, 1 .. 13 = slice sizes per result buffer, 14 .. 18 = decode tasks per codec class
constexpr int TILE_COLS = 1 + COVT_NUM_BUFFERS + NUM_OP_CLASSES_C;
constexpr int COL_CLASS0 = 1 + COVT_NUM_BUFFERS;

struct ResultBuffers { void* ptr[COVT_NUM_BUFFERS]; };

// One stream to decode. Tasks are kept in one dense list per codec class (the batch path counts them per tile in k0_scan_tiles
// and writes them in k0_fill_layers; the stream path sorts them on the host), so that every lane of a class kernel has work.
struct DeviceTask {
    uint64_t src_offset;   // into the batch blob
    uint8_t* dst;          // absolute device pointer of the output slice (16-byte aligned)
    uint32_t byte_length;
    uint32_t num_values;
    uint8_t op;            // covt_op; COVT_OP_NONE = nothing to do (absent slot, unsupported, or taken by the look-back kernel)
    uint8_t num_bits, no_shift, exact_length;
    uint32_t status;       // out
    uint32_t consumed;     // out
    uint32_t ref;          // batch path: layer * 8 + slot (where the status goes); stream path: index of the covt_stream_desc
};

// Device-side state of one batch decode. A batch is decoded in one or more SEGMENTS (contiguous tile ranges): with host
// input the upload of segment i+1 overlaps the decode of segment i, and the host never waits to learn a segment's sizes.
struct SegState {
    uint64_t base[TILE_COLS];       // totals of the segments already decoded (col 0 = layers, 1.. = result-buffer elements)
    uint64_t cap[TILE_COLS];        // capacity of the layer table / result buffers (exact for one segment, an estimate otherwise);
                                    // for the class columns: capacity of the class's task list, which is per segment
    uint64_t seg_total[TILE_COLS];  // totals of the current segment (written by the column scan)
    uint32_t overflow;              // sticky: an estimate was too small -> every later kernel is a no-op, the host decodes again
    uint32_t seg_layers, seg_layer_base, pad;
};

// codec classes = one small kernel each (the instruction working set of a kernel must stay cache-resident)
enum OpClass { CLASS_BYTE_RLE = 0, CLASS_RLE = 1, CLASS_VARINT32 = 2, CLASS_VARINT64 = 3, CLASS_PFOR = 4, NUM_OP_CLASSES = 5 };
static_assert(NUM_OP_CLASSES == NUM_OP_CLASSES_C, "TILE_COLS counts the codec classes");
// first task of every class in the task table
struct ClassOffsets { uint64_t off[NUM_OP_CLASSES]; };

// A large 32-bit varint stream, decoded by many CTAs in two wait-free passes (k1a aggregates -> segmented scan -> k1b).
struct BigStream {
    uint64_t src_offset;   // into the blob
    uint8_t* dst;          // absolute device pointer of the output slice
    uint32_t byte_length;
    uint32_t num_values;
    uint32_t first_chunk;  // index of its first superchunk (K1_SC_BYTES) in the launch-wide numbering
    uint32_t n_chunks;
    uint32_t* status_out;  // where to report the status (a DeviceTask)
    uint32_t* consumed_out;
    uint8_t post, num_bits, no_shift, pad;
};
// Per-chunk record: after k1a the chunk's aggregate (count, sum at even positions, sum at odd positions),
// after the scan the exclusive prefix of the chunk inside its stream (count, running x, running y).
// flags: bit 0 = first chunk of its stream, bit 1 = x/y interleaved sums
struct __align__(16) ChunkState { uint32_t count; int32_t a, b; uint32_t flags; };
constexpr int WORK_COUNTERS = 16;  // 3 per codec class (pass-1 ticket, queue length, pass-2 ticket) + the assembler's ticket
constexpr int FINAL_TOTALS = 9;  // [0] vertices (assembler) [1] payload bytes [2] output bytes (container walk + assembler) [3..7] algorithmic bytes per codec class, [8] assembler (profiling only)
constexpr int K1_WARPS = 8;
constexpr int K1_SC_WINDOWS = 16;  // 512-byte windows per superchunk (the unit one warp walks; one ChunkState each)
constexpr int K1_SC_BYTES = K1_SC_WINDOWS * 512;
constexpr int K1_SCAN_BLOCK = 1024;

static const uint8_t kBufElemSize[COVT_NUM_BUFFERS] = {1, 8, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1};

// launchers (covt_kernels.cu); every launcher returns the kernel's launch error
cudaError_t launch_k0_scan_tiles(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
                                 const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, uint64_t* tile_cols,
                                 uint32_t* tile_status, cudaStream_t st);
// exclusive scan over tiles of n_cols column-major columns (in place); block_sums: n_cols * ceil(n_tiles / 256) words; totals[n_cols]
cudaError_t launch_scan_tile_cols(uint64_t* tile_cols, uint32_t n_tiles, uint64_t* block_sums, uint64_t* totals, cudaStream_t st, uint32_t n_cols = TILE_COLS);
cudaError_t launch_seg_begin(SegState* seg, uint32_t* work_counters /*[WORK_COUNTERS]*/, cudaStream_t st);
cudaError_t launch_seg_end(SegState* seg, uint32_t* first_layer_end /*nullable*/, cudaStream_t st);
cudaError_t launch_k0_fill_layers(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t tile_base, uint32_t container,
                                  const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, const uint64_t* tile_cols,
                                  ResultBuffers bufs, covt_layer* layers, DeviceTask* tasks, ClassOffsets class_off, uint32_t* first_layer,
                                  const SegState* seg, uint64_t* totals /* [FINAL_TOTALS] */, cudaStream_t st);
// One codec class over ITS dense task list (pass 1: small streams by threads; pass 2: queued large streams, a warp each).
// counters: 3 zeroed words (see covt_kernels.cu); big_queue: n_tasks words. seg != nullptr (batch path): the task count is
// seg->seg_total[COL_CLASS0 + class] and n_tasks only bounds the grids. status_words != nullptr: a stream's status (when not OK) goes
// to status_words[task.ref] (the layer table or the property-column records viewed as uint32 words); nullptr (stream path): status
// and bytes consumed go back to the task. blocks_per_sm: 0 = fill the GPU with this kernel alone.
cudaError_t launch_decode_class(int op_class, const uint8_t* blob, DeviceTask* tasks, uint32_t n_tasks, uint32_t* counters, uint32_t* big_queue,
                                const SegState* seg, uint32_t* status_words, int sm_count, int blocks_per_sm, cudaStream_t st,
                                cudaStream_t big_st = nullptr /* second pass on its own stream, after pass1_done */, cudaEvent_t pass1_done = nullptr);
const char* op_class_name(int op_class);
cudaError_t launch_assemble_layers(covt_layer* layers, uint32_t n_layers_bound, ResultBuffers bufs, uint32_t flags,
                                   uint32_t* work_counter, const SegState* seg, uint64_t* totals /* [FINAL_TOTALS] */,
                                   uint32_t* tile_err /* [n_tiles of the batch], 0xffffffff = no failing layer */, int sm_count, cudaStream_t st);
// k1a_aggregate + 3 segmented-scan kernels + k1b_decode; block_states needs ceil(n_chunks / K1_SCAN_BLOCK) entries
cudaError_t launch_k1_varint_stream(const uint8_t* blob, const BigStream* streams, uint32_t n_streams, uint32_t n_chunks,
                                    ChunkState* states, ChunkState* block_states, cudaStream_t st);
// per-tile status; with COVT_FLAG_PROFILE_KERNELS also the algorithmic bytes per kernel (totals[3 ..])
cudaError_t launch_finalize(const covt_layer* layers, const uint32_t* tile_err, uint32_t n_tiles, uint32_t n_layers_bound, uint32_t flags,
                            uint32_t* tile_status, uint64_t* totals /* [FINAL_TOTALS] */, const SegState* seg, cudaStream_t st);
// ---- property columns (covt_props.cuh) ----
// per-tile columns of the property scan: 0 = columns, 1 = dictionaries, 2..8 = elements per value buffer, 9..13 = tasks per codec class
constexpr int PROP_COL_COLUMNS = 0, PROP_COL_DICTS = 1, PROP_COL_BUF0 = 2, PROP_COL_CLASS0 = PROP_COL_BUF0 + COVT_NUM_PROP_BUFFERS;
constexpr int PROP_COLS = PROP_COL_CLASS0 + NUM_OP_CLASSES_C;
constexpr int PROP_AUX_WORDS = 8;  // per column: [0] PRESENT status [1] DATA status [2] flags [3] DATA byteLength [4,5] DATA offset
static const uint8_t kPropBufElemSizeHost[COVT_NUM_PROP_BUFFERS] = {1, 8, 4, 8, 1, 4, 4};
struct PropOut {
    covt_prop_column* cols;
    covt_prop_dictionary* dicts;
    uint32_t* aux;             // PROP_AUX_WORDS words per column, then one status word per dictionary
    uint64_t aux_dict_base;    // word index of dictionary 0's status
    DeviceTask* tasks;
    uint64_t class_off[NUM_OP_CLASSES_C];
    void* buf[COVT_NUM_PROP_BUFFERS];
};
// One thread per layer of the layer table. fill = false: pcols[col * n_layers + layer] = per-layer sums (out is not touched);
// fill = true: pcols holds the exclusive prefixes. totals (fill only): [0] += payload bytes of the property streams, [1] += bytes
// of the property buffers written
cudaError_t launch_k0_props(bool fill, const uint8_t* blob, const uint64_t* tile_offsets, const covt_layer* layers, uint32_t n_layers, uint32_t container,
                            const uint32_t* tj_fields, uint32_t tj_layers, uint64_t* pcols, const PropOut& out, uint64_t* totals, cudaStream_t st);
cudaError_t launch_prop_finish(const uint8_t* blob, uint32_t n_cols, uint32_t n_dicts, const PropOut& out, cudaStream_t st);
// ---- stream encoders (covt_encode.cu) ----
constexpr uint32_t ENC_VARINT_PIECE = 4096;  // values per piece of a varint stream (one warp)
// One unit of encode work: a range of a varint stream, or a whole RLE / Byte-RLE / FastPFOR stream.
struct EncPiece {
    uint64_t stream_values;  // byte offset of the STREAM's first value in the device copy of the caller's values
    uint8_t* scratch;        // 16-byte aligned, bounded range the piece is written to (FastPFOR: zeroed)
    uint64_t out_offset;     // byte offset of the piece in the result arena (set by the host between encode and compaction)
    uint32_t first_index;    // index of the piece's first value in its stream (the value before it feeds its delta)
    uint32_t num_values;
    uint32_t byte_length;    // out
    uint32_t status;         // out
    uint32_t stream;         // index of the covt_encode_desc
    uint8_t op, num_bits, pad[2];
};
cudaError_t launch_encode_pieces(const uint8_t* values, EncPiece* varint_pieces, uint32_t n_varint, EncPiece* rle_pieces, uint32_t n_rle,
                                 EncPiece* pfor_pieces, uint32_t n_pfor, uint32_t flags, cudaStream_t st);
cudaError_t launch_encode_compact(const EncPiece* pieces, uint32_t n_pieces, uint8_t* arena, cudaStream_t st);
uint32_t host_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags);
int host_op_class_of(uint32_t op);  // OpClass of a covt_op, -1 for COVT_OP_NONE

}  // namespace covt
