// covt_internal.h — host<->kernel interface inside libcovt_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/covt_b200.h"

namespace covt {

// value post-processing of the varint / FastPFOR pipelines
enum PostKind { POST_PLAIN = 0, POST_ZZ = 1, POST_ZZ_DELTA = 2, POST_ZZ_DELTA_XY = 3, POST_DELTA_MORTON = 4 };

constexpr int TILE_COLS = 1 + COVT_NUM_BUFFERS;  // column 0 = layers per tile, 1.. = slice sizes per result buffer

struct ResultBuffers { void* ptr[COVT_NUM_BUFFERS]; };

// One stream to decode. The batch path keeps COVT_NUM_SLOTS of them per layer (task index = layer * 8 + slot,
// written by k0_fill_layers); the stream path builds them on the host from covt_stream_desc.
struct DeviceTask {
    uint64_t src_offset;   // into the batch blob
    uint8_t* dst;          // absolute device pointer of the output slice (16-byte aligned)
    uint32_t byte_length;
    uint32_t num_values;
    uint8_t op;            // covt_op; COVT_OP_NONE = nothing to do (absent slot, unsupported, or taken by the look-back kernel)
    uint8_t num_bits, no_shift, exact_length;
    uint32_t status;       // out
    uint32_t consumed;     // out
    uint32_t pad;
};

// codec classes = one small kernel each (the instruction working set of a kernel must stay cache-resident)
enum OpClass { CLASS_BYTE_RLE = 0, CLASS_RLE = 1, CLASS_VARINT32 = 2, CLASS_VARINT64 = 3, CLASS_PFOR = 4, NUM_OP_CLASSES = 5 };

// a large 32-bit delta-varint stream handled by the multi-CTA look-back kernel
struct BigStream {
    uint64_t src_offset;   // into the blob
    uint8_t* dst;          // absolute device pointer of the output slice
    uint32_t byte_length;
    uint32_t num_values;
    uint32_t first_chunk;  // index of its first chunk in the global chunk list
    uint32_t n_chunks;
    uint32_t* status_out;  // where to report the status (task or layer stream slot)
    uint32_t* consumed_out;
    uint8_t post, num_bits, no_shift, pad;
};
struct ChunkRef { uint32_t stream, chunk; };
// look-back record of one chunk, read and written with ONE 128-bit access (flag + payload are never torn):
// flag 0 = nothing yet, 1 = aggregate of this chunk (a/b = sums at even/odd positions relative to the chunk start),
// 2 = inclusive prefix of the stream up to and including this chunk (a/b = running x/y)
struct __align__(16) ChunkState { uint32_t flag, count; int32_t a, b; };
constexpr int K1_WARPS = 8;
constexpr int K1_TILE_BYTES = K1_WARPS * 512;

static const uint8_t kBufElemSize[COVT_NUM_BUFFERS] = {1, 8, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1};

// launchers (covt_kernels.cu); every launcher returns the kernel's launch error
cudaError_t launch_k0_scan_tiles(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                 const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, uint64_t* tile_cols,
                                 uint32_t* tile_status, cudaStream_t st);
cudaError_t launch_scan_tile_cols(uint64_t* tile_cols, uint32_t n_tiles, uint64_t* block_sums, uint64_t* totals, cudaStream_t st);
cudaError_t launch_k0_fill_layers(const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                  const uint32_t* tj_fields, uint32_t tj_layers, uint32_t flags, const uint64_t* tile_cols,
                                  ResultBuffers bufs, covt_layer* layers, DeviceTask* tasks, uint32_t* first_layer, cudaStream_t st);
// one codec class over a task table; work_counter must be zero
cudaError_t launch_decode_class(int op_class, const uint8_t* blob, DeviceTask* tasks, uint32_t n_tasks, uint32_t* work_counter,
                                int sm_count, cudaStream_t st);
const char* op_class_name(int op_class);
cudaError_t launch_assemble_layers(covt_layer* layers, const DeviceTask* tasks, uint32_t n_layers, ResultBuffers bufs, uint32_t flags,
                                   uint32_t* work_counter, int sm_count, cudaStream_t st);
cudaError_t launch_k1_varint_stream(const uint8_t* blob, const BigStream* streams, const ChunkRef* chunks, uint32_t n_chunks,
                                    ChunkState* states, uint32_t* ticket, cudaStream_t st);
cudaError_t launch_finalize(const covt_layer* layers, const uint32_t* first_layer, uint32_t n_tiles, uint32_t flags,
                            uint32_t* tile_status, uint64_t* totals /* [0]=vertices [1]=payload bytes [2]=output bytes */,
                            cudaStream_t st);
uint32_t host_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags);

}  // namespace covt
