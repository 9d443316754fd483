// covt_api.cu — host side of libcovt_b200: contexts, batches, results and the C ABI of include/covt_b200.h.
//
// The host does no per-tile work: it uploads the blob, launches the device-side container walk,
// reads back ONE small totals record to size the result buffers, launches the decode kernels and
// hands out device pointers. There is no CPU fallback: without a CUDA device every call fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "covt_internal.h"

using namespace covt;

namespace {

std::mutex g_err_mutex;
std::string g_create_error;

struct KernelRecord {
    std::string name;
    cudaEvent_t a, b;
    uint64_t alg_bytes;
};

}  // namespace

struct covt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // kernels, allocations, result read-backs
    cudaStream_t copy_stream = nullptr;  // host->device segments of covt_decode_batch, overlapped with the kernels
    cudaStream_t out_stream = nullptr;   // device->host segments of covt_decode_batch_to_host (PCIe is full duplex: its own stream)
    cudaStream_t big_stream = nullptr;   // second passes of the codec classes (large streams, a warp each): beside the next class's first pass
    cudaEvent_t ev_pass1[covt::NUM_OP_CLASSES] = {}, ev_big_done = nullptr;
    bool side_big = true;                // env COVT_SERIAL_BIG=1: second passes on the main stream (round-1 behaviour)
    int class_chains = 0;                // env COVT_CHAINS=1|2|3|5: how many chains of codec kernels run side by side; 0 = by segment size (see decode_segments)
    std::string err;
    uint64_t* h_totals = nullptr;  // pinned scratch for the one device->host size read-back
    covt::SegState* h_seg = nullptr;     // pinned mirror of the device-side segment state
    uint64_t seg_bytes = 64ull << 20;    // minimum upload/decode segment size of covt_decode_batch (env COVT_SEG_BYTES overrides: tests)
    uint32_t max_segments = 24;          // env COVT_MAX_SEGMENTS overrides (1 M tiles end to end: 8 -> 54.4 ms, 16 -> 53.5, 24 -> 53.2, 32 -> 53.3)
    uint32_t seg_min_tiles = 32768;      // env COVT_SEG_MIN_TILES overrides
    bool debug = false;                  // env COVT_DEBUG: host-side phase times on stderr
    bool serial_classes = true;          // env COVT_CONCURRENT=1 runs the five codec kernels side by side on their own streams
                                         // (measured SLOWER on B200, 6.61 vs 5.75 ms per 262k tiles: each kernel alone already
                                         // fills the SMs' register file, so the grids only get in each other's way)
    cudaStream_t class_stream[covt::NUM_OP_CLASSES] = {};  // the codec kernels of a segment run side by side
    cudaEvent_t ev_fork = nullptr, ev_join[covt::NUM_OP_CLASSES] = {};
    std::vector<std::pair<void*, uint64_t>> dev_cache;   // parked device blocks (see dev_alloc_bytes)
    std::unordered_map<void*, uint64_t> dev_live;         // cache-eligible blocks in use
    std::vector<std::pair<void*, size_t>> pinned_cache;  // page-locked host blocks handed to results and taken back on free
    uint32_t overflow_retries = 0;       // how often an extrapolated capacity was too small (diagnostics, tests)
    std::vector<cudaEvent_t> event_pool;
};

struct covt_batch {
    covt_ctx* ctx = nullptr;
    uint8_t* d_blob = nullptr;
    uint64_t blob_len = 0;
    uint64_t* d_tile_offsets = nullptr;
    uint32_t n_tiles = 0;
    float h2d_ms = 0.f;
};

struct covt_result {
    covt_ctx* ctx = nullptr;
    uint32_t n_tiles = 0, n_layers = 0;
    covt_layer* d_layers = nullptr;
    uint32_t* d_tile_status = nullptr;
    uint32_t* d_first_layer = nullptr;
    void* arena = nullptr;               // ONE device allocation behind bufs[] (fewer, larger pool blocks: stable reuse)
    void* bufs[COVT_NUM_BUFFERS] = {};
    uint64_t counts[COVT_NUM_BUFFERS] = {};
    covt_layer* h_layers = nullptr;      // pinned, lazily fetched
    uint32_t* h_tile_status = nullptr;   // pinned, lazily fetched
    uint32_t* h_first_layer = nullptr;
    // property columns (COVT_FLAG_DECODE_PROPERTIES)
    uint32_t n_prop_cols = 0, n_prop_dicts = 0;
    covt_prop_column* d_prop_cols = nullptr;
    covt_prop_dictionary* d_prop_dicts = nullptr;
    void* prop_arena = nullptr;
    void* pbufs[COVT_NUM_PROP_BUFFERS] = {};
    uint64_t pcounts[COVT_NUM_PROP_BUFFERS] = {};
    covt_prop_column* h_prop_cols = nullptr;       // pinned, lazily fetched
    covt_prop_dictionary* h_prop_dicts = nullptr;
    covt_timing timing = {};
    std::vector<covt_kernel_time> kernel_times;
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof(m_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = m_;                                                                            \
            return e_ == cudaErrorMemoryAllocation ? COVT_ERR_OOM : COVT_ERR_CUDA;                    \
        }                                                                                             \
    } while (0)

namespace {

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int32_t fail(covt_ctx* ctx, int32_t code, const char* msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// Device memory. Large blocks (>= 1 MiB) are kept by the context after use and handed out again: the stream-ordered pool
// alone re-maps physical memory whenever the sizes of successive requests differ (measured: 4 .. 640 ms for the 19.7 GB
// result arena of a 1 M tile batch), which would dominate the end-to-end time of back-to-back decode calls. Reuse is in
// stream order on ctx->stream, so no synchronisation is needed.
constexpr uint64_t DEV_CACHE_MIN = 1ull << 20;
constexpr size_t DEV_CACHE_BLOCKS = 24;

cudaError_t dev_alloc_bytes(covt_ctx* ctx, void** p, uint64_t bytes)
{
    bytes = std::max<uint64_t>(bytes, 64);
    if (bytes >= DEV_CACHE_MIN) {
        size_t best = SIZE_MAX;
        for (size_t i = 0; i < ctx->dev_cache.size(); i++) {
            const uint64_t have = ctx->dev_cache[i].second;
            if (have >= bytes && have <= bytes + bytes / 2 && (best == SIZE_MAX || have < ctx->dev_cache[best].second)) best = i;
        }
        if (best != SIZE_MAX) {
            *p = ctx->dev_cache[best].first;
            ctx->dev_live[*p] = ctx->dev_cache[best].second;
            ctx->dev_cache.erase(ctx->dev_cache.begin() + best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation && !ctx->dev_cache.empty()) {
        // out of memory with blocks parked in the cache: give them back and try once more
        (void)cudaGetLastError();
        for (auto& b : ctx->dev_cache) cudaFreeAsync(b.first, ctx->stream);
        ctx->dev_cache.clear();
        cudaStreamSynchronize(ctx->stream);
        e = cudaMallocAsync(p, bytes, ctx->stream);
    }
    if (e == cudaSuccess && bytes >= DEV_CACHE_MIN) ctx->dev_live[*p] = bytes;
    return e;
}
template <class T>
cudaError_t dev_alloc(covt_ctx* ctx, T** p, uint64_t count)
{
    return dev_alloc_bytes(ctx, reinterpret_cast<void**>(p), std::max<uint64_t>(count, 1) * sizeof(T) + 64);
}
void dev_free(covt_ctx* ctx, void* p)
{
    if (!p) return;
    auto it = ctx->dev_live.find(p);
    if (it != ctx->dev_live.end()) {
        const uint64_t bytes = it->second;
        ctx->dev_live.erase(it);
        if (ctx->dev_cache.size() < DEV_CACHE_BLOCKS) { ctx->dev_cache.emplace_back(p, bytes); return; }
        // cache full: drop the smallest parked block instead if this one is larger
        size_t smallest = 0;
        for (size_t i = 1; i < ctx->dev_cache.size(); i++) if (ctx->dev_cache[i].second < ctx->dev_cache[smallest].second) smallest = i;
        if (ctx->dev_cache[smallest].second < bytes) {
            cudaFreeAsync(ctx->dev_cache[smallest].first, ctx->stream);
            ctx->dev_cache[smallest] = {p, bytes};
            return;
        }
    }
    cudaFreeAsync(p, ctx->stream);
}

// Page-locked host memory is expensive to allocate (milliseconds): results borrow blocks from the context and give them back.
cudaError_t pinned_take(covt_ctx* ctx, void** p, size_t bytes)
{
    size_t best = SIZE_MAX;
    for (size_t i = 0; i < ctx->pinned_cache.size(); i++)
        if (ctx->pinned_cache[i].second >= bytes && (best == SIZE_MAX || ctx->pinned_cache[i].second < ctx->pinned_cache[best].second)) best = i;
    if (best != SIZE_MAX && ctx->pinned_cache[best].second <= 2 * bytes + 4096) {
        *p = ctx->pinned_cache[best].first;
        ctx->pinned_cache.erase(ctx->pinned_cache.begin() + best);
        return cudaSuccess;
    }
    return cudaMallocHost(p, std::max<size_t>(bytes, 64));
}
void pinned_give(covt_ctx* ctx, void* p, size_t bytes)
{
    if (!p) return;
    if (ctx->pinned_cache.size() >= 8) { cudaFreeHost(p); return; }
    ctx->pinned_cache.emplace_back(p, std::max<size_t>(bytes, 64));
}

// per-kernel timing (COVT_FLAG_PROFILE_KERNELS)
struct Profiler {
    covt_ctx* ctx;
    bool on;
    std::vector<KernelRecord> recs;
    void begin(const char* name, uint64_t alg_bytes)
    {
        if (!on) return;
        KernelRecord r;
        r.name = name;
        r.alg_bytes = alg_bytes;
        r.a = take_event();
        r.b = take_event();
        cudaEventRecord(r.a, ctx->stream);
        recs.push_back(r);
    }
    // timing events are pooled in the context: a profiled decode records ~25 pairs
    cudaEvent_t take_event()
    {
        cudaEvent_t e = nullptr;
        if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    void end()
    {
        if (!on) return;
        cudaEventRecord(recs.back().b, ctx->stream);
    }
    void collect(std::vector<covt_kernel_time>& out)
    {
        for (auto& r : recs) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            ctx->event_pool.push_back(r.a);
            ctx->event_pool.push_back(r.b);
            bool merged = false;
            for (auto& k : out)
                if (r.name == k.name) { k.ms += ms; k.launches++; k.algorithmic_bytes += r.alg_bytes; merged = true; break; }
            if (!merged) {
                covt_kernel_time k;
                memset(&k, 0, sizeof(k));
                snprintf(k.name, sizeof(k.name), "%s", r.name.c_str());
                k.ms = ms;
                k.launches = 1;
                k.algorithmic_bytes = r.alg_bytes;
                out.push_back(k);
            }
        }
        recs.clear();
    }
};

}  // namespace

// host copy of a record table of the result (pinned, fetched on first use)
template <class T>
static int32_t fetch_records(covt_result* res, T** host, const T* dev, uint64_t n)
{
    covt_ctx* ctx = res->ctx;
    if (*host) return COVT_OK;
    CK(cudaSetDevice(ctx->device));
    CK(pinned_take(ctx, reinterpret_cast<void**>(host), std::max<uint64_t>(n, 1) * sizeof(T)));
    if (n) CK(cudaMemcpyAsync(*host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return COVT_OK;
}

extern "C" {

int32_t covt_abi_version(void) { return COVT_ABI_VERSION; }

int32_t covt_create(int32_t device, covt_ctx** out)
{
    if (!out) return COVT_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        std::lock_guard<std::mutex> g(g_err_mutex);
        g_create_error = e != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(e)
                                          : "CUDA device ordinal out of range (libcovt_b200 has no CPU fallback)";
        return COVT_ERR_CUDA;
    }
    covt_ctx* ctx = new covt_ctx();
    ctx->device = device;
    auto bail = [&](const char* what, cudaError_t err) {
        std::lock_guard<std::mutex> g(g_err_mutex);
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return COVT_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        std::lock_guard<std::mutex> g(g_err_mutex);
        g_create_error = "libcovt_b200 carries sm_100a kernels only; this device is not a Blackwell B200";
        delete ctx;
        return COVT_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* sb = getenv("COVT_SEG_BYTES")) { const long long v = atoll(sb); if (v > 0) ctx->seg_bytes = (uint64_t)v; }
    ctx->debug = getenv("COVT_DEBUG") != nullptr;
    ctx->serial_classes = getenv("COVT_CONCURRENT") == nullptr;
    ctx->side_big = getenv("COVT_SERIAL_BIG") == nullptr;
    if (const char* ch = getenv("COVT_CHAINS")) { const int v = atoi(ch); if (v == 1 || v == 2 || v == 3 || v == 5) ctx->class_chains = v; }
    if ((e = cudaStreamCreateWithFlags(&ctx->big_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (auto& ev : ctx->ev_pass1) if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_big_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int c = 0; c < NUM_OP_CLASSES; c++) {
        if ((e = cudaStreamCreateWithFlags(&ctx->class_stream[c], cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_join[c], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if (const char* mt = getenv("COVT_SEG_MIN_TILES")) { const long long v = atoll(mt); if (v > 0) ctx->seg_min_tiles = (uint32_t)std::min<long long>(v, 1ll << 30); }
    if (const char* ms = getenv("COVT_MAX_SEGMENTS")) { const long long v = atoll(ms); if (v > 0) ctx->max_segments = (uint32_t)std::min<long long>(v, 4096); }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_totals), 64 * sizeof(uint64_t))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_seg), sizeof(SegState))) != cudaSuccess) return bail("cudaMallocHost", e);
    // keep freed result buffers in the stream-ordered pool: batches are decoded back to back
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return COVT_OK;
}

void covt_destroy(covt_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->dev_cache) cudaFreeAsync(b.first, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->h_seg) cudaFreeHost(ctx->h_seg);
    for (auto& b : ctx->pinned_cache) cudaFreeHost(b.first);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    if (ctx->big_stream) cudaStreamDestroy(ctx->big_stream);
    for (auto& e : ctx->ev_pass1) if (e) cudaEventDestroy(e);
    if (ctx->ev_big_done) cudaEventDestroy(ctx->ev_big_done);
    for (int c = 0; c < NUM_OP_CLASSES; c++) {
        if (ctx->class_stream[c]) cudaStreamDestroy(ctx->class_stream[c]);
        if (ctx->ev_join[c]) cudaEventDestroy(ctx->ev_join[c]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int32_t covt_last_error(covt_ctx* ctx, char* buf, size_t buf_len)
{
    if (!buf || !buf_len) return COVT_ERR_INVALID_ARG;
    std::string m;
    if (ctx) m = ctx->err;
    else { std::lock_guard<std::mutex> g(g_err_mutex); m = g_create_error; }
    snprintf(buf, buf_len, "%s", m.c_str());
    return COVT_OK;
}

int32_t covt_trim(covt_ctx* ctx)
{
    if (!ctx) return COVT_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    for (auto& b : ctx->dev_cache) cudaFreeAsync(b.first, ctx->stream);
    ctx->dev_cache.clear();
    CK(cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    return COVT_OK;
}

int32_t covt_host_register(covt_ctx* ctx, void* ptr, size_t bytes)
{
    if (!ctx || !ptr) return COVT_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return COVT_OK;
}
int32_t covt_host_unregister(covt_ctx* ctx, void* ptr)
{
    if (!ctx || !ptr) return COVT_ERR_INVALID_ARG;
    CK(cudaHostUnregister(ptr));
    return COVT_OK;
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
int32_t covt_batch_upload(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, covt_batch** out)
{
    if (!ctx || !out || (!blob && n_tiles) || !tile_offsets) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_batch_upload: null argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_tiles; i++)
        if (tile_offsets[i + 1] < tile_offsets[i]) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_batch_upload: tile_offsets must be non-decreasing");
    CK(cudaSetDevice(ctx->device));
    covt_batch* b = new covt_batch();
    b->ctx = ctx;
    b->n_tiles = n_tiles;
    b->blob_len = tile_offsets[n_tiles];
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    step(cudaEventCreate(&e0));
    step(cudaEventCreate(&e1));
    // 256 bytes of zero padding: kernels read whole 16-byte windows and one word past unaligned words
    step(dev_alloc_bytes(ctx, reinterpret_cast<void**>(&b->d_blob), b->blob_len + 256));
    step(dev_alloc(ctx, &b->d_tile_offsets, (uint64_t)n_tiles + 1));
    if (e == cudaSuccess) {
        step(cudaEventRecord(e0, ctx->stream));
        step(cudaMemsetAsync(b->d_blob + b->blob_len, 0, 256, ctx->stream));
        if (b->blob_len) step(cudaMemcpyAsync(b->d_blob, blob, b->blob_len, cudaMemcpyHostToDevice, ctx->stream));
        step(cudaMemcpyAsync(b->d_tile_offsets, tile_offsets, ((uint64_t)n_tiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        step(cudaEventRecord(e1, ctx->stream));
        step(cudaStreamSynchronize(ctx->stream));
    }
    if (e == cudaSuccess) cudaEventElapsedTime(&b->h2d_ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) {  // every error path gives the device blocks and the batch object back
        cudaStreamSynchronize(ctx->stream);
        covt_batch_free(b);
        CK(e);
    }
    *out = b;
    return COVT_OK;
}

void covt_batch_free(covt_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    dev_free(b->ctx, b->d_blob);
    dev_free(b->ctx, b->d_tile_offsets);
    delete b;
}

// ------------------------------------------------------------------------------------------------
// decode of a batch resident in (or on its way to) HBM, one or more segments of tiles
// ------------------------------------------------------------------------------------------------
// Property columns of a batch that is resident in HBM (COVT_FLAG_DECODE_PROPERTIES): CovtParser.decodePropertyColumn
// (CovtParser.java:276-390) for every property column of every tile, after the geometry path of the batch (kernels: covt_props.cuh).
// One pass over all tiles: count -> scan -> ONE size read-back -> records + tasks -> the codec-class kernels -> finish.
static int32_t decode_properties(covt_ctx* ctx, covt_batch* batch, uint32_t container, const covt_tilejson* tilejson, uint32_t flags, covt_result* R)
{
    const uint32_t n_layers = R->n_layers;
    if (!n_layers) return COVT_OK;
    cudaStream_t st = ctx->stream;
    Profiler prof = {ctx, (flags & COVT_FLAG_PROFILE_KERNELS) != 0, {}};
    uint64_t *d_pcols = nullptr, *d_block_sums = nullptr, *d_totals = nullptr;
    uint32_t *d_tj = nullptr, *d_counter = nullptr, *d_aux = nullptr, *d_queue = nullptr;
    DeviceTask* d_tasks = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int32_t rc = COVT_OK;
    auto cleanup = [&]() {
        dev_free(ctx, d_pcols); dev_free(ctx, d_block_sums); dev_free(ctx, d_totals); dev_free(ctx, d_tj); dev_free(ctx, d_counter);
        dev_free(ctx, d_aux); dev_free(ctx, d_queue); dev_free(ctx, d_tasks);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    };
#define CKP(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof(m_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = m_;                                                                            \
            rc = e_ == cudaErrorMemoryAllocation ? COVT_ERR_OOM : COVT_ERR_CUDA;                      \
            cudaStreamSynchronize(st);                                                                \
            cleanup();                                                                                \
            return rc;                                                                                \
        }                                                                                             \
    } while (0)
    const uint32_t nb = (n_layers + 255) / 256;
    CKP(cudaEventCreate(&ev0));
    CKP(cudaEventCreate(&ev1));
    CKP(dev_alloc(ctx, &d_pcols, (uint64_t)PROP_COLS * n_layers));
    CKP(dev_alloc(ctx, &d_block_sums, (uint64_t)PROP_COLS * nb));
    CKP(dev_alloc(ctx, &d_totals, PROP_COLS + 2));
    CKP(cudaMemsetAsync(d_totals, 0, (PROP_COLS + 2) * sizeof(uint64_t), st));
    CKP(dev_alloc(ctx, &d_counter, 16));
    uint32_t tj_layers = 0;
    if (tilejson && tilejson->n_vector_layers && tilejson->n_fields) {
        tj_layers = tilejson->n_vector_layers;
        CKP(dev_alloc(ctx, &d_tj, tj_layers));
        CKP(cudaMemcpyAsync(d_tj, tilejson->n_fields, tj_layers * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    CKP(cudaMemsetAsync(d_counter, 0, 16 * sizeof(uint32_t), st));
    CKP(cudaEventRecord(ev0, st));
    PropOut po = {};
    prof.begin("k0_props_scan", 0);
    CKP(launch_k0_props(false, batch->d_blob, batch->d_tile_offsets, R->d_layers, n_layers, container, d_tj, tj_layers, d_pcols, po, d_totals + PROP_COLS, st));
    CKP(launch_scan_tile_cols(d_pcols, n_layers, d_block_sums, d_totals, st, PROP_COLS));
    prof.end();
    uint64_t* h = ctx->h_totals;  // pinned scratch (64 words)
    CKP(cudaMemcpyAsync(h, d_totals, PROP_COLS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    const uint64_t n_cols = h[PROP_COL_COLUMNS], n_dicts = h[PROP_COL_DICTS];
    uint64_t n_tasks = 0, class_n[NUM_OP_CLASSES];
    for (int c = 0; c < NUM_OP_CLASSES; c++) { class_n[c] = h[PROP_COL_CLASS0 + c]; po.class_off[c] = n_tasks; n_tasks += class_n[c]; }
    if (n_cols * PROP_AUX_WORDS + n_dicts > 0xffffff00ull || n_tasks > 0xffffff00ull) { cleanup(); return fail(ctx, COVT_ERR_INVALID_ARG, "too many property columns in one batch"); }
    uint64_t arena_bytes = 0, buf_off[COVT_NUM_PROP_BUFFERS];
    for (int b = 0; b < COVT_NUM_PROP_BUFFERS; b++) {
        R->pcounts[b] = h[PROP_COL_BUF0 + b];
        buf_off[b] = arena_bytes;
        arena_bytes += (R->pcounts[b] * kPropBufElemSizeHost[b] + 64 + 255) & ~255ull;
    }
    R->n_prop_cols = (uint32_t)n_cols;
    R->n_prop_dicts = (uint32_t)n_dicts;
    CKP(dev_alloc_bytes(ctx, &R->prop_arena, arena_bytes + 256));
    for (int b = 0; b < COVT_NUM_PROP_BUFFERS; b++) {
        R->pbufs[b] = static_cast<uint8_t*>(R->prop_arena) + buf_off[b];
        po.buf[b] = R->pbufs[b];
    }
    CKP(dev_alloc(ctx, &R->d_prop_cols, n_cols));
    CKP(dev_alloc(ctx, &R->d_prop_dicts, n_dicts));
    CKP(dev_alloc(ctx, &d_aux, n_cols * PROP_AUX_WORDS + n_dicts));
    CKP(dev_alloc(ctx, &d_tasks, n_tasks));
    CKP(dev_alloc(ctx, &d_queue, n_tasks));
    po.cols = R->d_prop_cols;
    po.dicts = R->d_prop_dicts;
    po.aux = d_aux;
    po.aux_dict_base = n_cols * PROP_AUX_WORDS;
    po.tasks = d_tasks;
    prof.begin("k0_props_fill", 0);
    CKP(launch_k0_props(true, batch->d_blob, batch->d_tile_offsets, R->d_layers, n_layers, container, d_tj, tj_layers, d_pcols, po, d_totals + PROP_COLS, st));
    prof.end();
    for (int c = 0; c < NUM_OP_CLASSES; c++) {
        if (!class_n[c]) continue;
        prof.begin((std::string(op_class_name(c)) + "_props").c_str(), 0);
        CKP(launch_decode_class(c, batch->d_blob, d_tasks + po.class_off[c], (uint32_t)class_n[c], d_counter + 3 * c, d_queue + po.class_off[c], nullptr, d_aux,
                                ctx->sm_count, 0, st));
        prof.end();
    }
    prof.begin("k_prop_finish", 0);
    CKP(launch_prop_finish(batch->d_blob, (uint32_t)n_cols, (uint32_t)n_dicts, po, st));
    prof.end();
    CKP(cudaEventRecord(ev1, st));
    CKP(cudaMemcpyAsync(h, d_totals + PROP_COLS, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    R->timing.decode_ms += ms;
    R->timing.payload_bytes += h[0];  // property streams count like geometry streams: compressed bytes read
    R->timing.output_bytes += h[1];
    R->timing.kernel_launches += 7 + 2 * NUM_OP_CLASSES;
    prof.collect(R->kernel_times);
    cleanup();
#undef CKP
    return COVT_OK;
}

// seg_starts: S+1 tile indices, every segment non-empty. uploaded: nullptr (the whole blob is resident) or one event per
// segment, recorded on the copy stream after that segment's bytes and tile offsets arrived.
// With S > 1 the capacities of the result buffers are extrapolated from segment 0 (bytes ratio + 10 %); if a later segment
// does not fit, the device raises SegState::overflow, every later kernel becomes a no-op and the batch is decoded again as one
// segment with exact sizes (correct, only slower).
static int32_t decode_segments(covt_ctx* ctx, covt_batch* batch, uint32_t container, const covt_tilejson* tilejson, uint32_t flags,
                               const std::vector<uint32_t>& seg_starts, const uint64_t* h_tile_offsets,
                               const std::vector<cudaEvent_t>* uploaded, covt_result** out, const covt_host_sink* sink = nullptr)
{
    *out = nullptr;
    const uint32_t n_tiles = batch->n_tiles;
    const uint32_t S = (uint32_t)seg_starts.size() - 1;
    cudaStream_t st = ctx->stream;
    covt_result* R = new covt_result();
    R->ctx = ctx;
    R->n_tiles = n_tiles;
    Profiler prof = {ctx, (flags & COVT_FLAG_PROFILE_KERNELS) != 0, {}};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Host sink (covt_decode_batch_to_host): the slice of every wanted buffer that a segment produced goes to the caller's
    // page-locked memory on its own stream while later segments are uploaded and decoded. The host learns a segment's element
    // totals from a snapshot of SegState::base copied right behind the segment's kernels; it waits for segment i-1's snapshot
    // only after it has queued segment i, so the GPU never runs dry.
    uint64_t* h_snap = nullptr;                 // pinned: S x TILE_COLS running totals
    std::vector<cudaEvent_t> seg_done(sink ? seg_starts.size() - 1 : 0, nullptr);
    cudaEvent_t ev_out0 = nullptr, ev_out1 = nullptr;
    uint64_t sink_prev[COVT_NUM_BUFFERS] = {};  // elements of every buffer already sent
    bool sink_small = false;
    auto sink_cleanup = [&]() {
        for (auto e : seg_done) if (e) cudaEventDestroy(e);
        if (ev_out0) cudaEventDestroy(ev_out0);
        if (ev_out1) cudaEventDestroy(ev_out1);
        if (h_snap) pinned_give(ctx, h_snap, (seg_starts.size() - 1) * TILE_COLS * sizeof(uint64_t));
        h_snap = nullptr;
    };
    // sends what segment sg added to the wanted buffers (R->bufs are known from segment 0 on); returns the first CUDA error
    auto sink_send = [&](uint32_t sg) -> cudaError_t {
        cudaError_t e = cudaEventSynchronize(seg_done[sg]);
        if (e != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(ctx->out_stream, seg_done[sg], 0)) != cudaSuccess) return e;
        const uint64_t* base = h_snap + (uint64_t)sg * TILE_COLS;
        for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
            if (!sink->ptr[b]) continue;
            const uint64_t end = base[1 + b];
            if (end > sink->capacity[b]) { sink_small = true; continue; }
            if (end <= sink_prev[b]) continue;
            const uint64_t es = kBufElemSize[b];
            e = cudaMemcpyAsync(static_cast<uint8_t*>(sink->ptr[b]) + sink_prev[b] * es, static_cast<const uint8_t*>(R->bufs[b]) + sink_prev[b] * es,
                                (end - sink_prev[b]) * es, cudaMemcpyDeviceToHost, ctx->out_stream);
            if (e != cudaSuccess) return e;
            sink_prev[b] = end;
        }
        return cudaSuccess;
    };

    uint64_t *d_cols = nullptr, *d_block_sums = nullptr, *d_totals = nullptr;
    uint32_t *d_tj = nullptr, *d_counter = nullptr, *d_tile_err = nullptr;
    SegState* d_seg = nullptr;
    DeviceTask* d_tasks = nullptr;
    uint32_t* d_queue = nullptr;  // large streams handed from pass 1 to pass 2 of a codec class (same layout as d_tasks)
    uint32_t max_seg_tiles = 1;
    for (uint32_t i = 0; i < S; i++) max_seg_tiles = std::max(max_seg_tiles, seg_starts[i + 1] - seg_starts[i]);
    const uint32_t nb = (max_seg_tiles + 255) / 256;
    int32_t rc = COVT_OK;
    auto cleanup_tmp = [&]() {
        dev_free(ctx, d_cols);
        dev_free(ctx, d_block_sums);
        dev_free(ctx, d_totals);
        dev_free(ctx, d_tj);
        dev_free(ctx, d_counter);
        dev_free(ctx, d_tile_err);
        dev_free(ctx, d_tasks);
        dev_free(ctx, d_queue);
        dev_free(ctx, d_seg);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (sink && ctx->out_stream) cudaStreamSynchronize(ctx->out_stream);
        sink_cleanup();
    };
#define CKR(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof(m_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = m_;                                                                            \
            rc = e_ == cudaErrorMemoryAllocation ? COVT_ERR_OOM : COVT_ERR_CUDA;                      \
            cudaStreamSynchronize(st);                                                                \
            if (ctx->big_stream) cudaStreamSynchronize(ctx->big_stream);                              \
            cleanup_tmp();                                                                            \
            covt_result_free(R);                                                                      \
            return rc;                                                                                \
        }                                                                                             \
    } while (0)

    CKR(cudaEventCreate(&ev0));
    CKR(cudaEventCreate(&ev1));
    if (sink) {
        if (!ctx->out_stream) CKR(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
        CKR(pinned_take(ctx, reinterpret_cast<void**>(&h_snap), (uint64_t)S * TILE_COLS * sizeof(uint64_t)));
        for (auto& e : seg_done) CKR(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CKR(cudaEventCreate(&ev_out0));
        CKR(cudaEventCreate(&ev_out1));
        CKR(cudaEventRecord(ev_out0, ctx->out_stream));
    }
    CKR(dev_alloc(ctx, &d_cols, (uint64_t)TILE_COLS * max_seg_tiles));
    CKR(dev_alloc(ctx, &d_block_sums, (uint64_t)TILE_COLS * nb));
    CKR(dev_alloc(ctx, &d_totals, 32));
    CKR(dev_alloc(ctx, &d_counter, 16));
    CKR(dev_alloc(ctx, &d_seg, 1));
    CKR(dev_alloc(ctx, &R->d_tile_status, (uint64_t)n_tiles + 1));
    CKR(dev_alloc(ctx, &R->d_first_layer, (uint64_t)n_tiles + 2));
    CKR(dev_alloc(ctx, &d_tile_err, (uint64_t)n_tiles + 1));
    uint32_t tj_layers = 0;
    if (tilejson && tilejson->n_vector_layers && tilejson->n_fields) {
        tj_layers = tilejson->n_vector_layers;
        CKR(dev_alloc(ctx, &d_tj, tj_layers));
        CKR(cudaMemcpyAsync(d_tj, tilejson->n_fields, tj_layers * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    CKR(cudaMemsetAsync(d_totals, 0, 32 * sizeof(uint64_t), st));
    CKR(cudaMemsetAsync(d_counter, 0, 16 * sizeof(uint32_t), st));
    CKR(cudaMemsetAsync(d_seg, 0, sizeof(SegState), st));
    CKR(cudaMemsetAsync(R->d_first_layer, 0, ((uint64_t)n_tiles + 2) * sizeof(uint32_t), st));
    CKR(cudaMemsetAsync(d_tile_err, 0xff, ((uint64_t)n_tiles + 1) * sizeof(uint32_t), st));

    CKR(cudaEventRecord(ev0, st));
    uint32_t launches = 0;
    ResultBuffers rb = {};
    ClassOffsets class_off = {};
    uint64_t task_cap[NUM_OP_CLASSES] = {};
    uint64_t layers_per_seg_bound = 0;
    for (uint32_t sg = 0; sg < S && n_tiles; sg++) {
        const uint32_t t0 = seg_starts[sg], nt = seg_starts[sg + 1] - t0;
        if (uploaded) CKR(cudaStreamWaitEvent(st, (*uploaded)[sg], 0));
        // ---- K0 pass 1: layers per tile + slice sizes; column scan -> this segment's totals ----
        prof.begin("k0_scan_tiles", 0);
        CKR(launch_k0_scan_tiles(batch->d_blob, batch->d_tile_offsets + t0, nt, t0, container, d_tj, tj_layers, flags, d_cols, R->d_tile_status + t0, st));
        prof.end();
        prof.begin("scan_tile_cols", 0);
        CKR(launch_scan_tile_cols(d_cols, nt, d_block_sums, d_seg->seg_total, st));
        prof.end();
        launches += 4;
        if (sg == 0) {
            // the ONE size read-back of the call: segment 0's totals size (or, for several segments, extrapolate) everything
            CKR(cudaMemcpyAsync(ctx->h_totals, d_seg->seg_total, TILE_COLS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CKR(cudaStreamSynchronize(st));
            SegState* hs = ctx->h_seg;
            memset(hs, 0, sizeof(SegState));
            double scale = 1.0;
            if (S > 1) {
                const double seg_bytes = (double)(h_tile_offsets[seg_starts[1]] - h_tile_offsets[0]);
                scale = seg_bytes > 0 ? (double)batch->blob_len / seg_bytes * 1.10 : 1e9;
            }
            for (int c = 0; c < COL_CLASS0; c++)
                hs->cap[c] = S > 1 ? (uint64_t)((double)ctx->h_totals[c] * scale) + 65536 : ctx->h_totals[c];
            if (hs->cap[0] > 0xffffffffull / (sizeof(covt_layer) / 4) - 1)  // DeviceTask::ref = word index into the layer table must fit 32 bits
                { cleanup_tmp(); covt_result_free(R); return fail(ctx, COVT_ERR_INVALID_ARG, "too many layers in one batch"); }
            // the task lists are per segment: a segment that needs more entries than extrapolated counts as an overflow too
            const double seg_scale = S > 1 ? (double)max_seg_tiles / nt * 1.25 : 1.0;
            uint64_t task_total = 0;
            for (int c = 0; c < NUM_OP_CLASSES; c++) {
                task_cap[c] = S > 1 ? (uint64_t)((double)ctx->h_totals[COL_CLASS0 + c] * seg_scale) + 4096 : ctx->h_totals[COL_CLASS0 + c];
                hs->cap[COL_CLASS0 + c] = task_cap[c];
                class_off.off[c] = task_total;
                task_total += task_cap[c];
            }
            layers_per_seg_bound = S > 1 ? (uint64_t)((double)ctx->h_totals[0] * seg_scale) + 4096 : ctx->h_totals[0];
            uint64_t arena_bytes = 0, buf_off[COVT_NUM_BUFFERS];
            for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
                buf_off[b] = arena_bytes;
                arena_bytes += (hs->cap[1 + b] * kBufElemSize[b] + 64 + 255) & ~255ull;
            }
            const double t_alloc0 = now_ms();
            CKR(dev_alloc_bytes(ctx, &R->arena, arena_bytes + 256));
            for (int b = 0; b < COVT_NUM_BUFFERS; b++) {
                R->bufs[b] = hs->cap[1 + b] ? static_cast<uint8_t*>(R->arena) + buf_off[b] : nullptr;
                rb.ptr[b] = R->bufs[b];
            }
            if (ctx->debug) fprintf(stderr, "[covt] result arena %.2f GB allocated in %.2f ms (segments %u)\n", arena_bytes / 1e9, now_ms() - t_alloc0, S);
            CKR(dev_alloc(ctx, &R->d_layers, hs->cap[0]));
            CKR(dev_alloc(ctx, &d_tasks, task_total));
            CKR(dev_alloc(ctx, &d_queue, task_total));
            CKR(cudaMemcpyAsync(d_seg->cap, hs->cap, sizeof(hs->cap), cudaMemcpyHostToDevice, st));
        }
        CKR(launch_seg_begin(d_seg, d_counter, st));
        // ---- K0 pass 2: the layer table + eight decode tasks per layer ----
        prof.begin("k0_fill_layers", 0);
        CKR(launch_k0_fill_layers(batch->d_blob, batch->d_tile_offsets + t0, nt, t0, container, d_tj, tj_layers, flags, d_cols, rb, R->d_layers, d_tasks,
                                  class_off, R->d_first_layer + t0, d_seg, d_totals + 16, st));
        prof.end();
        // ---- every stream of every layer: one kernel per codec class ----
        // Small segments (the range of one GPU of eight, or one upload segment) gain 5 % from two chains; at 1 M tiles per call the gain is
        // within the run-to-run spread of the concurrent schedule (15.47 - 16.07 ms against 15.60 +- 0.01 ms serial), so large segments stay serial.
        const int chains = ctx->class_chains ? ctx->class_chains : (nt <= 262144u ? 2 : 1);
        if (!prof.on && ctx->serial_classes && chains > 1) {
            // The first passes of the five classes are independent. They run as a few CHAINS of kernels with full grids — chain 0 on the
            // main stream, the others beside it — so that the tail of a kernel (its last tickets) overlaps the start of a kernel of
            // another chain instead of leaving the GPU half empty (131 072 tiles: 2.42 -> 2.29 ms with two chains). Giving every
            // class a fixed SHARE of each SM (the round-1 experiment further down) was slower instead.
            static const int chain2[NUM_OP_CLASSES] = {/*BYTE_RLE*/ 0, /*RLE*/ 1, /*VARINT32*/ 0, /*VARINT64*/ 1, /*PFOR*/ 0};
            static const int chain3[NUM_OP_CLASSES] = {/*BYTE_RLE*/ 1, /*RLE*/ 2, /*VARINT32*/ 0, /*VARINT64*/ 2, /*PFOR*/ 1};
            static const int chain5[NUM_OP_CLASSES] = {1, 2, 0, 3, 4};
            const int* chain = chains == 2 ? chain2 : (chains == 3 ? chain3 : chain5);
            const int n_chains = chains == 2 ? 2 : (chains == 3 ? 3 : 5);
            CKR(cudaEventRecord(ctx->ev_fork, st));
            for (int k = 1; k < n_chains; k++) CKR(cudaStreamWaitEvent(ctx->class_stream[k], ctx->ev_fork, 0));
            // (launch order = class order; "longest kernels first" measured slower: 2.36 vs 2.29 ms per 131 072 tiles with two chains)
            for (int c = 0; c < NUM_OP_CLASSES; c++) {
                cudaStream_t sc = chain[c] == 0 ? st : ctx->class_stream[chain[c]];
                CKR(launch_decode_class(c, batch->d_blob, d_tasks + class_off.off[c], (uint32_t)std::min<uint64_t>(task_cap[c], 0xffffff00ull), d_counter + 3 * c,
                                        d_queue + class_off.off[c], d_seg, reinterpret_cast<uint32_t*>(R->d_layers), ctx->sm_count, 0, sc,
                                        ctx->big_stream, ctx->ev_pass1[c]));
            }
            for (int k = 1; k < n_chains; k++) {
                CKR(cudaEventRecord(ctx->ev_join[k], ctx->class_stream[k]));
                CKR(cudaStreamWaitEvent(st, ctx->ev_join[k], 0));
            }
            CKR(cudaEventRecord(ctx->ev_big_done, ctx->big_stream));
            CKR(cudaStreamWaitEvent(st, ctx->ev_big_done, 0));
        } else if (prof.on || ctx->serial_classes) {
            // (profiling brackets every class with events on the main stream: everything stays there)
            const bool side = ctx->side_big && !prof.on;
            for (int c = 0; c < NUM_OP_CLASSES; c++) {
                prof.begin(op_class_name(c), 0);
                CKR(launch_decode_class(c, batch->d_blob, d_tasks + class_off.off[c], (uint32_t)std::min<uint64_t>(task_cap[c], 0xffffff00ull), d_counter + 3 * c,
                                        d_queue + class_off.off[c], d_seg, reinterpret_cast<uint32_t*>(R->d_layers), ctx->sm_count, 0, st,
                                        side ? ctx->big_stream : nullptr, ctx->ev_pass1[c]));
                prof.end();
            }
            if (side) {  // the assembler reads what the second passes wrote
                CKR(cudaEventRecord(ctx->ev_big_done, ctx->big_stream));
                CKR(cudaStreamWaitEvent(st, ctx->ev_big_done, 0));
            }
        } else {
            // experiment (off by default, see covt_ctx::serial_classes): the five codec kernels are independent, run them
            // side by side
            static const int order[NUM_OP_CLASSES] = {CLASS_VARINT32, CLASS_PFOR, CLASS_RLE, CLASS_VARINT64, CLASS_BYTE_RLE};
            static const int share[NUM_OP_CLASSES] = {/*BYTE_RLE*/ 2, /*RLE*/ 3, /*VARINT32*/ 5, /*VARINT64*/ 1, /*PFOR*/ 2};
            CKR(cudaEventRecord(ctx->ev_fork, st));
            for (int i = 0; i < NUM_OP_CLASSES; i++) {
                const int c = order[i];
                CKR(cudaStreamWaitEvent(ctx->class_stream[c], ctx->ev_fork, 0));
                CKR(launch_decode_class(c, batch->d_blob, d_tasks + class_off.off[c], (uint32_t)std::min<uint64_t>(task_cap[c], 0xffffff00ull), d_counter + 3 * c,
                                        d_queue + class_off.off[c], d_seg, reinterpret_cast<uint32_t*>(R->d_layers), ctx->sm_count, share[c], ctx->class_stream[c]));
                CKR(cudaEventRecord(ctx->ev_join[c], ctx->class_stream[c]));
                CKR(cudaStreamWaitEvent(st, ctx->ev_join[c], 0));
            }
        }
        // ---- geometry assembly ----
        prof.begin("k_assemble_layers", 0);
        CKR(launch_assemble_layers(R->d_layers, (uint32_t)std::min<uint64_t>(layers_per_seg_bound, 0xffffff00ull), rb, flags, d_counter + 15, d_seg, d_totals + 16, d_tile_err, ctx->sm_count, st));
        prof.end();
        CKR(launch_seg_end(d_seg, sg + 1 == S ? R->d_first_layer + n_tiles : nullptr, st));
        launches += 13;  // k_seg_begin, k0_fill_layers, 5 codec kernels + 4 second passes, k_assemble_layers, k_seg_end
        if (sink) {
            CKR(cudaMemcpyAsync(h_snap + (uint64_t)sg * TILE_COLS, d_seg->base, TILE_COLS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CKR(cudaEventRecord(seg_done[sg], st));
            if (sg > 0) CKR(sink_send(sg - 1));  // (segment sg is queued: the GPU stays busy while the host waits for sg - 1)
        }
    }
    if (sink && S && n_tiles) CKR(sink_send(S - 1));
    if (n_tiles) {
        prof.begin("k_tile_status", 0);
        CKR(launch_finalize(R->d_layers, d_tile_err, n_tiles, (uint32_t)std::min<uint64_t>(ctx->h_seg->cap[0], 0xffffff00ull), flags, R->d_tile_status, d_totals + 16, d_seg, st));
        prof.end();
        launches += prof.on ? 3 : 2;  // k_tile_status, k_layer_totals (+ k_alg_bytes when profiling)
    }
    CKR(cudaEventRecord(ev1, st));
    CKR(cudaMemcpyAsync(ctx->h_totals + 32, d_totals + 16, FINAL_TOTALS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CKR(cudaMemcpyAsync(ctx->h_seg, d_seg, sizeof(SegState), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    if (ctx->h_seg->overflow) {
        // an extrapolated capacity was too small: decode again as ONE segment with exact sizes
        if (uploaded) cudaStreamSynchronize(ctx->copy_stream);
        prof.collect(R->kernel_times);
        cleanup_tmp();
        covt_result_free(R);
        if (S == 1) return fail(ctx, COVT_ERR_CUDA, "internal error: exact capacities overflowed");
        ctx->overflow_retries++;
        std::vector<uint32_t> one = {0u, n_tiles};
        rc = decode_segments(ctx, batch, container, tilejson, flags, one, h_tile_offsets, nullptr, out, sink);
        if (rc == COVT_OK) (*out)->timing.capacity_retries = 1;
        return rc;
    }
    cudaEventElapsedTime(&R->timing.decode_ms, ev0, ev1);
    if (sink) {
        CKR(cudaEventRecord(ev_out1, ctx->out_stream));
        CKR(cudaStreamSynchronize(ctx->out_stream));
        cudaEventElapsedTime(&R->timing.d2h_ms, ev_out0, ev_out1);
        if (sink_small) {
            char m_[256];
            uint64_t need[COVT_NUM_BUFFERS];
            for (int b = 0; b < COVT_NUM_BUFFERS; b++) need[b] = ctx->h_seg->base[1 + b];
            snprintf(m_, sizeof(m_), "covt_decode_batch_to_host: a sink buffer is too small (coords need %llu elements, ids %llu)",
                     (unsigned long long)need[COVT_BUF_A_COORDS], (unsigned long long)need[COVT_BUF_S_IDS]);
            cleanup_tmp();
            covt_result_free(R);
            return fail(ctx, COVT_ERR_INVALID_ARG, m_);
        }
    }
    R->n_layers = (uint32_t)ctx->h_seg->base[0];
    for (int b = 0; b < COVT_NUM_BUFFERS; b++) R->counts[b] = ctx->h_seg->base[1 + b];
    R->timing.h2d_ms = batch->h2d_ms;
    R->timing.vertices = ctx->h_totals[32];
    R->timing.payload_bytes = ctx->h_totals[33];
    R->timing.output_bytes = ctx->h_totals[34];
    if (prof.on) {
        // algorithmic bytes per kernel: known only now that k_alg_bytes has summed them on the device;
        // booked on the first launch of each kernel (the records of one kernel are merged by name)
        const uint64_t meta_bytes = batch->blob_len > R->timing.payload_bytes ? batch->blob_len - R->timing.payload_bytes : 0;
        uint64_t n_task_entries = 0;
        for (int c = 0; c < NUM_OP_CLASSES; c++) n_task_entries += ctx->h_seg->base[COL_CLASS0 + c];
        std::vector<std::string> seen;
        for (auto& r : prof.recs) {
            if (std::find(seen.begin(), seen.end(), r.name) != seen.end()) continue;
            seen.push_back(r.name);
            if (r.name == "k0_scan_tiles") r.alg_bytes = meta_bytes + (uint64_t)n_tiles * (8 + TILE_COLS * 8 + 4);
            else if (r.name == "scan_tile_cols") r.alg_bytes = 2ull * n_tiles * TILE_COLS * 8;
            else if (r.name == "k0_fill_layers") r.alg_bytes = meta_bytes + (uint64_t)n_tiles * (8 + TILE_COLS * 8) + (uint64_t)R->n_layers * sizeof(covt_layer) + n_task_entries * sizeof(DeviceTask);
            else if (r.name == "k_assemble_layers") r.alg_bytes = ctx->h_totals[32 + 8];
            else if (r.name == "k_tile_status") r.alg_bytes = (uint64_t)n_tiles * 12;
            else for (int c = 0; c < NUM_OP_CLASSES; c++) if (r.name == op_class_name(c)) r.alg_bytes = ctx->h_totals[32 + 3 + c];
        }
        prof.collect(R->kernel_times);
    }
    R->timing.kernel_launches = launches;
    R->timing.segments = S;
    cleanup_tmp();
#undef CKR
    if ((flags & COVT_FLAG_DECODE_PROPERTIES) && n_tiles) {
        rc = decode_properties(ctx, batch, container, tilejson, flags, R);
        if (rc != COVT_OK) { covt_result_free(R); return rc; }
    }
    *out = R;
    return COVT_OK;
}

int32_t covt_batch_decode(covt_ctx* ctx, covt_batch* batch, uint32_t container, const covt_tilejson* tilejson, uint32_t flags,
                          covt_result** out)
{
    if (!ctx || !batch || !out || batch->ctx != ctx) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_batch_decode: bad argument");
    if (container != COVT_CONTAINER_GEN2B && container != COVT_CONTAINER_GEN3) return fail(ctx, COVT_ERR_INVALID_ARG, "unknown container kind");
    CK(cudaSetDevice(ctx->device));
    std::vector<uint32_t> one = {0u, batch->n_tiles};
    return decode_segments(ctx, batch, container, tilejson, flags, one, nullptr, nullptr, out);
}

// Host input: the blob goes up in segments on the copy stream while earlier segments are being decoded.
static int32_t decode_batch_host(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                 const covt_tilejson* tilejson, uint32_t flags, const covt_host_sink* sink, covt_result** out);
int32_t covt_decode_batch(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                          const covt_tilejson* tilejson, uint32_t flags, covt_result** out)
{
    return decode_batch_host(ctx, blob, tile_offsets, n_tiles, container, tilejson, flags, nullptr, out);
}
int32_t covt_decode_batch_to_host(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                  const covt_tilejson* tilejson, uint32_t flags, const covt_host_sink* sink, covt_result** out)
{
    if (!sink) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_decode_batch_to_host: null sink");
    return decode_batch_host(ctx, blob, tile_offsets, n_tiles, container, tilejson, flags, sink, out);
}
static int32_t decode_batch_host(covt_ctx* ctx, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                 const covt_tilejson* tilejson, uint32_t flags, const covt_host_sink* sink, covt_result** out)
{
    if (!ctx || !out || (!blob && n_tiles) || !tile_offsets) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_decode_batch: null argument");
    if (container != COVT_CONTAINER_GEN2B && container != COVT_CONTAINER_GEN3) return fail(ctx, COVT_ERR_INVALID_ARG, "unknown container kind");
    *out = nullptr;
    // (tile_offsets must be non-decreasing: checked segment by segment right before a segment's bytes are queued, so that the first
    // copy starts after 1/S of that pass instead of after all of it — two linear passes over 1 M offsets cost ~2 ms of a 53 ms call)
    auto offsets_ok = [&](uint32_t t0, uint32_t t1) {
        for (uint32_t i = t0; i < t1; i++)
            if (tile_offsets[i + 1] < tile_offsets[i]) return false;
        return true;
    };
    const uint64_t blob_len = tile_offsets[n_tiles];
    if (blob_len < tile_offsets[0]) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_decode_batch: tile_offsets must be non-decreasing");
    // segments of ~SEG_BYTES, balanced by payload bytes, every one non-empty
    // Every segment costs ~1 ms of fixed kernel time (13 launches that are latency-bound at small sizes), and a segment must hold
    // enough tiles to fill the GPU (a kernel cannot finish before its largest stream, which one warp decodes), so use few, large
    // segments: at most 24, at least seg_bytes and 32768 tiles each. Measured on B200: 1 M tiles / 2.77 GB — 1 segment 76.5 ms,
    // 8 segments 54.4 ms, 24 segments 53.2 ms per call (the host->device copy alone takes 49.8 ms); the 91 OMT fixture tiles x256
    // (23 296 tiles of 122 KB, layers of up to 60 000 features) — 8 segments 199 ms, 1 segment ~90 ms.
    const uint64_t SEG_BYTES = std::max<uint64_t>(ctx->seg_bytes, 4096);
    uint32_t want = (uint32_t)std::min<uint64_t>(ctx->max_segments, std::max<uint64_t>(1, (blob_len - tile_offsets[0]) / SEG_BYTES));
    // ... or at least 1 GiB of large tiles: those fill the GPU through their large streams (fixture sweep, 2 segments: 75.5 -> 68.3 ms per call)
    want = std::max<uint32_t>(1, std::min<uint32_t>(want, std::max<uint32_t>(n_tiles / ctx->seg_min_tiles, (uint32_t)((blob_len - tile_offsets[0]) >> 30))));
    if (tile_offsets[0] != 0) want = 1;  // the segment extrapolation assumes the blob starts at its first tile
    std::vector<uint32_t> starts(want + 1);
    covt_partition_tiles(tile_offsets, n_tiles, want, starts.data());
    starts.erase(std::unique(starts.begin(), starts.end()), starts.end());
    if (starts.size() < 2) starts = {0u, n_tiles};
    const uint32_t S = (uint32_t)starts.size() - 1;
    if (S == 1) {
        if (!offsets_ok(0, n_tiles)) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_decode_batch: tile_offsets must be non-decreasing");
        covt_batch* b = nullptr;
        int32_t rc = covt_batch_upload(ctx, blob, tile_offsets, n_tiles, &b);
        if (rc != COVT_OK) return rc;
        std::vector<uint32_t> one = {0u, n_tiles};
        rc = decode_segments(ctx, b, container, tilejson, flags, one, nullptr, nullptr, out, sink);
        covt_batch_free(b);
        return rc;
    }
    CK(cudaSetDevice(ctx->device));
    const double t_call0 = now_ms();
    covt_batch* b = new covt_batch();
    b->ctx = ctx;
    b->n_tiles = n_tiles;
    b->blob_len = blob_len;
    std::vector<cudaEvent_t> evs(S, nullptr);
    cudaEvent_t e_alloc = nullptr, e0 = nullptr, e1 = nullptr;
    auto destroy_events = [&]() {
        for (auto e : evs) if (e) cudaEventDestroy(e);
        if (e_alloc) cudaEventDestroy(e_alloc);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    step(dev_alloc_bytes(ctx, reinterpret_cast<void**>(&b->d_blob), blob_len + 256));
    step(dev_alloc(ctx, &b->d_tile_offsets, (uint64_t)n_tiles + 1));
    step(cudaEventCreateWithFlags(&e_alloc, cudaEventDisableTiming));
    step(cudaEventCreate(&e0));
    step(cudaEventCreate(&e1));
    step(cudaEventRecord(e_alloc, ctx->stream));
    step(cudaStreamWaitEvent(ctx->copy_stream, e_alloc, 0));
    step(cudaEventRecord(e0, ctx->copy_stream));
    step(cudaMemsetAsync(b->d_blob + blob_len, 0, 256, ctx->copy_stream));
    // ALL tile offsets first, in one copy (8 bytes per tile): a caller's offsets are often in pageable memory, and a pageable copy
    // queued between the blob segments holds the host back until the segments before it have gone up
    step(cudaMemcpyAsync(b->d_tile_offsets, tile_offsets, ((uint64_t)n_tiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    bool bad_offsets = false;
    for (uint32_t sg = 0; sg < S && e == cudaSuccess; sg++) {
        const uint32_t t0 = starts[sg], t1 = starts[sg + 1];
        const uint64_t o0 = tile_offsets[t0], o1 = tile_offsets[t1];
        if (t1 < t0 || o1 < o0 || o1 > blob_len || !offsets_ok(t0, t1)) { bad_offsets = true; break; }
        if (o1 > o0) step(cudaMemcpyAsync(b->d_blob + o0, blob + o0, o1 - o0, cudaMemcpyHostToDevice, ctx->copy_stream));
        step(cudaEventCreateWithFlags(&evs[sg], cudaEventDisableTiming));
        step(cudaEventRecord(evs[sg], ctx->copy_stream));
    }
    step(cudaEventRecord(e1, ctx->copy_stream));
    if (e != cudaSuccess || bad_offsets) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        destroy_events();
        covt_batch_free(b);
        if (bad_offsets) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_decode_batch: tile_offsets must be non-decreasing");
        CK(e);
    }
    const double t_enq = now_ms();
    int32_t rc = decode_segments(ctx, b, container, tilejson, flags, starts, tile_offsets, &evs, out, sink);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->debug) fprintf(stderr, "[covt] covt_decode_batch: enqueue uploads %.2f ms, decode_segments %.2f ms\n", t_enq - t_call0, now_ms() - t_enq);
    if (rc == COVT_OK) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) (*out)->timing.h2d_ms = ms;
    }
    destroy_events();
    covt_batch_free(b);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// stream path
// ------------------------------------------------------------------------------------------------
static uint32_t op_elem_size(uint32_t op)
{
    switch (op) {
    case COVT_OP_BYTE_RLE: return 1;
    case COVT_OP_RLE_U64: case COVT_OP_RLE_S64: case COVT_OP_VARINT_U64: case COVT_OP_VARINT_ZZ_DELTA_64:
    case COVT_OP_VARINT_U32_AS_I64: case COVT_OP_VARINT_ZZ_DELTA_AS_I64: case COVT_OP_VARINT_ZZ_AS_I64: return 8;
    default: return 4;
    }
}
static int op_big_post(uint32_t op)
{
    switch (op) {
    case COVT_OP_VARINT_U32: return POST_PLAIN;
    case COVT_OP_VARINT_ZZ: return POST_ZZ;
    case COVT_OP_VARINT_ZZ_DELTA: return POST_ZZ_DELTA;
    case COVT_OP_VARINT_ZZ_DELTA_XY: return POST_ZZ_DELTA_XY;
    case COVT_OP_VARINT_DELTA_MORTON: return POST_DELTA_MORTON;
    default: return -1;
    }
}

int32_t covt_resolve_op(uint32_t stream_type, uint32_t encoding, uint32_t column_type, uint32_t flags)
{
    return (int32_t)host_resolve_op(stream_type, encoding, column_type, flags);
}

int32_t covt_batch_decode_streams(covt_ctx* ctx, covt_batch* batch, covt_stream_desc* descs, uint32_t n, uint32_t flags, covt_result** out)
{
    if (!ctx || !batch || !out || (!descs && n) || batch->ctx != ctx) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_batch_decode_streams: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    covt_result* R = new covt_result();
    R->ctx = ctx;
    Profiler prof = {ctx, (flags & COVT_FLAG_PROFILE_KERNELS) != 0, {}};
    std::vector<DeviceTask> tasks(n);
    std::vector<BigStream> bigs;
    uint32_t n_big_chunks = 0;
    std::vector<uint32_t> big_task;
    uint64_t arena = 0, payload = 0, out_bytes = 0, big_alg = 0;
    uint64_t class_alg[NUM_OP_CLASSES] = {};
    const uint64_t BIG_BYTES = 1u << 18;  // streams of >= 256 KiB take the multi-CTA look-back kernel
    for (uint32_t i = 0; i < n; i++) {
        covt_stream_desc& d = descs[i];
        DeviceTask& t = tasks[i];
        memset(&t, 0, sizeof(t));
        uint32_t op = d.op ? d.op : host_resolve_op(d.stream_type, d.encoding, d.column_type, flags);
        d.status = COVT_OK;
        d.bytes_consumed = 0;
        d.out_count = 0;
        d.out_offset = arena;
        if (op == COVT_OP_NONE || op >= COVT_NUM_OPS) { d.status = COVT_ERR_UNSUPPORTED_ENCODING; t.status = d.status; continue; }
        // (overflow-safe: the descriptors come from untrusted callers)
        if (d.byte_offset > batch->blob_len || d.byte_length > batch->blob_len - d.byte_offset) { d.status = COVT_ERR_TRUNCATED; t.status = d.status; continue; }
        // the same policy as the container walk: a request for more than 256 values per payload byte cannot decode with any codec of
        // the path, and must not be able to reserve gigabytes of output
        if ((uint64_t)d.num_values > 256ull * ((uint64_t)d.byte_length + 16ull)) { d.status = COVT_ERR_TRUNCATED; t.status = d.status; continue; }
        const bool morton = op == COVT_OP_VARINT_DELTA_MORTON || op == COVT_OP_PFOR_DELTA_MORTON;
        const uint64_t cnt = morton ? 2ull * d.num_values : d.num_values;
        d.out_count = cnt;
        t.src_offset = d.byte_offset;
        t.dst = reinterpret_cast<uint8_t*>(arena);  // offset for now, rebased once the arena is allocated
        t.byte_length = d.byte_length;
        t.num_values = d.num_values;
        t.op = (uint8_t)op;
        t.num_bits = d.num_bits;
        t.no_shift = (flags & COVT_FLAG_MORTON_NO_SHIFT) ? 1 : 0;
        // FastPFOR calls carry an exact byteLength (DecodingUtils.java:316); varint/RLE calls only a start position
        t.exact_length = (op == COVT_OP_PFOR_ZZ_DELTA || op == COVT_OP_PFOR_ZZ_DELTA_XY || op == COVT_OP_PFOR_DELTA_MORTON) ? 1 : 0;
        const uint64_t ob = cnt * op_elem_size(op);
        arena += (ob + 15) & ~15ull;
        payload += d.byte_length;
        out_bytes += ob;
        const int post = op_big_post(op);
        if (post >= 0 && d.byte_length >= BIG_BYTES && !((op == COVT_OP_VARINT_ZZ_DELTA_XY) && (d.num_values & 1u))) {
            // large varint stream: byte_length is taken as exact (documented in DESIGN.md)
            BigStream b;
            memset(&b, 0, sizeof(b));
            b.src_offset = d.byte_offset;
            b.byte_length = d.byte_length;
            b.num_values = d.num_values;
            b.first_chunk = n_big_chunks;
            const uint64_t window = (d.byte_offset & 15) + d.byte_length;  // blob base is 256-byte aligned
            b.n_chunks = (uint32_t)((window + K1_SC_BYTES - 1) / K1_SC_BYTES);
            n_big_chunks += b.n_chunks;
            b.post = (uint8_t)post;
            b.num_bits = d.num_bits;
            b.no_shift = t.no_shift;
            bigs.push_back(b);
            big_task.push_back(i);
            t.op = COVT_OP_NONE;  // the warp-per-stream kernel skips it
            big_alg += d.byte_length + ob;
        } else {
            int c = -1;
            switch (op) {
            case COVT_OP_BYTE_RLE: c = CLASS_BYTE_RLE; break;
            case COVT_OP_RLE_U32: case COVT_OP_RLE_U64: case COVT_OP_RLE_S64: c = CLASS_RLE; break;
            case COVT_OP_VARINT_U64: case COVT_OP_VARINT_ZZ_DELTA_64: c = CLASS_VARINT64; break;
            case COVT_OP_PFOR_ZZ_DELTA: case COVT_OP_PFOR_ZZ_DELTA_XY: case COVT_OP_PFOR_DELTA_MORTON: c = CLASS_PFOR; break;
            default: c = CLASS_VARINT32; break;
            }
            class_alg[c] += d.byte_length + ob;
        }
    }
    DeviceTask* d_tasks = nullptr;
    BigStream* d_bigs = nullptr;
    ChunkState* d_block_states = nullptr;
    ChunkState* d_states = nullptr;
    uint32_t* d_counter = nullptr;
    uint32_t* d_queue = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int32_t rc = COVT_OK;
    auto cleanup_tmp = [&]() {
        dev_free(ctx, d_tasks);
        dev_free(ctx, d_bigs);
        dev_free(ctx, d_block_states);
        dev_free(ctx, d_states);
        dev_free(ctx, d_counter);
        dev_free(ctx, d_queue);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    };
#define CKR(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof(m_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = m_;                                                                            \
            rc = e_ == cudaErrorMemoryAllocation ? COVT_ERR_OOM : COVT_ERR_CUDA;                      \
            cleanup_tmp();                                                                            \
            covt_result_free(R);                                                                      \
            return rc;                                                                                \
        }                                                                                             \
    } while (0)
    CKR(cudaEventCreate(&ev0));
    CKR(cudaEventCreate(&ev1));
    R->counts[COVT_BUF_STREAM_ARENA] = arena;
    CKR(dev_alloc_bytes(ctx, &R->arena, arena + 64));
    R->bufs[COVT_BUF_STREAM_ARENA] = R->arena;
    CKR(dev_alloc(ctx, &d_tasks, n));
    CKR(dev_alloc(ctx, &d_queue, n));
    CKR(dev_alloc(ctx, &d_counter, 16));
    // dense task list per codec class (the class kernels walk their own list); tasks taken by the large-stream kernel and
    // rejected requests go last. ref = index of the covt_stream_desc the outcome belongs to.
    std::vector<DeviceTask> sorted(n);
    uint32_t class_count[NUM_OP_CLASSES] = {}, class_first[NUM_OP_CLASSES + 1] = {};
    for (uint32_t i = 0; i < n; i++) {
        tasks[i].dst = reinterpret_cast<uint8_t*>(R->bufs[COVT_BUF_STREAM_ARENA]) + reinterpret_cast<uintptr_t>(tasks[i].dst);
        tasks[i].ref = i;
        const int c = host_op_class_of(tasks[i].op);
        if (c >= 0) class_count[c]++;
    }
    for (int c = 0; c < NUM_OP_CLASSES; c++) class_first[c + 1] = class_first[c] + class_count[c];
    std::vector<uint32_t> pos_of(n);
    {
        uint32_t cursor[NUM_OP_CLASSES + 1];
        for (int c = 0; c <= NUM_OP_CLASSES; c++) cursor[c] = class_first[c];
        for (uint32_t i = 0; i < n; i++) {
            const int c = host_op_class_of(tasks[i].op);
            pos_of[i] = cursor[c >= 0 ? c : NUM_OP_CLASSES]++;
            sorted[pos_of[i]] = tasks[i];
        }
    }
    if (!bigs.empty()) {
        CKR(dev_alloc(ctx, &d_bigs, bigs.size()));
        CKR(dev_alloc(ctx, &d_block_states, (uint64_t)n_big_chunks / K1_SCAN_BLOCK + 2));
        CKR(dev_alloc(ctx, &d_states, n_big_chunks));
        for (size_t k = 0; k < bigs.size(); k++) {
            const uint32_t i = big_task[k];
            bigs[k].dst = tasks[i].dst;
            bigs[k].status_out = &d_tasks[pos_of[i]].status;
            bigs[k].consumed_out = &d_tasks[pos_of[i]].consumed;
        }
        CKR(cudaMemcpyAsync(d_bigs, bigs.data(), bigs.size() * sizeof(BigStream), cudaMemcpyHostToDevice, st));
    }
    if (n) CKR(cudaMemcpyAsync(d_tasks, sorted.data(), (uint64_t)n * sizeof(DeviceTask), cudaMemcpyHostToDevice, st));
    CKR(cudaMemsetAsync(d_counter, 0, 16 * sizeof(uint32_t), st));
    CKR(cudaEventRecord(ev0, st));
    uint32_t launches = 0;
    if (!bigs.empty()) {
        prof.begin("k1_varint_stream", big_alg);
        CKR(launch_k1_varint_stream(batch->d_blob, d_bigs, (uint32_t)bigs.size(), n_big_chunks, d_states, d_block_states, st));
        prof.end();
        launches += 5;
    }
    for (int c = 0; c < NUM_OP_CLASSES; c++) {
        if (!class_count[c]) continue;
        prof.begin(op_class_name(c), class_alg[c]);
        CKR(launch_decode_class(c, batch->d_blob, d_tasks + class_first[c], class_count[c], d_counter + 3 * c, d_queue + class_first[c], nullptr, nullptr,
                                ctx->sm_count, 0, st));
        prof.end();
        launches += c == CLASS_VARINT32 ? 1 : 2;  // pass 1 (+ pass 2 over the queued large streams)
    }
    CKR(cudaEventRecord(ev1, st));
    if (n) CKR(cudaMemcpyAsync(sorted.data(), d_tasks, (uint64_t)n * sizeof(DeviceTask), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&R->timing.decode_ms, ev0, ev1);
    for (uint32_t i = 0; i < n; i++) {
        if (descs[i].status != COVT_OK) { descs[i].out_count = 0; continue; }
        const DeviceTask& t = sorted[pos_of[i]];
        descs[i].status = t.status;
        descs[i].bytes_consumed = t.consumed;
        if (t.status != COVT_OK && t.status != COVT_ERR_VARINT_OVERLONG) descs[i].out_count = 0;
    }
    R->timing.h2d_ms = batch->h2d_ms;
    R->timing.payload_bytes = payload;
    R->timing.output_bytes = out_bytes;
    R->timing.kernel_launches = launches;
    if (prof.on) prof.collect(R->kernel_times);
    cleanup_tmp();
#undef CKR
    *out = R;
    return COVT_OK;
}

int32_t covt_decode_streams(covt_ctx* ctx, const uint8_t* blob, uint64_t blob_len, covt_stream_desc* descs, uint32_t n_streams,
                            uint32_t flags, covt_result** out)
{
    uint64_t offs[2] = {0, blob_len};
    covt_batch* b = nullptr;
    int32_t rc = covt_batch_upload(ctx, blob, offs, 1, &b);
    if (rc != COVT_OK) return rc;
    rc = covt_batch_decode_streams(ctx, b, descs, n_streams, flags, out);
    covt_batch_free(b);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// stream encoders: the inverse of the decode ops, EncodingUtils.java:39-230 on the GPU (covt_encode.cu)
// ------------------------------------------------------------------------------------------------
int32_t covt_encode_streams(covt_ctx* ctx, const void* values, uint64_t values_bytes, covt_encode_desc* descs, uint32_t n, uint32_t flags,
                            covt_result** out)
{
    if (!ctx || !out || (!descs && n) || (!values && values_bytes)) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_encode_streams: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<EncPiece> pieces[3];  // 0 = varint pieces, 1 = RLE / Byte-RLE streams, 2 = FastPFOR streams
    uint64_t scratch[3] = {0, 0, 0}, in_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        covt_encode_desc& d = descs[i];
        d.status = COVT_OK;
        d.out_offset = 0;
        d.byte_length = 0;
        const uint32_t op = d.op;
        int kind = -1;
        uint64_t elem = 4, cap = 0;
        const uint64_t nv = d.num_values;
        switch (op) {
        case COVT_OP_VARINT_U32: case COVT_OP_VARINT_ZZ: case COVT_OP_VARINT_ZZ_DELTA: case COVT_OP_VARINT_ZZ_DELTA_XY: kind = 0; elem = 4; cap = 5; break;
        case COVT_OP_VARINT_DELTA_MORTON: kind = 0; elem = 8; cap = 10; break;
        case COVT_OP_VARINT_U64: case COVT_OP_VARINT_ZZ_DELTA_64: kind = 0; elem = 8; cap = 10; break;
        case COVT_OP_BYTE_RLE: kind = 1; elem = 1; cap = nv + nv / 128 + 8; break;
        case COVT_OP_RLE_U32: kind = 1; elem = 4; cap = 10 * nv + nv / 128 + 16; break;
        case COVT_OP_RLE_U64: case COVT_OP_RLE_S64: kind = 1; elem = 8; cap = 10 * nv + nv / 128 + 16; break;
        case COVT_OP_PFOR_ZZ_DELTA: case COVT_OP_PFOR_ZZ_DELTA_XY: kind = 2; elem = 4; break;
        case COVT_OP_PFOR_DELTA_MORTON: kind = 2; elem = 8; break;
        default: break;
        }
        if (kind < 0) { d.status = COVT_ERR_UNSUPPORTED_ENCODING; continue; }
        if (d.value_offset > values_bytes || nv * elem > values_bytes - d.value_offset || (d.value_offset % elem) != 0 ||
            (op == COVT_OP_VARINT_ZZ_DELTA_XY || op == COVT_OP_PFOR_ZZ_DELTA_XY) && (nv & 1u)) {
            d.status = COVT_ERR_INVALID_ARG;
            continue;
        }
        in_bytes += nv * elem;
        EncPiece p;
        memset(&p, 0, sizeof(p));
        p.stream_values = d.value_offset;
        p.stream = i;
        p.op = (uint8_t)op;
        p.num_bits = d.num_bits;
        if (kind == 0) {
            for (uint64_t f = 0; f < nv || f == 0; f += ENC_VARINT_PIECE) {
                p.first_index = (uint32_t)f;
                p.num_values = (uint32_t)std::min<uint64_t>(ENC_VARINT_PIECE, nv - f);
                p.scratch = reinterpret_cast<uint8_t*>(scratch[0]);  // offset for now
                scratch[0] += ((uint64_t)p.num_values * cap + 64 + 15) & ~15ull;
                pieces[0].push_back(p);
                if (nv == 0) break;
            }
        } else {
            if (kind == 2) cap = 16 * nv + 1024 * (nv / 65536 + 1) + 2048;  // words <= ~2.3 n + per-page headers (see k_enc_pfor)
            p.num_values = d.num_values;
            p.scratch = reinterpret_cast<uint8_t*>(scratch[kind]);
            scratch[kind] += (cap + 64 + 15) & ~15ull;
            pieces[kind].push_back(p);
        }
    }
    const uint32_t np[3] = {(uint32_t)pieces[0].size(), (uint32_t)pieces[1].size(), (uint32_t)pieces[2].size()};
    const uint32_t n_pieces = np[0] + np[1] + np[2];
    covt_result* R = new covt_result();
    R->ctx = ctx;
    uint8_t* d_values = nullptr;
    uint8_t* d_scratch = nullptr;
    EncPiece* d_pieces = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    int32_t rc = COVT_OK;
    auto cleanup_tmp = [&]() {
        dev_free(ctx, d_values);
        dev_free(ctx, d_scratch);
        dev_free(ctx, d_pieces);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
    };
#define CKE(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char m_[512];                                                                             \
            snprintf(m_, sizeof(m_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = m_;                                                                            \
            rc = e_ == cudaErrorMemoryAllocation ? COVT_ERR_OOM : COVT_ERR_CUDA;                      \
            cudaStreamSynchronize(st);                                                                \
            cleanup_tmp();                                                                            \
            covt_result_free(R);                                                                      \
            return rc;                                                                                \
        }                                                                                             \
    } while (0)
    CKE(cudaEventCreate(&ev0));
    CKE(cudaEventCreate(&ev1));
    CKE(cudaEventCreate(&ev2));
    const uint64_t scratch_total = scratch[0] + scratch[1] + scratch[2];
    CKE(dev_alloc_bytes(ctx, reinterpret_cast<void**>(&d_values), values_bytes + 64));
    CKE(dev_alloc_bytes(ctx, reinterpret_cast<void**>(&d_scratch), scratch_total + 64));
    CKE(dev_alloc(ctx, &d_pieces, std::max<uint32_t>(n_pieces, 1)));
    std::vector<EncPiece> all;
    all.reserve(n_pieces);
    {
        uint64_t base = 0;
        for (int k = 0; k < 3; k++) {
            for (auto& p : pieces[k]) { p.scratch = d_scratch + base + reinterpret_cast<uintptr_t>(p.scratch); all.push_back(p); }
            base += scratch[k];
        }
    }
    CKE(cudaEventRecord(ev0, st));
    if (values_bytes) CKE(cudaMemcpyAsync(d_values, values, values_bytes, cudaMemcpyHostToDevice, st));
    if (n_pieces) CKE(cudaMemcpyAsync(d_pieces, all.data(), (uint64_t)n_pieces * sizeof(EncPiece), cudaMemcpyHostToDevice, st));
    if (scratch[2]) CKE(cudaMemsetAsync(d_scratch + scratch[0] + scratch[1], 0, scratch[2], st));  // FastPFOR words are OR-ed together
    CKE(cudaEventRecord(ev1, st));
    CKE(launch_encode_pieces(d_values, d_pieces, np[0], d_pieces + np[0], np[1], d_pieces + np[0] + np[1], np[2], flags, st));
    if (n_pieces) CKE(cudaMemcpyAsync(all.data(), d_pieces, (uint64_t)n_pieces * sizeof(EncPiece), cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
    // final place of every stream (16-byte aligned, in descriptor order) and of every piece inside its stream
    std::vector<uint64_t> stream_len(n, 0);
    for (auto& p : all) stream_len[p.stream] += p.byte_length;
    uint64_t arena = 0, out_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (descs[i].status != COVT_OK) continue;
        if (stream_len[i] > 0xffffffffull) { descs[i].status = COVT_ERR_INVALID_ARG; stream_len[i] = 0; continue; }
        descs[i].out_offset = arena;
        descs[i].byte_length = (uint32_t)stream_len[i];
        out_bytes += stream_len[i];
        arena += (stream_len[i] + 15) & ~15ull;
    }
    {
        std::vector<uint64_t> cursor(n, 0);
        for (auto& p : all) {  // (the pieces of a stream are in order inside their kind)
            p.out_offset = descs[p.stream].out_offset + cursor[p.stream];
            cursor[p.stream] += p.byte_length;
            if (descs[p.stream].status != COVT_OK) p.byte_length = 0;
        }
    }
    R->counts[COVT_BUF_STREAM_ARENA] = arena;
    CKE(dev_alloc_bytes(ctx, &R->arena, arena + 64));
    R->bufs[COVT_BUF_STREAM_ARENA] = R->arena;
    if (n_pieces) CKE(cudaMemcpyAsync(d_pieces, all.data(), (uint64_t)n_pieces * sizeof(EncPiece), cudaMemcpyHostToDevice, st));
    CKE(launch_encode_compact(d_pieces, n_pieces, reinterpret_cast<uint8_t*>(R->arena), st));
    CKE(cudaEventRecord(ev2, st));
    CKE(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&R->timing.h2d_ms, ev0, ev1);
    cudaEventElapsedTime(&R->timing.decode_ms, ev1, ev2);  // (device time of the encode: kernels + the size read-back between them)
    R->timing.payload_bytes = out_bytes;  // compressed bytes written
    R->timing.output_bytes = in_bytes;    // value bytes read
    R->timing.kernel_launches = (np[0] ? 1 : 0) + (np[1] ? 1 : 0) + (np[2] ? 1 : 0) + (n_pieces ? 1 : 0);
    cleanup_tmp();
#undef CKE
    *out = R;
    return COVT_OK;
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
uint32_t covt_result_num_tiles(const covt_result* res) { return res ? res->n_tiles : 0; }
uint32_t covt_result_num_layers(const covt_result* res) { return res ? res->n_layers : 0; }

int32_t covt_result_layers(covt_result* res, const covt_layer** layers)
{
    if (!res || !layers) return COVT_ERR_INVALID_ARG;
    covt_ctx* ctx = res->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!res->h_layers) {
        const double t0 = now_ms();  // (host clock around a synchronous copy: no events to leak on the error paths)
        CK(pinned_take(ctx, reinterpret_cast<void**>(&res->h_layers), std::max<uint64_t>(res->n_layers, 1) * sizeof(covt_layer)));
        if (res->n_layers) CK(cudaMemcpyAsync(res->h_layers, res->d_layers, (uint64_t)res->n_layers * sizeof(covt_layer), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        res->timing.d2h_ms += (float)(now_ms() - t0);
    }
    *layers = res->h_layers;
    return COVT_OK;
}

int32_t covt_result_tile_status(covt_result* res, const uint32_t** status, const uint32_t** first_layer)
{
    if (!res) return COVT_ERR_INVALID_ARG;
    covt_ctx* ctx = res->ctx;
    CK(cudaSetDevice(ctx->device));
    if (!res->h_tile_status) {
        const double t0 = now_ms();
        CK(pinned_take(ctx, reinterpret_cast<void**>(&res->h_tile_status), ((uint64_t)res->n_tiles + 1) * sizeof(uint32_t)));
        CK(pinned_take(ctx, reinterpret_cast<void**>(&res->h_first_layer), ((uint64_t)res->n_tiles + 2) * sizeof(uint32_t)));
        if (res->n_tiles && res->d_tile_status) {
            CK(cudaMemcpyAsync(res->h_tile_status, res->d_tile_status, (uint64_t)res->n_tiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(res->h_first_layer, res->d_first_layer, ((uint64_t)res->n_tiles + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            res->h_first_layer[0] = 0;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        res->timing.d2h_ms += (float)(now_ms() - t0);
    }
    if (status) *status = res->h_tile_status;
    if (first_layer) *first_layer = res->h_first_layer;
    return COVT_OK;
}

int32_t covt_result_buffer(const covt_result* res, uint32_t which, const void** device_ptr, uint64_t* count, uint32_t* elem_size)
{
    if (!res || which >= COVT_NUM_BUFFERS) return COVT_ERR_INVALID_ARG;
    if (device_ptr) *device_ptr = res->bufs[which];
    if (count) *count = res->counts[which];
    if (elem_size) *elem_size = kBufElemSize[which];
    return COVT_OK;
}

int32_t covt_result_read(covt_result* res, uint32_t which, uint64_t elem_offset, uint64_t count, void* host_dst)
{
    if (!res || which >= COVT_NUM_BUFFERS || (!host_dst && count)) return COVT_ERR_INVALID_ARG;
    covt_ctx* ctx = res->ctx;
    if (elem_offset > res->counts[which] || count > res->counts[which] - elem_offset) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_result_read: range outside the buffer");
    if (!count) return COVT_OK;
    CK(cudaSetDevice(ctx->device));
    const uint32_t es = kBufElemSize[which];
    CK(cudaMemcpyAsync(host_dst, reinterpret_cast<const uint8_t*>(res->bufs[which]) + elem_offset * es, count * es, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return COVT_OK;
}

int32_t covt_result_prop_columns(covt_result* res, const covt_prop_column** columns, uint32_t* n_columns)
{
    if (!res || !columns || !n_columns) return COVT_ERR_INVALID_ARG;
    const int32_t rc = fetch_records(res, &res->h_prop_cols, res->d_prop_cols, res->n_prop_cols);
    if (rc != COVT_OK) return rc;
    *columns = res->h_prop_cols;
    *n_columns = res->n_prop_cols;
    return COVT_OK;
}

int32_t covt_result_prop_dictionaries(covt_result* res, const covt_prop_dictionary** dictionaries, uint32_t* n_dictionaries)
{
    if (!res || !dictionaries || !n_dictionaries) return COVT_ERR_INVALID_ARG;
    const int32_t rc = fetch_records(res, &res->h_prop_dicts, res->d_prop_dicts, res->n_prop_dicts);
    if (rc != COVT_OK) return rc;
    *dictionaries = res->h_prop_dicts;
    *n_dictionaries = res->n_prop_dicts;
    return COVT_OK;
}

int32_t covt_result_prop_buffer(const covt_result* res, uint32_t which, const void** device_ptr, uint64_t* count, uint32_t* elem_size)
{
    if (!res || which >= COVT_NUM_PROP_BUFFERS) return COVT_ERR_INVALID_ARG;
    if (device_ptr) *device_ptr = res->pbufs[which];
    if (count) *count = res->pcounts[which];
    if (elem_size) *elem_size = kPropBufElemSizeHost[which];
    return COVT_OK;
}

int32_t covt_result_prop_read(covt_result* res, uint32_t which, uint64_t elem_offset, uint64_t count, void* host_dst)
{
    if (!res || which >= COVT_NUM_PROP_BUFFERS || (!host_dst && count)) return COVT_ERR_INVALID_ARG;
    covt_ctx* ctx = res->ctx;
    if (elem_offset > res->pcounts[which] || count > res->pcounts[which] - elem_offset) return fail(ctx, COVT_ERR_INVALID_ARG, "covt_result_prop_read: range outside the buffer");
    if (!count) return COVT_OK;
    CK(cudaSetDevice(ctx->device));
    const uint32_t es = kPropBufElemSizeHost[which];
    CK(cudaMemcpyAsync(host_dst, reinterpret_cast<const uint8_t*>(res->pbufs[which]) + elem_offset * es, count * es, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return COVT_OK;
}

int32_t covt_result_timing(const covt_result* res, covt_timing* out)
{
    if (!res || !out) return COVT_ERR_INVALID_ARG;
    *out = res->timing;
    return COVT_OK;
}

int32_t covt_result_kernel_times(const covt_result* res, covt_kernel_time* out, uint32_t cap, uint32_t* n)
{
    if (!res || !n) return COVT_ERR_INVALID_ARG;
    *n = (uint32_t)res->kernel_times.size();
    for (uint32_t i = 0; i < cap && i < *n; i++) out[i] = res->kernel_times[i];
    return COVT_OK;
}

void covt_result_free(covt_result* res)
{
    if (!res) return;
    covt_ctx* ctx = res->ctx;
    cudaSetDevice(ctx->device);
    dev_free(ctx, res->arena);
    dev_free(ctx, res->d_layers);
    dev_free(ctx, res->d_tile_status);
    dev_free(ctx, res->d_first_layer);
    pinned_give(ctx, res->h_layers, std::max<uint64_t>(res->n_layers, 1) * sizeof(covt_layer));
    pinned_give(ctx, res->h_tile_status, ((uint64_t)res->n_tiles + 1) * sizeof(uint32_t));
    pinned_give(ctx, res->h_first_layer, ((uint64_t)res->n_tiles + 2) * sizeof(uint32_t));
    dev_free(ctx, res->prop_arena);
    dev_free(ctx, res->d_prop_cols);
    dev_free(ctx, res->d_prop_dicts);
    pinned_give(ctx, res->h_prop_cols, std::max<uint64_t>(res->n_prop_cols, 1) * sizeof(covt_prop_column));
    pinned_give(ctx, res->h_prop_dicts, std::max<uint64_t>(res->n_prop_dicts, 1) * sizeof(covt_prop_dictionary));
    delete res;
}

// ------------------------------------------------------------------------------------------------
// batch scheduler helper: contiguous tile ranges balanced by payload bytes (no collective: tiles share nothing)
// ------------------------------------------------------------------------------------------------
int32_t covt_partition_tiles(const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t n_parts, uint32_t* starts)
{
    if (!tile_offsets || !starts || !n_parts) return COVT_ERR_INVALID_ARG;
    const uint64_t base = tile_offsets[0], total = tile_offsets[n_tiles] - base;
    starts[0] = 0;
    uint32_t t = 0;
    for (uint32_t p = 1; p < n_parts; p++) {
        // first tile whose start offset reaches p/n_parts of the bytes (binary search: the offsets are non-decreasing; on offsets
        // that are not, the result is still a non-decreasing list of tile indices in range, and the decode call rejects the batch)
        const uint64_t target = base + (uint64_t)(((__uint128_t)total * p) / n_parts);
        t = (uint32_t)(std::lower_bound(tile_offsets + t, tile_offsets + n_tiles, target) - tile_offsets);
        starts[p] = t;
    }
    starts[n_parts] = n_tiles;
    return COVT_OK;
}

}  // extern "C"
