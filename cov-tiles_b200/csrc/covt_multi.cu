// covt_multi.cu — the batch scheduler of libcovt_b200: ONE call decodes a host batch on several GPUs of one box.
//
// The reference entry point being replaced is one call (CovtParser.decodeCovt, J/decoder/CovtParser.java:53) made from one JVM
// thread. Tiles share nothing (CovtParser.java:56-131 keeps no cross-layer state; every delta chain starts from 0,
// DecodingUtils.java:57,97-98,396), so the scheduler cuts the batch into contiguous tile ranges balanced by payload bytes
// (covt_partition_tiles) and hands range g to GPU g: every GPU has its own covt_ctx (its own CUDA streams, pinned staging, device
// block cache and result arena) and its own persistent host thread, which runs the ordinary single-GPU covt_decode_batch on its
// range. No collective, no peer traffic: the only shared resources are host memory bandwidth and the PCIe fabric.
// The result handle exposes one covt_result per GPU (results stay resident on the GPU that decoded them).
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/covt_b200.h"

namespace {

struct Job {
    const uint8_t* blob = nullptr;
    const uint64_t* tile_offsets = nullptr;  // the caller's array (absolute offsets)
    uint32_t t0 = 0, t1 = 0;
    uint32_t container = 0, flags = 0;
    const covt_tilejson* tilejson = nullptr;
};

struct Worker {
    int32_t device = 0;
    covt_ctx* ctx = nullptr;
    std::thread thread;
    std::mutex m;
    std::condition_variable cv;
    bool has_job = false, done = false, quit = false;
    Job job;
    int32_t rc = COVT_OK;
    covt_result* res = nullptr;
    uint64_t* offs = nullptr;    // re-based tile offsets of the range, PAGE-LOCKED: covt_decode_batch copies them up segment by segment on
                                 // its copy stream, and a copy from pageable memory would stall that stream's host side behind the blob copies
    size_t offs_cap = 0;
    std::string err;
    double wall_ms = 0.0;
};

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void worker_main(Worker* w)
{
    cudaSetDevice(w->device);
    for (;;) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        const Job j = w->job;
        lk.unlock();
        const double t_start = now_ms();
        const uint32_t n = j.t1 - j.t0;
        const uint64_t base = j.tile_offsets[j.t0];
        if (w->offs_cap < (size_t)n + 1) {
            if (w->offs) cudaFreeHost(w->offs);
            w->offs = nullptr;
            w->offs_cap = 0;
            const size_t cap = ((size_t)n + 1) * 5 / 4 + 1024;
            if (cudaMallocHost(reinterpret_cast<void**>(&w->offs), cap * sizeof(uint64_t)) == cudaSuccess) w->offs_cap = cap;
        }
        w->res = nullptr;
        if (!w->offs) {
            w->rc = COVT_ERR_OOM;
            w->err = "cudaMallocHost failed for the re-based tile offsets";
        } else {
            for (uint32_t i = 0; i <= n; i++) w->offs[i] = j.tile_offsets[j.t0 + i] - base;
            w->rc = covt_decode_batch(w->ctx, j.blob + base, w->offs, n, j.container, j.tilejson, j.flags, &w->res);
        }
        if (w->rc != COVT_OK && w->offs) {
            char buf[512];
            covt_last_error(w->ctx, buf, sizeof(buf));
            w->err = buf;
        }
        w->wall_ms = now_ms() - t_start;
        lk.lock();
        w->has_job = false;
        w->done = true;
        lk.unlock();
        w->cv.notify_all();
    }
}

}  // namespace

struct covt_multi {
    std::vector<Worker*> workers;
    std::string err;
    std::mutex call_mutex;  // one covt_decode_batch_multi at a time per handle
};

struct covt_multi_result {
    covt_multi* owner = nullptr;
    std::vector<covt_result*> parts;
    std::vector<uint32_t> starts;  // parts + 1 tile indices
    std::vector<int32_t> devices;
    std::vector<double> wall_ms;
};

namespace {
std::mutex g_multi_err_mutex;
std::string g_multi_create_error;
}

extern "C" {

int32_t covt_create_multi(uint32_t device_count, const int32_t* device_ids, covt_multi** out)
{
    if (!out) return COVT_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        std::lock_guard<std::mutex> g(g_multi_err_mutex);
        g_multi_create_error = "no CUDA device (libcovt_b200 has no CPU fallback)";
        return COVT_ERR_CUDA;
    }
    if (device_count == 0) device_count = (uint32_t)n_dev;  // 0 = every visible GPU
    covt_multi* M = new covt_multi();
    for (uint32_t i = 0; i < device_count; i++) {
        const int32_t dev = device_ids ? device_ids[i] : (int32_t)i;
        covt_ctx* ctx = nullptr;
        const int32_t rc = covt_create(dev, &ctx);
        if (rc != COVT_OK) {
            char buf[512];
            covt_last_error(nullptr, buf, sizeof(buf));
            {
                std::lock_guard<std::mutex> g(g_multi_err_mutex);
                g_multi_create_error = buf;
            }
            covt_destroy_multi(M);
            return rc;
        }
        Worker* w = new Worker();
        w->device = dev;
        w->ctx = ctx;
        w->thread = std::thread(worker_main, w);
        M->workers.push_back(w);
    }
    *out = M;
    return COVT_OK;
}

void covt_destroy_multi(covt_multi* M)
{
    if (!M) return;
    for (Worker* w : M->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->thread.joinable()) w->thread.join();
        if (w->offs) { cudaSetDevice(w->device); cudaFreeHost(w->offs); }
        covt_destroy(w->ctx);
        delete w;
    }
    delete M;
}

int32_t covt_multi_last_error(covt_multi* M, char* buf, size_t buf_len)
{
    if (!buf || !buf_len) return COVT_ERR_INVALID_ARG;
    std::string m;
    if (M) m = M->err;
    else { std::lock_guard<std::mutex> g(g_multi_err_mutex); m = g_multi_create_error; }
    snprintf(buf, buf_len, "%s", m.c_str());
    return COVT_OK;
}

uint32_t covt_multi_device_count(const covt_multi* M) { return M ? (uint32_t)M->workers.size() : 0; }

int32_t covt_multi_context(covt_multi* M, uint32_t part, covt_ctx** ctx, int32_t* device)
{
    if (!M || part >= M->workers.size()) return COVT_ERR_INVALID_ARG;
    if (ctx) *ctx = M->workers[part]->ctx;
    if (device) *device = M->workers[part]->device;
    return COVT_OK;
}

int32_t covt_decode_batch_multi(covt_multi* M, const uint8_t* blob, const uint64_t* tile_offsets, uint32_t n_tiles, uint32_t container,
                                const covt_tilejson* tilejson, uint32_t flags, covt_multi_result** out)
{
    if (!M || !out || (!blob && n_tiles) || !tile_offsets) return COVT_ERR_INVALID_ARG;
    *out = nullptr;
    std::lock_guard<std::mutex> call(M->call_mutex);
    const uint32_t P = (uint32_t)M->workers.size();
    covt_multi_result* R = new covt_multi_result();
    R->owner = M;
    R->starts.resize(P + 1);
    int32_t rc = covt_partition_tiles(tile_offsets, n_tiles, P, R->starts.data());
    if (rc != COVT_OK) { delete R; M->err = "covt_partition_tiles failed"; return rc; }
    for (uint32_t p = 0; p < P; p++) {
        Worker* w = M->workers[p];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = Job();
        w->job.blob = blob;
        w->job.tile_offsets = tile_offsets;
        w->job.t0 = R->starts[p];
        w->job.t1 = R->starts[p + 1];
        w->job.container = container;
        w->job.flags = flags;
        w->job.tilejson = tilejson;
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    R->parts.assign(P, nullptr);
    R->devices.resize(P);
    R->wall_ms.resize(P);
    for (uint32_t p = 0; p < P; p++) {
        Worker* w = M->workers[p];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        R->parts[p] = w->res;
        R->devices[p] = w->device;
        R->wall_ms[p] = w->wall_ms;
        w->res = nullptr;
        if (w->rc != COVT_OK && rc == COVT_OK) {
            rc = w->rc;
            char m[640];
            snprintf(m, sizeof(m), "device %d: %s", w->device, w->err.c_str());
            M->err = m;
        }
    }
    if (rc != COVT_OK) {
        covt_multi_result_free(R);
        return rc;
    }
    *out = R;
    return COVT_OK;
}

uint32_t covt_multi_result_parts(const covt_multi_result* R) { return R ? (uint32_t)R->parts.size() : 0; }

int32_t covt_multi_result_part(const covt_multi_result* R, uint32_t part, covt_result** res, uint32_t* first_tile, uint32_t* n_tiles, int32_t* device)
{
    if (!R || part >= R->parts.size()) return COVT_ERR_INVALID_ARG;
    if (res) *res = R->parts[part];
    if (first_tile) *first_tile = R->starts[part];
    if (n_tiles) *n_tiles = R->starts[part + 1] - R->starts[part];
    if (device) *device = R->devices[part];
    return COVT_OK;
}

int32_t covt_multi_result_timing(const covt_multi_result* R, covt_timing* out)
{
    if (!R || !out) return COVT_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    for (covt_result* r : R->parts) {
        covt_timing t;
        if (!r || covt_result_timing(r, &t) != COVT_OK) continue;
        out->h2d_ms = std::max(out->h2d_ms, t.h2d_ms);
        out->decode_ms = std::max(out->decode_ms, t.decode_ms);
        out->d2h_ms = std::max(out->d2h_ms, t.d2h_ms);
        out->kernel_launches += t.kernel_launches;
        out->payload_bytes += t.payload_bytes;
        out->output_bytes += t.output_bytes;
        out->vertices += t.vertices;
        out->segments += t.segments;
        out->capacity_retries += t.capacity_retries;
    }
    return COVT_OK;
}

// Brings result buffer `which` of every part back to the host, all GPUs copying at once: part p goes to host_dst[p], which must
// hold count(p) * element size bytes (covt_result_buffer of the part). Pinned destinations run at full PCIe rate.
int32_t covt_multi_result_read(covt_multi_result* R, uint32_t which, void* const* host_dst)
{
    if (!R || !host_dst || which >= COVT_NUM_BUFFERS) return COVT_ERR_INVALID_ARG;
    const size_t P = R->parts.size();
    std::vector<int32_t> rcs(P, COVT_OK);
    std::vector<std::thread> th;
    for (size_t p = 0; p < P; p++)
        th.emplace_back([&, p] {
            uint64_t count = 0;
            rcs[p] = covt_result_buffer(R->parts[p], which, nullptr, &count, nullptr);
            if (rcs[p] == COVT_OK && count) rcs[p] = covt_result_read(R->parts[p], which, 0, count, host_dst[p]);
        });
    for (auto& t : th) t.join();
    for (size_t p = 0; p < P; p++)
        if (rcs[p] != COVT_OK) { R->owner->err = "covt_multi_result_read: a device->host copy failed"; return rcs[p]; }
    return COVT_OK;
}

void covt_multi_result_free(covt_multi_result* R)
{
    if (!R) return;
    // a result is freed on the device that owns it (the library calls cudaSetDevice itself)
    for (covt_result* r : R->parts) covt_result_free(r);
    delete R;
}

}  // extern "C"
