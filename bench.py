#!/usr/bin/env python
"""bench.py — COVT tile-batch decode on N B200s (BASELINE.json metric: compressed GB/s & Mvertices/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tiles|fixtures|varint1g] [--impl reference]

One "step" = one pass of the hot path (container walk, every stream codec, geometry assembly) over one batch.
Default workload = BASELINE config 5: ONE batch of 1 048 576 synthetic mixed-geometry gen-2b tiles (seed = tile index) PARTITIONED by
tile index over the N GPUs (strong scaling: covt_partition_tiles cuts contiguous ranges balanced by payload bytes, rank r decodes
range r; tiles share nothing, so there is no collective on the data path). --scaling weak: every rank its own --tiles tiles.

  value      compressed GB/s, whole job, inputs already resident in HBM, timed with CUDA events on the library's
             launching stream (first kernel start -> last kernel end), max over ranks
  e2e        the same metric through the reference-facing C-ABI call covt_decode_batch with HOST (pinned) buffers:
             host->device copy, decode and the device->host read of the per-tile status + layer index inside the timed region
  e2e_host   e2e plus the device->host read of the assembled GeoArrow-style buffers and ids into pinned memory (what a
             List<Layer> caller with a host-side consumer gets)
  h2d_ceiling  a bare pinned->device copy of the same bytes on all ranks at once: what the box gives the e2e figure at most
  library_scheduler (N > 1)  the same host batch through ONE covt_decode_batch_multi call from rank 0 over all N GPUs
  roofline   dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak (+ roofline.step: whole step)
  cpu_baseline / --impl reference: the CPU oracle (C restatement of the reference Java decoder; no JVM in this image)
             timed on the box's host cores — test infrastructure used only as the checker/baseline, never as the product.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "covt_tile_batch_decode_compressed_GBps"
CACHE_DIR = os.environ.get("COVT_BENCH_CACHE", "/tmp/covt_bench_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ workloads
def make_tiles(first_tile, n_tiles, container=0):
    """Config 5 tiles; cached on local disk because the driver runs several arms back to back on one box."""
    from tools.gen import gen as G
    os.makedirs(CACHE_DIR, exist_ok=True)
    key = os.path.join(CACHE_DIR, "tiles_c%d_%d_%d" % (container, first_tile, n_tiles))
    if os.path.exists(key + ".json"):
        try:
            meta = json.load(open(key + ".json"))
            blob = np.fromfile(key + ".blob", dtype=np.uint8)
            offs = np.fromfile(key + ".offs", dtype=np.uint64)
            if len(blob) == meta["bytes"] and len(offs) == n_tiles + 1:
                return blob, offs, meta["truth"]
        except Exception:
            pass
    t0 = time.time()
    blob, offs, truth = G.tiles(first_tile, n_tiles, G.default_params(container=container))
    log("[bench] generated %d tiles (%.1f MB) in %.1f s" % (n_tiles, len(blob) / 1e6, time.time() - t0))
    try:
        import shutil
        if shutil.disk_usage(CACHE_DIR).free > 4 * len(blob) + (8 << 30):  # never fill the box's disk (or a RAM-backed /tmp)
            blob.tofile(key + ".blob")
            offs.tofile(key + ".offs")
            json.dump({"bytes": int(len(blob)), "truth": truth}, open(key + ".json", "w"))
    except Exception as e:  # a full disk must not fail the bench
        log("[bench] cache write failed:", e)
    return blob, offs, truth


def make_fixture_sweep(replicas, decode_pfor=None):
    """Config 2: the reference's gen-2b OMT fixture tiles z2-z14 in one batch, replicated for timing. decode_pfor(tile array,
    offset, byte_length, num_values) -> values: when given, the FastPFOR topology streams are transcoded to ORC RLE first
    (tools/gen/rewrite.py: the "RLE topology streams" variant BASELINE config 2 names)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    tiles = [b for n, b in util.load_fixture_tiles() if n.startswith("omt/")]
    if decode_pfor is not None:
        from tools.gen import rewrite
        n_streams = 0
        for i, data in enumerate(tiles):
            arr = np.frombuffer(data + bytes(64), dtype=np.uint8)
            todo = rewrite.topology_pfor_streams(data)
            tiles[i] = rewrite.transcode_topology_to_rle(data, {off: decode_pfor(arr, off, bl, nv) for off, bl, nv in todo})
            n_streams += len(todo)
        log("[bench] transcoded %d FastPFOR topology streams to ORC RLE" % n_streams)
    blob1, offs1 = util.concat_tiles(tiles)
    blob = np.tile(blob1, replicas)
    sizes = np.tile(np.diff(offs1), replicas)
    offs = np.zeros(len(sizes) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(sizes)
    return blob, offs, None


def read_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): an NVML polling thread
    (same counters nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.* prints; a 100 ms nvidia-smi loop is too coarse
    for a timed region of a few hundred milliseconds), with the nvidia-smi loop as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        import threading
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.stop_flag = threading.Event()
        self.thread = self.p = self.f = None
        try:
            import pynvml as N
            N.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps ordinals: resolve through the PCI bus id of the CUDA device
            try:
                import torch
                bus = torch.cuda.get_device_properties(device).pci_bus_id
                h = N.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (torch.cuda.get_device_properties(device).pci_domain_id, bus,
                                                                            torch.cuda.get_device_properties(device).pci_device_id)).encode())
            except Exception:
                h = N.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = (("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap))

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for nm, bit in names:
                            if r & bit:
                                self.reasons.add(nm)
                        self.power.append(N.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    except Exception:
                        pass
                    self.stop_flag.wait(0.005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                          stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                out.update({"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(min(self.samples)), "sm_max_mhz": self.max_mhz,
                            "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml",
                            "power_w_max": max(self.power) if self.power else None})
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm = []
        reasons = set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        out["source"] = "nvidia-smi"
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, restated in C (oracle/), on the host cores."""
    if rank != 0:
        return
    from oracle import oracle as O
    abi = O.abi
    blob, offs, truth, cfg, container, flags = build_workload(args, 0, for_cpu=args.scaling == "strong")
    threads = os.cpu_count() or 1
    if args.workload == "varint1g":
        threads = 1  # one delta chain: sequential on the CPU
        sample = min(len(blob), 128 << 20)
        nv = min(truth["stream_values"], (sample // 4) & ~1)

        def one_step():
            vals, st_, cons = O.decode_stream(blob[:sample + 8], abi.OP_VARINT_ZZ_DELTA_XY, num_values=nv)
            return cons, len(vals) // 2
        sample_note = "the first %d values of the stream per step, one thread" % nv
    else:
        def one_step():
            rc, p, v, _cs = O.decode_batch_timed(blob, offs, container, flags & ~abi.FLAG_DECODE_PROPERTIES, n_threads=threads)
            if flags & abi.FLAG_DECODE_PROPERTIES:  # (the property oracle is single-threaded: the batch is walked once more)
                pr = O.decode_properties(blob, offs, container, flags)
                p += int(pr.payload_bytes)
            return p, v
        sample_note = "%d tiles per step (%s)" % (len(offs) - 1, "the whole batch the GPU arm partitions" if scaling_of(args) == "strong" else "the rank-0 batch of the GPU arm")
    cfg.update({"container": "gen-2b", "flags": "CLOSE_RINGS"})
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    pb = vx = 0
    for _ in range(args.steps):
        p, v = one_step()
        pb += p
        vx += v
    dt = time.perf_counter() - t0
    gbps = pb / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": gbps, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling_of(args),
            "vs_baseline": None, "dtype": "int32", "data": cfg.pop("data", "synthetic"), "config": cfg,
            "mvertices_per_s": vx / dt / 1e6,
            "workload_stats": {"payload_bytes": pb / args.steps, "vertices": vx / args.steps},
            "cpu_baseline": {"value": gbps, "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": sample_note + "; C restatement of the Java decoder (JVM unavailable), -O2, pthreads"},
            "e2e": {"value": gbps, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def scaling_of(args):
    return "strong" if (args.workload == "tiles" and args.scaling == "strong") else "weak"


def run_library_scheduler(args, covt, abi, world, container, flags):
    """Rank 0 alone: the whole host batch through ONE covt_decode_batch_multi call (the library's batch scheduler: one host thread
    + context + pinned staging per GPU, covt_partition_tiles on the call path). The other ranks wait at a barrier meanwhile."""
    import torch
    blob, offs, truth = make_tiles(0, args.tiles)
    pinned = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = blob
    payload = None
    md = covt.MultiDecoder(list(range(world)))
    try:
        def one():
            res = md.decode_batch(pinned.numpy(), offs, container, flags)
            t = res.timing()
            verts = 0
            for p in res.parts:
                p["result"].touch_tile_status()
            res.free()
            return t
        for _ in range(2):
            one()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t = one()
        dt = (time.perf_counter() - t0) / args.steps
        payload = t["payload_bytes"]
        return {"n_gpus": md.n_devices, "ms_per_call": dt * 1e3, "e2e_GBps": payload / dt / 1e9, "device_ms_max_over_gpus": t["decode_ms"],
                "h2d_ms_max_over_gpus": t["h2d_ms"], "vertices": t["vertices"], "ok": t["vertices"] == truth["vertices"],
                "note": "one process, one call, N GPUs: covt_decode_batch_multi from rank 0 with host input (pinned) while the other ranks idle"}
    finally:
        md.close() if hasattr(md, "close") else None


def build_workload(args, rank, for_cpu=False, world=1, sync=None):
    """-> blob, tile offsets, truth, config, container, flags of THIS rank's share. sync(): a barrier over the ranks (strong
    scaling: rank 0 generates the one batch into the disk cache, the others read it)."""
    import covt_loader
    covt = covt_loader.load()
    abi = covt.abi
    flags = abi.FLAG_DEFAULT
    container = abi.CONTAINER_GEN2B
    if args.workload == "tiles" and (args.scaling == "strong" or for_cpu):
        n = args.tiles
        if rank == 0 or sync is None:
            blob, offs, truth = make_tiles(0, n)
        if sync is not None and world > 1:
            sync()
            if rank != 0:
                blob, offs, truth = make_tiles(0, n)  # from the cache rank 0 wrote (generated again if the disk was full)
        cfg = {"workload": "config5: ONE batch of %d synthetic mixed-geometry gen-2b tiles, seed = tile index, 2 layers/tile, "
                           "partitioned by tile index over the GPUs" % n,
               "tiles": n, "partition": "covt_partition_tiles: contiguous tile ranges balanced by payload bytes, no collective",
               "l2": "inputs (%.2f GB) and outputs far larger than the 126 MB L2" % (len(blob) / 1e9)}
        truth = dict(truth, whole_batch=True)
        if world > 1 and not for_cpu:
            # this rank's range of the ONE batch: covt_partition_tiles (the split the library scheduler makes inside
            # covt_decode_batch_multi) through the one-process-per-GPU helper of the package
            blob, offs, _first_tile = covt.scheduler.rank_slice(blob, offs, rank, world)
            blob, offs = blob.copy(), offs.astype(np.uint64)
    elif args.workload == "tiles":
        n = args.tiles
        blob, offs, truth = make_tiles(rank * n, n)
        cfg = {"workload": "config5 (weak scaling): %d synthetic mixed-geometry gen-2b tiles per GPU, seed = tile index, 2 layers/tile" % n,
               "tiles_per_gpu": n, "partition": "tile index ranges, no collective", "l2": "inputs (%.2f GB) and outputs far larger than the 126 MB L2" % (len(blob) / 1e9)}
    elif args.workload == "fixtures":
        decode_pfor = None
        if args.rle_topology and for_cpu:
            # the reference arm prepares its workload with its own decoder (the oracle)
            from oracle import oracle as O

            def decode_pfor(arr, off, bl, nv):
                return O.decode_stream(arr, abi.OP_PFOR_ZZ_DELTA, byte_offset=off, byte_length=bl, num_values=nv)[0]
        elif args.rle_topology:
            # the GPU arm prepares it with the product (no oracle on this path)
            prep = covt_loader.load().Decoder(int(os.environ.get("LOCAL_RANK", 0)))

            def decode_pfor(arr, off, bl, nv):
                vals, st_, _ = prep.decode_stream(arr, abi.OP_PFOR_ZZ_DELTA, byte_offset=off, byte_length=bl, num_values=nv)
                if st_ != 0:
                    raise RuntimeError("transcoding: FastPFOR stream status %d" % st_)
                return vals
        blob, offs, truth = make_fixture_sweep(args.replicas, decode_pfor)
        flags |= abi.FLAG_ID_DVZZ_IS_RLE
        if args.props:
            flags |= abi.FLAG_DECODE_PROPERTIES
        cfg = {"workload": "config2%s%s: the reference's 91 gen-2b OMT fixture tiles z2-z14 in one batch, x%d replicas" % (
            " (RLE topology streams: FastPFOR topology transcoded to ORC RLE)" if args.rle_topology else "",
            " + property columns (COVT_FLAG_DECODE_PROPERTIES)" if args.props else "", args.replicas),
               "l2": "inputs larger than L2"}
    elif args.workload == "varint1g":
        from tools.gen import gen as G
        os.makedirs(CACHE_DIR, exist_ok=True)
        key = os.path.join(CACHE_DIR, "varint_%d_%d" % (args.stream_bytes, rank))
        if os.path.exists(key + ".json"):
            blob = np.fromfile(key + ".blob", dtype=np.uint8)
            nvals = json.load(open(key + ".json"))["n"]
        else:
            t0 = time.time()
            blob, nvals = G.varint_stream(args.stream_bytes, seed=0xC0717 + rank)
            log("[bench] generated a %d-byte varint stream in %.1f s" % (len(blob), time.time() - t0))
            try:
                import shutil
                if shutil.disk_usage(CACHE_DIR).free > 4 * len(blob) + (8 << 30):
                    blob.tofile(key + ".blob")
                    json.dump({"n": nvals}, open(key + ".json", "w"))
            except Exception as e:
                log("[bench] cache write failed:", e)
        offs = np.array([0, len(blob)], dtype=np.uint64)
        truth = {"stream_values": nvals}
        cfg = {"workload": "config3: one PLAIN VERTEX_BUFFER / VARINT_DELTA_ZIG_ZAG stream of %d bytes per GPU (%d ints), seed 0xC0717" % (len(blob), nvals),
               "l2": "input 1 GiB and output 2.8 GB far larger than the 126 MB L2"}
    else:
        raise SystemExit("unknown workload " + args.workload)
    cfg["data"] = "reference fixtures (replicated)" if args.workload == "fixtures" else "synthetic"
    return blob, offs, truth, cfg, container, flags


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs (and, by first touch, the host memory) of the NUMA node its GPU hangs off: with 8 ranks
    uploading at once, pinned buffers on the wrong socket halve the host->device rate."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import covt_loader
    covt = covt_loader.load()
    abi = covt.abi
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dec = covt.Decoder(local_rank)  # raises without the CUDA library / device: no CPU fallback

    def sync_ranks():
        if world > 1:
            dist.barrier()
    blob_np, offs, truth, cfg, container, flags = build_workload(args, rank, world=world, sync=sync_ranks)
    n_tiles = len(offs) - 1
    # pinned host copy of the inputs (what a caller hands to covt_decode_batch)
    pinned = torch.empty(len(blob_np), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = blob_np
    offs_pinned = torch.empty(len(offs), dtype=torch.int64, pin_memory=True)
    offs_pinned.numpy().view(np.uint64)[:] = offs
    del blob_np
    blob_ptr, offs_ptr = pinned.data_ptr(), offs_pinned.data_ptr()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream_mode = args.workload == "varint1g"
    if stream_mode:
        import ctypes as C
        descs = (abi.StreamDesc * 1)()

        def fresh_descs():
            descs[0] = abi.StreamDesc(byte_offset=0, byte_length=pinned.numel(), num_values=truth["stream_values"],
                                      op=abi.OP_VARINT_ZZ_DELTA_XY)
            return descs

        class _D:  # same call shapes as the batch path
            @staticmethod
            def decode(batch, container, fl):
                return dec.decode_streams(batch, fresh_descs(), fl)

            @staticmethod
            def decode_batch_raw(bp, op_, n, container, fl):
                h = C.c_void_p()
                dec._check(covt.lib().covt_decode_streams(dec._h, bp, pinned.numel(), fresh_descs(), 1, fl, C.byref(h)))
                return covt.Result(dec, h)
        runner = _D
    else:
        runner = dec

    # ---------------- device-resident decode: inputs already in HBM ----------------
    batch = dec.upload_raw(blob_ptr, offs_ptr, n_tiles, pinned.numel())
    pflags = flags | abi.FLAG_PROFILE_KERNELS
    for _ in range(max(args.warmup, 3)):
        r = runner.decode(batch, container, flags)
        r.free()
    barrier()
    sampler = ClockSampler(local_rank)
    dev_ms = 0.0
    launches = 0
    payload = verts = outb = 0
    bad_tiles = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        r = runner.decode(batch, container, flags)
        t = r.timing()
        dev_ms += t["decode_ms"]
        launches += t["kernel_launches"]
        payload, verts, outb = t["payload_bytes"], t["vertices"], t["output_bytes"]
        r.free()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    # per-kernel breakdown: a separate pass with one CUDA event pair per kernel, everything on the main stream (in the timed pass
    # above the codec classes run one after the other too, but the SECOND pass of a class — its queued large streams — runs on a side
    # stream beside the next class's first pass, so individual durations would overlap)
    ktimes = {}
    prof_ms = 0.0
    for _ in range(args.steps):
        r = runner.decode(batch, container, pflags)
        prof_ms += r.timing()["decode_ms"]
        for k in r.kernel_times():
            e = ktimes.setdefault(k["name"], {"ms": 0.0, "launches": 0, "bytes": 0})
            e["ms"] += k["ms"]
            e["launches"] += k["launches"]
            e["bytes"] += k["algorithmic_bytes"]
        r.free()
    barrier()
    # parity guard outside the timed region: every tile decoded, totals equal what was encoded
    r = runner.decode(batch, container, flags)
    if stream_mode:
        assert descs[0].status == 0 and descs[0].out_count == truth["stream_values"], "stream decode failed"
        verts = truth["stream_values"] // 2
        st = np.zeros(1, np.uint32)
    else:
        st, _first = r.tile_status()
    bad_tiles = int((st != 0).sum())
    if truth is not None and not stream_mode:
        L = r.layers
        sums = torch.tensor([int(L["n_vertices"].sum()), int(L["n_rings"].sum()), int(L["n_parts"].sum())], dtype=torch.float64, device="cuda")
        if world > 1 and truth.get("whole_batch"):
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)  # the truth counts belong to the whole batch, the ranks hold its partitions
        assert int(sums[0].item()) == truth["vertices"], "decoded vertex count differs from what was encoded"
        assert int(sums[1].item()) == truth["rings"] and int(sums[2].item()) == truth["parts"]
    # what a host-side consumer reads back: the assembled GeoArrow-style buffers + ids (element counts of this rank's result)
    host_bufs = [] if stream_mode else [abi.BUF_S_IDS, abi.BUF_S_GEOMETRY_TYPES, abi.BUF_A_GEOM_OFFSETS, abi.BUF_A_PART_OFFSETS,
                                        abi.BUF_A_RING_OFFSETS, abi.BUF_A_COORDS]
    host_counts = {b: r.device_buffer(b)[1] for b in host_bufs}
    r.free()
    batch.free()

    # ---------------- end to end through the C-ABI call with host buffers ----------------
    for _ in range(2):
        r = runner.decode_batch_raw(blob_ptr, offs_ptr, n_tiles, container, flags)
        r.touch_tile_status()
        r.free()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = runner.decode_batch_raw(blob_ptr, offs_ptr, n_tiles, container, flags)
        r.touch_tile_status()  # device->host read of the step's result index: per-tile status + first-layer table
        r.free()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d_bytes = pinned.numel() + offs_pinned.numel() * 8
    d2h_bytes = n_tiles * 4 + (n_tiles + 1) * 4 + (1 + abi.NUM_BUFFERS + 4) * 8

    # ---------------- the ceiling of e2e: a bare pinned->device copy of the same bytes, all ranks at once ----------------
    dev_in = torch.empty(pinned.numel(), dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        dev_in.copy_(pinned, non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        dev_in.copy_(pinned, non_blocking=True)
    ev1.record()
    barrier()
    h2d_copy_ms = ev0.elapsed_time(ev1)
    del dev_in

    # ---------------- end to end INTO HOST MEMORY: the same call + the read-back of the assembled buffers and ids ----------------
    e2e_host_s, d2h_host_bytes = 0.0, 0
    if host_bufs and not args.no_e2e_host:
        sizes = {b: host_counts[b] * np.dtype(abi.BUF_DTYPES[b]).itemsize for b in host_bufs}
        try:
            host_out = {b: torch.empty(max(int(sizes[b] * 1.02) + 4096, 16), dtype=torch.uint8, pin_memory=True) for b in host_bufs}
        except Exception as e:  # not enough lockable host memory on this box: report the figure as absent, not as a failure
            log("[bench] e2e_host skipped:", e)
            host_out = None
        if host_out is not None:
            sink = abi.HostSink()
            for b in host_bufs:
                sink.ptr[b] = host_out[b].data_ptr()
                sink.capacity[b] = host_out[b].numel() // np.dtype(abi.BUF_DTYPES[b]).itemsize

            def host_step():
                # covt_decode_batch_to_host: the read-back of a segment's results overlaps the upload and decode of the next ones
                r = runner.decode_batch_to_host_raw(blob_ptr, offs_ptr, n_tiles, sink, container, flags)
                r.touch_tile_status()
                got = sum(r.device_buffer(b)[1] * np.dtype(abi.BUF_DTYPES[b]).itemsize for b in host_bufs)
                r.free()
                return got
            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                d2h_host_bytes = host_step()
            barrier()
            e2e_host_s = time.perf_counter() - t0
            del host_out

    # ---------------- reduce over ranks ----------------
    tt = torch.tensor([dev_ms, e2e_s * 1e3, wall_ms, h2d_copy_ms, e2e_host_s * 1e3], dtype=torch.float64, device="cuda")
    ss = torch.tensor([payload, verts, outb, launches, bad_tiles, h2d_bytes, d2h_bytes, d2h_host_bytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, wall_ms_max, h2d_copy_ms_max, e2e_host_ms_max = [float(x) for x in tt.tolist()]
    payload_all, verts_all, outb_all, launches_all, bad_all, h2d_all, d2h_all, d2h_host_all = [float(x) for x in ss.tolist()]

    # ---------------- N > 1: the in-library batch scheduler, ONE call from ONE process over all N GPUs ----------------
    sched = None
    if world > 1 and args.workload == "tiles" and args.scaling == "strong" and not args.no_scheduler:
        del pinned, offs_pinned
        # the other ranks must wait on the HOST: an NCCL barrier would park a spinning kernel on the GPUs the scheduler is using
        cpu_group = dist.new_group(backend="gloo")
        barrier()
        if rank == 0:
            try:
                sched = run_library_scheduler(args, covt, abi, world, container, flags)
            except Exception as e:
                sched = {"unavailable": str(e)[:200]}
        dist.barrier(group=cpu_group)
        barrier()

    if rank == 0:
        peak, peak_src = read_peaks()
        steps = args.steps
        value = payload_all * steps / (dev_ms_max * 1e-3) / 1e9
        mverts = verts_all * steps / (dev_ms_max * 1e-3) / 1e6
        top = max(ktimes.items(), key=lambda kv: kv[1]["ms"]) if ktimes else (None, None)
        roof = None
        if top[0]:
            k = top[1]
            achieved = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0.0
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
                size_now = {"tiles_per_gpu": n_tiles, "payload_bytes_per_gpu": payload}.get(tj.get("size_key"))
                traffic = tj.get(top[0]) if tj.get("size") == size_now else None  # measured under ncu at exactly this size
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": k["bytes"] / max(k["launches"], 1),
                    "ms_per_launch": k["ms"] / max(k["launches"], 1), "share_of_step": k["ms"] / max(prof_ms, 1e-9),
                    "serialised_ms_per_step": prof_ms / steps,
                    "step": {"algorithmic_bytes": payload + outb, "achieved": (payload + outb) * steps / (dev_ms * 1e-3) / 1e9,
                             "frac": (payload + outb) * steps / (dev_ms * 1e-3) / 1e9 / peak,
                             "note": "whole step on rank 0: (payload read + final outputs written) / device time of all kernels"},
                    "kernels": {n: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps,
                                        "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0)} for n, v in ktimes.items()}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline and stream_mode:
            from oracle import oracle as O
            hb = pinned.numpy()
            sample = min(len(hb), 256 << 20)
            t0 = time.perf_counter()
            vals, st_, cons = O.decode_stream(hb[:sample + 8], abi.OP_VARINT_ZZ_DELTA_XY, num_values=min(truth["stream_values"], (sample // 4) & ~1))
            dt = time.perf_counter() - t0
            cpu = {"value": cons / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "mvertices_per_s": len(vals) / 2 / dt / 1e6,
                   "sample": "the first %d values (%.0f MB) of the same stream, one thread (a single delta chain is sequential on "
                             "the CPU); C restatement of DecodingUtils.decodeZigZagDeltaVarintCoordinates" % (len(vals), cons / 1e6)}
        elif world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            threads = os.cpu_count() or 1
            hb = pinned.numpy()
            O.decode_batch_timed(hb[: int(offs[min(2048, n_tiles)])], offs[: min(2048, n_tiles) + 1], container, flags, n_threads=threads)
            t0 = time.perf_counter()
            rc, pb, vx, _cs = O.decode_batch_timed(hb, offs, container, flags, n_threads=threads)
            dt = time.perf_counter() - t0
            cpu = {"value": pb / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "mvertices_per_s": vx / dt / 1e6,
                   "sample": "one pass over the same %d tiles (%.2f GB payload, %.1f s); C restatement of the reference Java "
                             "decoder (JVM unavailable), -O2, pthreads" % (n_tiles, pb / 1e9, dt)}
        host_binding = "each rank bound to the NUMA node of its GPU" if numa_node is not None else None
        cfg.update({"container": "gen-2b", "flags": "CLOSE_RINGS"})
        stats = {"payload_bytes": payload_all, "vertices": verts_all, "output_bytes": outb_all, "bad_tiles": bad_all,
                 "payload_bytes_rank0": payload, "vertices_rank0": verts, "output_bytes_rank0": outb, "tiles_rank0": n_tiles}
        if args.workload == "fixtures" and bad_all:
            stats["bad_tiles_cause"] = ("the committed omt/4_8_10 tile: its water_name layer labels a varint-coded ICE vertex buffer as FastPFOR "
                                      "(SURVEY 0-8b); one per replica, COVT_ERR_COUNT_MISMATCH like the oracle")
        h2d_ceiling = h2d_all * steps / (h2d_copy_ms_max * 1e-3) / 1e9 if h2d_copy_ms_max > 0 else None
        e2e_ms = e2e_ms_max / steps
        e2e = {"value": payload_all * steps / (e2e_ms_max * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(h2d_all),
               "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_ms,
               "mvertices_per_s": verts_all * steps / (e2e_ms_max * 1e-3) / 1e6,
               "h2d_ceiling_GBps": h2d_ceiling, "h2d_copy_ms_per_step": h2d_copy_ms_max / steps,
               "h2d_GBps_in_call": h2d_all / (e2e_ms * 1e-3) / 1e9,
               "fraction_of_h2d_ceiling": (h2d_all / (e2e_ms * 1e-3) / 1e9) / h2d_ceiling if h2d_ceiling else None,
               "note": "results stay device-resident (GeoArrow-style buffers); the device->host read is the per-tile status + layer index. "
                       "h2d_ceiling = a bare pinned->device copy of the same input bytes on all ranks at once"}
        e2e_host = None
        if e2e_host_ms_max > 0:
            e2e_host = {"value": payload_all * steps / (e2e_host_ms_max * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_host_ms_max / steps,
                        "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all + d2h_host_all),
                        "mvertices_per_s": verts_all * steps / (e2e_host_ms_max * 1e-3) / 1e6,
                        "note": "the same batch through covt_decode_batch_to_host: ids, geometry types and the assembled geom/part/ring offsets + "
                                "coordinates land in pinned host memory, read back segment by segment while later segments are uploaded and "
                                "decoded: what a List<Layer> caller with a host-side consumer gets"}
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "int32",
                "data": cfg.pop("data"), "config": cfg, "workload_stats": stats, "host_binding": host_binding,
                "mvertices_per_s": mverts, "wall_ms_per_step": wall_ms_max / steps,
                "step_roofline_frac": roof["step"]["frac"] if roof else None,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_host": e2e_host, "library_scheduler": sched,
                "gpu_launches": int(launches_all), "clocks": clocks}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else any library prints (NCCL's version banner,
    make output of the in-tree builds) was redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tiles", choices=["tiles", "fixtures", "varint1g"])
    ap.add_argument("--stream-bytes", type=int, default=1 << 30, help="config 3 stream size")
    ap.add_argument("--tiles", type=int, default=1 << 20, help="tiles per GPU (config 5: 1 048 576)")
    ap.add_argument("--replicas", type=int, default=256, help="fixture-sweep replicas (config 2)")
    ap.add_argument("--rle-topology", action="store_true", help="config 2 with every topology stream as ORC RLE (BASELINE wording)")
    ap.add_argument("--props", action="store_true", help="config 2 with the property columns decoded too (SURVEY 8 f1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="tiles workload on N GPUs: strong = ONE batch of --tiles tiles partitioned over the GPUs (BASELINE config 5); weak = --tiles per GPU")
    ap.add_argument("--no-e2e-host", action="store_true", help="skip the end-to-end-into-host-memory leg (it pins as much host memory as the results take)")
    ap.add_argument("--no-scheduler", action="store_true", help="N > 1: skip the covt_decode_batch_multi leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
