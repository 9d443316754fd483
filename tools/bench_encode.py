#!/usr/bin/env python
"""bench_encode.py — the GPU stream encoders (covt_encode_streams) on one B200.

  python tools/bench_encode.py [--steps K] [--warmup W] [--stream-bytes B] [--replicas R]

Two workloads, each also a round-trip check (the encoders must give back the bytes the values were decoded from):
  varint1g   the values of BASELINE config 3 (one PLAIN VERTEX_BUFFER / VARINT_DELTA_ZIG_ZAG stream, 1 GiB compressed, ~708 M ints)
             -> encodeZigZagDeltaCoordinates + encodeVarints: must reproduce the stream byte for byte
  fixtures   every geometry / id stream of the reference's 129 fixture tiles (decoded on the GPU), x R replicas in ONE call
             -> must reproduce the fixture bytes (FastPFOR streams of more than 65 536 values excepted)
Reported per workload: device time of the encode (CUDA events inside the library, covt_timing.decode_ms), GB/s of values read and
of compressed bytes written, the end-to-end time of the call with host values (host->device copy of the values included), and the
CPU restatement of EncodingUtils (tools/gen, one thread) on a bounded sample as the baseline beside it. One JSON line on stdout."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def timed_encode(dec, abi, values, descs, steps, warmup):
    dev_ms, wall_ms = [], []
    last = None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = dec.encode_streams(values, descs)
        t1 = time.perf_counter()
        if it >= warmup:
            dev_ms.append(res.timing()["decode_ms"])
            wall_ms.append((t1 - t0) * 1e3)
        if last is not None:
            last.free()
        last = res
    return last, float(np.mean(dev_ms)), float(np.mean(wall_ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stream-bytes", type=int, default=1 << 30)
    ap.add_argument("--replicas", type=int, default=64)
    args = ap.parse_args()
    import covt_loader
    import util
    from tools.gen import gen as G
    covt = covt_loader.load()
    covt.build()
    abi = covt.abi
    dec = covt.Decoder(0)
    out = {"what": "GPU stream encoders (covt_encode_streams), 1 B200", "steps": args.steps, "warmup": args.warmup, "workloads": {}}

    # ---- config 3 backwards: values -> the 1 GiB stream
    enc, n = G.varint_stream(args.stream_bytes, seed=0xC0717)
    enc = np.asarray(enc, dtype=np.uint8)
    vals, st, cons = dec.decode_stream(np.concatenate([enc, np.zeros(64, np.uint8)]), abi.OP_VARINT_ZZ_DELTA_XY, byte_length=len(enc), num_values=n)
    assert st == 0 and cons == len(enc)
    descs = (abi.EncodeDesc * 1)()
    descs[0] = abi.EncodeDesc(value_offset=0, num_values=n, op=abi.OP_VARINT_ZZ_DELTA_XY)
    res, dev_ms, wall_ms = timed_encode(dec, abi, vals, descs, args.steps, args.warmup)
    got = res.buffer(abi.BUF_STREAM_ARENA, descs[0].out_offset, descs[0].byte_length)
    same = len(got) == len(enc) and bool(np.array_equal(got, enc))
    res.free()
    t0 = time.perf_counter()
    sample = min(n, 1 << 24) & ~1
    cpu = G.encode_varints(G.encode_zigzag_delta_coordinates(vals[:sample]).astype(np.int64) & 0xFFFFFFFF)
    cpu_s = time.perf_counter() - t0
    out["workloads"]["varint1g"] = {
        "values": int(n), "value_bytes": int(vals.nbytes), "compressed_bytes": int(len(enc)), "round_trip_identical": same,
        "device_ms": dev_ms, "values_GBps": vals.nbytes / dev_ms / 1e6, "compressed_GBps": len(enc) / dev_ms / 1e6,
        "e2e_ms_host_values": wall_ms,
        "cpu_port_1_thread": {"sample_values": int(sample), "seconds": cpu_s, "compressed_GBps": len(cpu) / cpu_s / 1e9},
    }
    log("[encode] varint1g: %.2f ms on the device (%.1f GB/s of values, %.1f GB/s compressed), call %.1f ms, identical=%s" % (
        dev_ms, vals.nbytes / dev_ms / 1e6, len(enc) / dev_ms / 1e6, wall_ms, same))
    del vals, got

    # ---- every fixture stream, x replicas
    fixtures = util.load_fixture_tiles()
    blob, offs = util.concat_tiles([b for _, b in fixtures])
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    r = dec.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    layers = r.layers
    bufs = {b: r.buffer(b) for b in set(abi.SLOT_BUF)}
    buf = bytearray()
    one = []
    for L in layers:
        key = "%s/%s" % (fixtures[int(L["tile"])][0], util.layer_name(blob, L))
        for s in range(abi.NUM_SLOTS):
            S = L["streams"][s]
            if S["encoding"] == abi.ENC_ABSENT or S["op"] == abi.OP_NONE or S["status"] != 0 or (key in util.KNOWN_MISLABELLED and s == abi.SLOT_VBUF):
                continue
            nv = int(S["num_values"])
            ice = s == abi.SLOT_VBUF and L["geom_column_type"] in (abi.CT_ICE, abi.CT_ICE_MORTON_CODE)
            b = abi.SLOT_BUF[s]
            o = int(L["out"][b])
            a = bufs[b][o:o + (2 * nv if ice else nv)]
            buf += bytes((-len(buf)) % 8)
            morton = int(S["op"]) in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON)
            nv = a.size // 2 if morton else a.size  # (ICE vertex buffers: the header counts vertices, the stream holds 2 ints each)
            one.append((len(buf), nv, int(S["op"]), int(L["num_bits"]), int(S["byte_offset"]), int(S["byte_length"])))
            buf += a.tobytes()
    r.free()
    buf += bytes((-len(buf)) % 8)
    stride = len(buf)
    values = np.tile(np.frombuffer(bytes(buf), np.uint8), args.replicas)
    descs = (abi.EncodeDesc * (len(one) * args.replicas))()
    k = 0
    for rep in range(args.replicas):
        for (vo, nv, op, nb, _, _) in one:
            descs[k] = abi.EncodeDesc(value_offset=rep * stride + vo, num_values=nv, op=op, num_bits=nb)
            k += 1
    res, dev_ms, wall_ms = timed_encode(dec, abi, values, descs, args.steps, args.warmup)
    arena = res.buffer(abi.BUF_STREAM_ARENA)
    comp = sum(int(d.byte_length) for d in descs)
    bad = 0
    for k, (vo, nv, op, nb, off, bl) in enumerate(one):  # the last replica
        d = descs[(args.replicas - 1) * len(one) + k]
        ok = d.byte_length == bl and np.array_equal(arena[d.out_offset:d.out_offset + d.byte_length], blob[off:off + bl])
        if not ok and not (abi.OP_NAMES[op].startswith("pfor") and nv > 65536):
            bad += 1
            if bad <= 5:
                log("[encode] differs: op %s, %d values, %d bytes vs %d in the fixture" % (abi.OP_NAMES[op], nv, d.byte_length, bl))
    res.free()
    t0 = time.perf_counter()
    cpu_bytes = 0
    for (vo, nv, op, nb, off, bl) in one[:2000]:
        name = abi.OP_NAMES[op]
        a = np.frombuffer(bytes(buf[vo:vo + 8 * nv + 8]), np.uint8)
        if name == "byte_rle":
            cpu_bytes += len(G.encode_byte_rle(a[:nv]))
        elif name.startswith("rle"):
            v = a[:4 * nv].view(np.int32).astype(np.int64) if name == "rle_u32" else a[:8 * nv].view(np.int64)
            cpu_bytes += len(G.encode_rle(v, signed=False))
        elif name == "pfor_zz_delta":
            cpu_bytes += len(G.encode_fastpfor(a[:4 * nv].view(np.int32), zigzag=True, delta=True))
        elif name == "varint_zz_delta":
            cpu_bytes += len(G.encode_varints(a[:4 * nv].view(np.int32).astype(np.int64), zigzag=True, delta=True))
    cpu_s = time.perf_counter() - t0
    out["workloads"]["fixtures"] = {
        "streams": len(descs), "replicas": args.replicas, "value_bytes": int(len(values)), "compressed_bytes": comp,
        "streams_differing_from_the_fixture_bytes": bad,
        "device_ms": dev_ms, "values_GBps": len(values) / dev_ms / 1e6, "compressed_GBps": comp / dev_ms / 1e6, "e2e_ms_host_values": wall_ms,
        "cpu_port_1_thread": {"sample_streams": min(2000, len(one)), "seconds": cpu_s, "compressed_GBps": cpu_bytes / max(cpu_s, 1e-9) / 1e9,
                              "note": "ctypes call overhead per stream included"},
    }
    log("[encode] fixtures x%d: %d streams, %.2f ms on the device (%.2f GB/s of values), call %.1f ms, %d streams differ" % (
        args.replicas, len(descs), dev_ms, len(values) / dev_ms / 1e6, wall_ms, bad))
    print(json.dumps(out))
    dec.close()


if __name__ == "__main__":
    main()
