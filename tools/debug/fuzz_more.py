"""Debug helper: many more seeds of the mutation fuzz tests than the test suite runs (prints every discrepancy, does not stop).
usage: python tools/debug/fuzz_more.py [first_seed] [n_seeds]"""
import importlib.util
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import covt_loader  # noqa: E402
import util  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools.gen import gen as G  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


tgs, tgb = load("test_gpu_streams"), load("test_gpu_batch")
covt = covt_loader.load()
covt.build()
dec = covt.Decoder(0)
abi = covt.abi
fixtures = util.load_fixture_tiles()
first, count = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 10)
bad = 0
for seed in range(first, first + count):
    try:
        tgs._batch_check(covt, O, dec, tgs._fuzz_cases(abi, G, seed))
    except AssertionError as e:
        bad += 1
        print("STREAM seed %d: %s" % (seed, str(e)[:300]))
    # batch fuzz: fixture mutants and synthetic-tile mutants, several flag sets
    clean = [(n, b) for n, b in fixtures if n.startswith("omt/") and not n.startswith("omt/8_")
             and not any(k.startswith(n + "/") for k in util.KNOWN_MISLABELLED)]
    small = [b for _, b in sorted(clean, key=lambda t: len(t[1]))[5 + seed % 20: 10 + seed % 20]]
    synth_blob, synth_offs, _ = G.tiles(seed * 100, 6, G.default_params(container=seed % 3))
    synth = [bytes(synth_blob[int(synth_offs[i]):int(synth_offs[i + 1])]) for i in range(6)]
    for label, base, container, nf in (("fixtures", small, abi.CONTAINER_GEN2B, None),
                                       ("synthetic", synth, abi.CONTAINER_GEN2B if seed % 3 == 0 else abi.CONTAINER_GEN3,
                                        [0] * 16 if seed % 3 == 2 else None)):
        quirk = abi.FLAG_ID_DVZZ_IS_RLE if label == "fixtures" else 0  # the fixtures' mislabelled id streams
        for flags in (abi.FLAG_CLOSE_RINGS | quirk, quirk | abi.FLAG_MORTON_NO_SHIFT, abi.FLAG_CLOSE_RINGS | quirk | abi.FLAG_SKIP_ASSEMBLY,
                      abi.FLAG_CLOSE_RINGS | quirk | abi.FLAG_ID_WIDTH_32):
            tiles, good = tgb._mutants(np.random.default_rng(seed), base, 600, 50)
            blob, offs = util.concat_tiles(tiles)
            res = dec.decode_batch(blob, offs, container, flags, n_fields=nf)
            ref = O.decode_batch(blob, offs, container, flags, n_fields=nf)
            try:
                st, first_layer = res.tile_status()
                assert np.array_equal(first_layer, ref.first_layer), "first_layer differs"
                assert np.array_equal(st == 0, ref.tile_status == 0), "tile status OK-ness differs"
                # (with ID_WIDTH_32 the 64-bit ids of good tiles are legitimately flagged as overlong)
                assert (flags & abi.FLAG_ID_WIDTH_32) or not st[good].any(), "a good tile was poisoned"
                util.compare_results(abi, res, ref)
            except AssertionError as e:
                bad += 1
                print("BATCH %s seed %d flags %#x: %s" % (label, seed, flags, str(e)[:300]))
                gl, wl = res.layers, ref.layers
                if len(gl) == len(wl):
                    diff = np.nonzero(((gl["status"] == 0) != (wl["status"] == 0)) |
                                      ((gl["streams"]["status"] == 0) != (wl["streams"]["status"] == 0)).any(axis=1))[0]
                    for li in diff[:2]:
                        t = int(wl["tile"][li])
                        print("   layer %d tile %d: gpu layer status %d oracle %d" % (li, t, gl["status"][li], wl["status"][li]))
                        for sl in range(abi.NUM_SLOTS):
                            S, W = gl["streams"][li][sl], wl["streams"][li][sl]
                            if W["encoding"] != abi.ENC_ABSENT:
                                print("      slot %s op %s nv %d bl %d off %d: gpu %d oracle %d" % (
                                    abi.SLOT_NAMES[sl], abi.OP_NAMES[W["op"]], W["num_values"], W["byte_length"],
                                    int(W["byte_offset"]) - int(offs[t]), S["status"], W["status"]))
                        open(os.path.join(ROOT, "gpurun_out", "fuzz_tile_%d_%s_%x.bin" % (seed, label, flags)), "wb").write(
                            bytes(blob[int(offs[t]):int(offs[t + 1])]))
            res.free()
    # property columns: mutants of fixture tiles (and of their gen-3 re-wraps) that keep their property columns
    pbase = [b for _, b in sorted([(n, b) for n, b in fixtures if n.startswith(("omt/", "bing/", "amazon/")) and not n.startswith("omt/8_")],
                                  key=lambda t: len(t[1]))[seed % 30: seed % 30 + 6]]
    for label, base, container in (("props gen-2b", pbase, abi.CONTAINER_GEN2B),
                                   ("props gen-3", [util.rewrap_gen3(abi, O, b, props=True, gen=G)[0] for b in pbase], abi.CONTAINER_GEN3)):
        flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
        tiles, good = tgb._mutants(np.random.default_rng(seed), base, 800, 40)
        blob, offs = util.concat_tiles(tiles)
        res = dec.decode_batch(blob, offs, container, flags | abi.FLAG_DECODE_PROPERTIES)
        try:
            util.compare_props(abi, blob, util.GpuProps(abi, res), O.decode_properties(blob, offs, container, flags))
            util.compare_results(abi, res, O.decode_batch(blob, offs, container, flags))
        except AssertionError as e:
            bad += 1
            print("%s seed %d: %s" % (label, seed, str(e)[:300]))
        res.free()
    print("seed %d done" % seed, flush=True)
print("discrepancies:", bad)
