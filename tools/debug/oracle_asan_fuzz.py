"""The oracle is the arbiter of every parity test, so it must itself be memory-safe on the corrupt inputs of the fuzz tests:
runs the mutation corpora of tests/test_gpu_batch.py / test_gpu_streams.py (CPU side only) and the property-column decode
through the AddressSanitizer + UBSan build of oracle/covt_oracle.c.

  COVT_ORACLE_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \\
      python tools/debug/oracle_asan_fuzz.py [first_seed] [n_seeds]
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
assert os.environ.get("COVT_ORACLE_ASAN"), "set COVT_ORACLE_ASAN=1 (and LD_PRELOAD libasan)"
import util  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools.gen import gen as G  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


tgs, tgb = load("test_gpu_streams"), load("test_gpu_batch")
abi = O.abi
fixtures = util.load_fixture_tiles()
first, count = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 6)
clean = [(n, b) for n, b in fixtures if n.startswith("omt/") and not n.startswith("omt/8_")]
n_tiles = n_streams = 0
for seed in range(first, first + count):
    for op, enc, n, nbits, exact in tgs._fuzz_cases(abi, G, seed):
        blob = np.concatenate([np.asarray(enc, np.uint8), np.zeros(64, np.uint8)])
        O.decode_stream(blob, op, byte_offset=0, byte_length=len(enc) + (0 if exact else 37), num_values=n, num_bits=nbits)
        n_streams += 1
    small = [b for _, b in sorted(clean, key=lambda t: len(t[1]))[seed % 20: seed % 20 + 5]]
    synth_blob, synth_offs, _ = G.tiles(seed * 100, 6, G.default_params(container=seed % 3))
    synth = [bytes(synth_blob[int(synth_offs[i]):int(synth_offs[i + 1])]) for i in range(6)]
    for base, container, nf in ((small, abi.CONTAINER_GEN2B, None),
                                (synth, abi.CONTAINER_GEN2B if seed % 3 == 0 else abi.CONTAINER_GEN3, [0] * 16 if seed % 3 == 2 else None)):
        tiles, _ = tgb._mutants(np.random.default_rng(seed), base, 600, 50)
        blob, offs = util.concat_tiles(tiles)
        for flags in (abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE, abi.FLAG_MORTON_NO_SHIFT | abi.FLAG_ID_WIDTH_32, abi.FLAG_SKIP_ASSEMBLY):
            O.decode_batch(blob, offs, container, flags, n_fields=nf)
        if container == abi.CONTAINER_GEN2B:
            O.decode_properties(blob, offs)  # property walk of corrupt tiles: statuses, never a crash
        n_tiles += len(tiles)
    print("seed %d done" % seed, flush=True)
print("oracle survived %d mutated tiles x 3 flag sets and %d mutated streams under ASan/UBSan" % (n_tiles, n_streams))
