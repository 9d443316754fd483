"""Debug helper: run selected cases of tests/test_gpu_streams.py::_fuzz_cases alone and inside their batch (GPU vs oracle)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import covt_loader  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools.gen import gen as G  # noqa: E402

spec = importlib.util.spec_from_file_location("tgs", os.path.join(ROOT, "tests", "test_gpu_streams.py"))
tgs = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tgs)

covt = covt_loader.load()
covt.build()
dec = covt.Decoder(0)
abi = covt.abi
seed, which = int(sys.argv[1]), [int(x) for x in sys.argv[2:]]
cases = tgs._fuzz_cases(abi, G, seed)
for w in which:
    for lo, hi in ((w, w + 1), (max(0, w - 3), w + 3), (0, len(cases))):
        sub = cases[lo:hi]
        blob = bytearray()
        descs = (abi.StreamDesc * len(sub))()
        for i, (op, payload, n, nbits, exact) in enumerate(sub):
            blob += bytes((i * 7 + 3) % 16 + (1 if i % 3 == 0 else 0))
            off = len(blob)
            blob += bytes(payload)
            slack = 0 if exact else 37
            descs[i] = abi.StreamDesc(byte_offset=off, byte_length=len(payload) + slack, num_values=n, num_bits=nbits, op=op)
            if not exact:
                blob += bytes([0x80] * 3) + bytes(slack - 3)
        blob += bytes(64)
        blob = np.frombuffer(bytes(blob), dtype=np.uint8)
        res = dec.decode_streams(blob, descs, abi.FLAG_DEFAULT)
        d = descs[w - lo]
        want, wst, wcons = O.decode_stream(blob, d.op, byte_offset=d.byte_offset, byte_length=d.byte_length, num_values=d.num_values,
                                           num_bits=d.num_bits, flags=abi.FLAG_DEFAULT)
        print("case %d in [%d,%d): op %s off %d (mod 16 = %d) len %d n %d | gpu status %d consumed %d count %d | oracle status %d consumed %d" % (
            w, lo, hi, abi.OP_NAMES[d.op], d.byte_offset, d.byte_offset % 16, d.byte_length, d.num_values, d.status, d.bytes_consumed,
            d.out_count, wst, wcons))
        res.free()
