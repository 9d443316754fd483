import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import covt_loader
covt = covt_loader.load(); abi = covt.abi
from oracle import oracle as O
from tools.gen import gen as G
dec = covt.Decoder(0)
rng = np.random.default_rng(1)
for n in (70000, 140000):
  for align in (0, 3, 13):
    v = rng.integers(-(1 << 26), 1 << 26, n).astype(np.int64)
    for op, enc in ((abi.OP_VARINT_ZZ_DELTA, G.encode_varints(v, zigzag=True, delta=True)),
                    (abi.OP_VARINT_ZZ_DELTA_XY, G.encode_varints(G.encode_zigzag_delta_coordinates(v.astype(np.int32)).astype(np.int64) & 0xFFFFFFFF))):
        blob = np.concatenate([np.zeros(align, np.uint8), enc, np.zeros(64, np.uint8)])
        got, st, cons = dec.decode_stream(blob, op, byte_offset=align, byte_length=len(enc), num_values=n)
        want, wst, wcons = O.decode_stream(blob, op, byte_offset=align, byte_length=len(enc), num_values=n)
        bad = np.nonzero(got != want)[0]
        # value index of chunk boundaries
        ends = np.nonzero((enc & 0x80) == 0)[0]  # terminator byte positions
        print("n", n, "align", align, "op", abi.OP_NAMES[op], "bytes", len(enc), "st", st, wst, "cons", cons, wcons, "nbad", len(bad), "first bad", bad[:3])
        if len(bad):
            b = bad[0]
            pos = ends[b] + align
            print("   first bad value ends at window byte", pos, "chunk", pos // 512, "lane", (pos % 512) // 16, "byte", pos % 16, "starts at", (ends[b-1] + 1 + align) if b else align)
            print("   got-want", (got[bad[:4]].astype(np.int64) - want[bad[:4]]), "want[b-1]", want[b-1] if b else None, "vals", want[b], got[b])
