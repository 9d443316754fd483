import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import covt_loader
covt = covt_loader.load(); abi = covt.abi
from tools.gen import gen as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dec = covt.Decoder(0)
blob, offs, truth = G.tiles(0, n, G.default_params())
pinned = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True); pinned.numpy()[:] = blob
op = torch.empty(len(offs), dtype=torch.int64, pin_memory=True); op.numpy().view(np.uint64)[:] = offs
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = dec.decode_batch_raw(pinned.data_ptr(), op.data_ptr(), n, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    t1 = time.perf_counter()
    r.touch_tile_status(); t2 = time.perf_counter()
    t = r.timing()
    r.free(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("iter", it, "call %.1f ms status %.1f ms free %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items()})
