import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import covt_loader
covt = covt_loader.load(); abi = covt.abi
import util
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tiles = [b for n, b in util.load_fixture_tiles() if n.startswith("omt/")]
blob1, offs1 = util.concat_tiles(tiles)
blob = np.tile(blob1, reps); sizes = np.tile(np.diff(offs1), reps)
offs = np.zeros(len(sizes) + 1, dtype=np.uint64); offs[1:] = np.cumsum(sizes)
n = len(sizes)
dec = covt.Decoder(0)
pinned = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True); pinned.numpy()[:] = blob
op = torch.empty(len(offs), dtype=torch.int64, pin_memory=True); op.numpy().view(np.uint64)[:] = offs
flags = abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = dec.decode_batch_raw(pinned.data_ptr(), op.data_ptr(), n, abi.CONTAINER_GEN2B, flags)
    t1 = time.perf_counter()
    r.touch_tile_status(); t2 = time.perf_counter()
    t = r.timing()
    r.free(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("iter", it, "call %.1f ms status %.1f ms free %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items()})
