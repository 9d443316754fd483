"""The tile writer's host logic without a GPU (cov-tiles_b200/converter.py): stream selection ("encode both ways, keep the shorter"),
ICE Morton dictionary building, gen-2b / gen-3 metadata and payload order — with the stream ENCODERS swapped for the CPU restatement
of EncodingUtils (tools/gen), the tiles must equal those of the CPU restatement of the converter (covt_gen_append_layer) byte for
byte. tests/test_gpu_converter.py runs the same comparison with the real GPU encoders."""
import sys

import numpy as np

import covt_loader
from test_gpu_converter import OPTION_SETS, random_layer


def _cpu_plan_run(abi, gen):
    def run(self):
        self.out = []
        raw = bytes(self.buf) + bytes(8)
        for off, n, op, nb in self.req:
            name = abi.OP_NAMES[op]
            if name == "byte_rle":
                e = gen.encode_byte_rle(np.frombuffer(raw, np.uint8, n, off))
            elif name == "rle_u32":
                e = gen.encode_rle(np.frombuffer(raw, np.int32, n, off).astype(np.int64))
            elif name == "rle_u64":
                e = gen.encode_rle(np.frombuffer(raw, np.int64, n, off))
            elif name == "pfor_zz_delta":
                e = gen.encode_fastpfor(np.frombuffer(raw, np.int32, n, off), True, True)
            elif name == "pfor_zz_delta_xy":
                e = gen.encode_fastpfor(gen.encode_zigzag_delta_coordinates(np.frombuffer(raw, np.int32, n, off)), False, False)
            elif name == "varint_zz_delta_xy":
                e = gen.encode_varints(gen.encode_zigzag_delta_coordinates(np.frombuffer(raw, np.int32, n, off)).astype(np.int64) & 0xFFFFFFFF)
            elif name == "varint_zz_delta":
                e = gen.encode_varints(np.frombuffer(raw, np.int32, n, off).astype(np.int64), True, True)
            elif name == "varint_u64":
                e = gen.encode_varints(np.frombuffer(raw, np.int64, n, off))
            elif name == "varint_zz_delta_64":
                e = gen.encode_varints(np.frombuffer(raw, np.int64, n, off), True, True)
            elif name in ("varint_delta_morton", "pfor_delta_morton"):
                v = np.frombuffer(raw, np.int32, 2 * n, off).reshape(-1, 2)
                codes = np.array([gen.encode_morton(int(a), int(b), nb) for a, b in v], dtype=np.int64)
                e = gen.encode_varints(codes, delta=True) if name.startswith("varint") else gen.encode_fastpfor(codes.astype(np.int32), False, True)
            else:
                raise AssertionError(name)
            self.out.append(bytes(e))
    return run


def test_converter_host_logic_equals_the_reference_converter_restatement(gen, monkeypatch):
    covt = covt_loader.load()
    abi = covt.abi
    conv_mod = sys.modules[covt.__name__ + ".converter"]
    monkeypatch.setattr(conv_mod._Plan, "run", _cpu_plan_run(abi, gen))
    conv = covt.CovtConverter(None)
    rng = np.random.default_rng(31)
    n_tiles = 0
    for container in (0, 1):
        tiles = []
        for k, options in enumerate(OPTION_SETS * 2):
            layers = [random_layer(rng, int(n), name="l%d" % i, extent=int(rng.choice([4096, 8192])), with_ids=bool((k + i) % 3),
                                   id_kind=["seq", "big", "walk"][(k + i) % 3], step=int(rng.choice([3, 60, 900])))
                      for i, n in enumerate(rng.choice([0, 1, 3, 40, 300], size=int(rng.integers(1, 4))))]
            for L in layers:
                L["options"] = options
            tiles.append(layers)
        for layers, got in zip(tiles, conv.convert_tiles(tiles, container)):
            assert got == bytes(gen.make_tile(layers, container, layers[0]["options"])), (container, hex(layers[0]["options"]))
            n_tiles += 1
    assert n_tiles == 4 * len(OPTION_SETS)
