"""The C ABI from plain C (examples/decode_tiles.c): no Python, no torch between the caller and libcovt_b200.so.

CPU: the example compiles against include/covt_b200.h, links the library and FAILS LOUDLY without a GPU (no CPU fallback).
GPU: its per-layer output (counts, status, FNV-1a of the assembled coordinates) equals the oracle's on fixture tiles."""
import os
import subprocess

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def example_binary(tmp_path_factory, covt):
    covt.build()
    out = str(tmp_path_factory.mktemp("c_example") / "decode_tiles")
    libdir = os.path.join(ROOT, "cov-tiles_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "decode_tiles.c"), "-o", out,
                           "-L", libdir, "-lcovt_b200", "-Wl,-rpath," + libdir])
    return out


def test_c_example_builds_and_refuses_to_run_without_a_gpu(example_binary, tmp_path):
    tile = tmp_path / "t.covt"
    tile.write_bytes(dict(util.load_fixture_tiles())["omt/5_16_21"])
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([example_binary, str(tile)], env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "covt_create failed" in p.stderr and p.stdout == ""


def _fnv(ints):
    h = 1469598103934665603
    for v in ints.astype(np.uint32).tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_c_example_output_equals_oracle(example_binary, oracle, fixtures, tmp_path):
    abi = oracle.abi
    names = ["omt/5_16_21", "omt/5_16_20", "omt/6_33_42", "omt/7_66_84", "omt/4_8_10"]
    have = dict(fixtures)
    names = [n for n in names if n in have] or [n for n, _ in fixtures if n.startswith("omt/5_")]
    paths = []
    for n in names:
        f = tmp_path / (n.replace("/", "_") + ".covt")
        f.write_bytes(have[n])
        paths.append(str(f))
    flags = abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE
    p = subprocess.run([example_binary, "--flags", str(flags)] + paths, capture_output=True, text=True)
    assert p.returncode in (0, 3), p.stderr
    blob, offs = util.concat_tiles([have[n] for n in names])
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
    lines = [ln.split(" ") for ln in p.stdout.strip().split("\n")]
    assert len(lines) == len(ref.layers) > 20
    coords = ref.buffer(abi.BUF_A_COORDS)
    checked = 0
    for ln, L in zip(lines, ref.layers):
        tile, li, name = int(ln[0]), int(ln[1]), ln[2]
        feats, parts, rings, verts, ncoords, status = (int(x) for x in ln[3:9])
        assert (tile, li, name) == (int(L["tile"]), int(L["layer_index"]), util.layer_name(blob, L))
        assert feats == int(L["num_features"]) and (status == 0) == (int(L["status"]) == 0)
        if status != 0:
            continue
        assert (parts, rings, verts, ncoords) == (int(L["n_parts"]), int(L["n_rings"]), int(L["n_vertices"]), int(L["n_coords"]))
        o = int(L["out"][abi.BUF_A_COORDS])
        assert int(ln[9], 16) == _fnv(coords[o:o + 2 * ncoords]), name
        checked += 1
    assert checked > 20
    # the example also encodes every decoded geometry_types stream again on the GPU (covt_encode_streams from plain C)
    import re
    m = re.search(r"re-encoded (\d+) geometry_types streams on the GPU: (\d+) identical", p.stderr)
    assert m and int(m.group(1)) == checked and m.group(1) == m.group(2), p.stderr
