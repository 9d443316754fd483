"""CPU: the C-ABI shared library loads, exports every symbol include/covt_b200.h declares, agrees with the ctypes
mirror on struct layout, fails loudly without a GPU (no CPU fallback), and its host-only entry points
(covt_resolve_op = the dispatch table of CovtParser.decodeGeometryColumn; covt_partition_tiles) are correct.
No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "covt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(covt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(covt):
    covt.build()
    names = _declared_functions()
    assert len(names) >= 23
    lib = covt.lib()
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(covt.ABI_SYMBOLS) == names, "ABI_SYMBOLS and the header disagree"
    assert lib.covt_abi_version() == covt.abi.ABI_VERSION
    # the product library carries sm_100a code only and does not link the oracle
    so = os.path.join(ROOT, "cov-tiles_b200", "libcovt_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "covt_oracle" not in ldd


def test_struct_layout_matches_header(covt, tmp_path):
    """sizeof/offsetof as the C compiler sees include/covt_b200.h == the ctypes mirror (what Panama FFM would lay out)."""
    abi = covt.abi
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "covt_b200.h"\nint main(void){\n'
                    'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(covt_stream_ref), sizeof(covt_layer), sizeof(covt_stream_desc),'
                    ' sizeof(covt_timing), sizeof(covt_kernel_time), sizeof(covt_tilejson));\n'
                    'printf("%zu %zu %zu %zu\\n", offsetof(covt_layer, streams), offsetof(covt_layer, out), offsetof(covt_layer, n_parts),'
                    ' offsetof(covt_stream_desc, out_offset));\n'
                    'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(covt_prop_column), sizeof(covt_prop_dictionary), offsetof(covt_prop_column, status),'
                    ' offsetof(covt_prop_column, validity_offset), offsetof(covt_prop_column, data_num_values), offsetof(covt_prop_dictionary, offsets_offset),'
                    ' offsetof(covt_layer, header_offset));\n'
                    'printf("%zu %zu %zu %zu\\n", sizeof(covt_encode_desc), offsetof(covt_encode_desc, op), offsetof(covt_encode_desc, out_offset),'
                    ' offsetof(covt_encode_desc, status));\nreturn 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    a, b, c, d = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    assert [int(x) for x in d.split()] == [C.sizeof(abi.EncodeDesc), abi.EncodeDesc.op.offset, abi.EncodeDesc.out_offset.offset,
                                            abi.EncodeDesc.status.offset]
    assert [int(x) for x in c.split()] == [C.sizeof(abi.PropColumn), C.sizeof(abi.PropDictionary), abi.PropColumn.status.offset,
                                            abi.PropColumn.validity_offset.offset, abi.PropColumn.data_num_values.offset,
                                            abi.PropDictionary.offsets_offset.offset, abi.Layer.header_offset.offset]
    assert [int(x) for x in a.split()] == [C.sizeof(abi.StreamRef), C.sizeof(abi.Layer), C.sizeof(abi.StreamDesc),
                                            C.sizeof(abi.Timing), C.sizeof(abi.KernelTime), C.sizeof(abi.TileJson)]
    assert [int(x) for x in b.split()] == [abi.Layer.streams.offset, abi.Layer.out.offset, abi.Layer.n_parts.offset,
                                            abi.StreamDesc.out_offset.offset]


def test_no_cpu_fallback(covt):
    """Without a CUDA device every entry point fails with COVT_ERR_CUDA and says so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = covt.lib().covt_create(0, C.byref(h))
    assert rc == covt.abi.ERR_CUDA and not h.value
    buf = C.create_string_buffer(256)
    covt.lib().covt_last_error(None, buf, 256)
    assert b"CUDA" in buf.value
    with pytest.raises(covt.CovtError):
        covt.Decoder(0)
    with pytest.raises(covt.CovtError):
        covt.DecodingUtils.decodeVarint(np.array([10], np.uint8), [0], 1)


def test_dispatch_table(covt, oracle):
    """covt_resolve_op reproduces CovtParser.decodeGeometryColumn :405-510 / decodedIds :552-572
    and agrees with the oracle for every (stream type, encoding, column type, flag) combination."""
    abi = covt.abi
    lib = covt.lib()
    for flags in (0, abi.FLAG_ID_DVZZ_IS_RLE, abi.FLAG_ID_WIDTH_32, abi.FLAG_ID_DVZZ_IS_RLE | abi.FLAG_ID_WIDTH_32):
        for st in range(13):
            for enc in range(10):
                for ct in range(5):
                    assert lib.covt_resolve_op(st, enc, ct, flags) == oracle.lib().covt_oracle_resolve_op(st, enc, ct, flags), (st, enc, ct, flags)
    r = lib.covt_resolve_op
    assert r(abi.ST_GEOMETRY_TYPES, abi.ENC_PLAIN, 0, 0) == abi.OP_BYTE_RLE  # any label, always Byte-RLE (:405-406)
    for st in (abi.ST_GEOMETRY_OFFSETS, abi.ST_PART_OFFSETS, abi.ST_RING_OFFSETS):
        assert r(st, abi.ENC_RLE, 0, 0) == abi.OP_RLE_U32
        assert r(st, abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, 0, 0) == abi.OP_PFOR_ZZ_DELTA
        assert r(st, abi.ENC_VARINT, 0, 0) == abi.OP_NONE  # IllegalArgumentException (:425-427)
    assert r(abi.ST_VERTEX_BUFFER, abi.ENC_VARINT_DELTA_ZIG_ZAG, abi.CT_ICE_MORTON_CODE, 0) == abi.OP_VARINT_DELTA_MORTON
    assert r(abi.ST_VERTEX_BUFFER, abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, abi.CT_ICE_MORTON_CODE, 0) == abi.OP_PFOR_DELTA_MORTON
    assert r(abi.ST_VERTEX_BUFFER, abi.ENC_VARINT_DELTA_ZIG_ZAG, abi.CT_PLAIN, 0) == abi.OP_VARINT_ZZ_DELTA_XY
    assert r(abi.ST_VERTEX_BUFFER, abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG, abi.CT_ICE, 0) == abi.OP_PFOR_ZZ_DELTA_XY
    assert r(abi.ST_VERTEX_OFFSETS, abi.ENC_VARINT_DELTA_ZIG_ZAG, abi.CT_ICE, 0) == abi.OP_VARINT_ZZ_DELTA
    assert r(abi.ST_DATA, abi.ENC_RLE, 0, 0) == abi.OP_RLE_U64
    assert r(abi.ST_DATA, abi.ENC_VARINT, 0, abi.FLAG_ID_WIDTH_32) == abi.OP_VARINT_U32_AS_I64
    assert r(abi.ST_DATA, abi.ENC_VARINT_DELTA_ZIG_ZAG, 0, abi.FLAG_ID_DVZZ_IS_RLE) == abi.OP_RLE_U64


def test_partition_tiles(covt):
    rng = np.random.default_rng(3)
    sizes = rng.integers(0, 9000, 10000).astype(np.uint64)
    offs = np.zeros(len(sizes) + 1, np.uint64)
    offs[1:] = np.cumsum(sizes)
    for parts in (1, 2, 3, 4, 8, 16):
        s = covt.partition_tiles(offs, parts)
        assert s[0] == 0 and s[-1] == len(sizes) and np.all(np.diff(s.astype(np.int64)) >= 0)
        share = np.diff(offs[s].astype(np.float64))
        assert share.max() - share.min() <= 2 * 9000, "ranges are balanced by payload bytes to within one tile"
    # degenerate inputs: more parts than tiles, empty tiles, empty batch
    s = covt.partition_tiles(np.array([0, 5, 5, 9], np.uint64), 8)
    assert s[0] == 0 and s[-1] == 3 and np.all(np.diff(s.astype(np.int64)) >= 0)
    s = covt.partition_tiles(np.array([0], np.uint64), 4)
    assert list(s) == [0, 0, 0, 0, 0]
    # the definition, against a plain restatement: part p starts at the first tile whose start offset reaches p / n_parts of the bytes
    # (many empty tiles = runs of equal offsets; a first offset that is not 0)
    for trial in range(20):
        sizes = rng.integers(0, 50, 3000) * (rng.random(3000) < 0.4)
        offs = np.zeros(len(sizes) + 1, np.uint64)
        offs[1:] = np.cumsum(sizes)
        offs += np.uint64(trial * 1000)
        for parts in (2, 5, 24, 64):
            base, total = int(offs[0]), int(offs[-1] - offs[0])
            want = [0] + [int(np.searchsorted(offs[:-1], base + total * p // parts, side="left")) for p in range(1, parts)] + [len(sizes)]
            want = list(np.maximum.accumulate(want))
            assert list(covt.partition_tiles(offs, parts)) == want


def test_java_binding_names_every_symbol_it_binds(covt):
    """integration/java/CovtGpuDecoder.java (un-built: no JDK here) binds only symbols the library exports."""
    src = open(os.path.join(ROOT, "integration", "java", "CovtGpuDecoder.java")).read()
    bound = set(re.findall(r'h\("(covt_[a-z0-9_]+)"', src))
    assert len(bound) >= 10
    for name in bound:
        assert hasattr(covt.lib(), name), name
    assert "sizeof(covt_layer)" in src and str(C.sizeof(covt.abi.Layer)) in src
