"""ctypes binding of the CPU oracle (oracle/libcovt_oracle.so).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _load_abi():
    spec = importlib.util.spec_from_file_location("covt_abi", os.path.join(_ROOT, "cov-tiles_b200", "abi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


abi = _load_abi()


def build(force=False):
    if os.environ.get("COVT_ORACLE_ASAN"):
        # the AddressSanitizer / UBSan build of the oracle (tools/debug/oracle_asan_fuzz.sh; needs LD_PRELOAD of libasan)
        subprocess.check_call(["make", "-C", _HERE, "asan"], stdout=subprocess.DEVNULL)
        return os.path.join(_HERE, "libcovt_oracle_asan.so")
    so = os.path.join(_HERE, "libcovt_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("covt_oracle.c", "covt_oracle.h")] + [os.path.join(_ROOT, "include", "covt_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libcovt_oracle.so"], stdout=subprocess.DEVNULL)
    return so


class OracleResult(C.Structure):
    _fields_ = [("n_tiles", C.c_uint32), ("n_layers", C.c_uint32), ("tile_status", C.POINTER(C.c_uint32)),
                ("first_layer", C.POINTER(C.c_uint32)), ("layers", C.POINTER(abi.Layer)),
                ("buffers", C.c_void_p * abi.NUM_BUFFERS), ("counts", C.c_uint64 * abi.NUM_BUFFERS),
                ("payload_bytes", C.c_uint64), ("vertices", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        _lib.covt_oracle_decode_stream.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(abi.StreamDesc), C.c_uint32, C.c_void_p, C.c_uint64]
        _lib.covt_oracle_decode_stream.restype = C.c_int32
        _lib.covt_oracle_resolve_op.argtypes = [C.c_uint32] * 4
        _lib.covt_oracle_resolve_op.restype = C.c_int32
        _lib.covt_oracle_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(abi.TileJson), C.c_uint32, C.c_uint32, C.POINTER(C.POINTER(OracleResult))]
        _lib.covt_oracle_decode_batch.restype = C.c_int32
        _lib.covt_oracle_decode_batch_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(abi.TileJson), C.c_uint32, C.c_uint32, u64p, u64p, u64p]
        _lib.covt_oracle_decode_batch_timed.restype = C.c_int32
        _lib.covt_oracle_result_free.argtypes = [C.POINTER(OracleResult)]
        _lib.covt_oracle_result_free.restype = None
        _lib.covt_oracle_parse_tile.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(abi.TileJson), C.c_uint32, C.c_uint32, C.POINTER(abi.Layer), C.c_uint32, C.POINTER(C.c_uint32), u64p]
        _lib.covt_oracle_parse_tile.restype = C.c_int32
        _lib.covt_oracle_decode_morton.argtypes = [C.c_int32, C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib.covt_oracle_decode_morton.restype = None
        del u8p
    return _lib


def _as_u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def make_tilejson(n_fields):
    """n_fields: list of field counts per vector layer (or None)."""
    if n_fields is None:
        return None, None
    arr = (C.c_uint32 * max(1, len(n_fields)))(*n_fields)
    tj = abi.TileJson(len(n_fields), C.cast(arr, C.POINTER(C.c_uint32)))
    return tj, arr


def decode_stream(blob, op=0, *, byte_offset=0, byte_length=None, num_values, stream_type=0, encoding=0,
                  column_type=0, num_bits=0, flags=abi.FLAG_DEFAULT):
    """One DecodingUtils call. Returns (values ndarray, status, bytes_consumed)."""
    b = _as_u8(blob)
    if byte_length is None:
        byte_length = len(b) - byte_offset
    d = abi.StreamDesc(byte_offset=byte_offset, byte_length=byte_length, num_values=num_values, stream_type=stream_type,
                       encoding=encoding, column_type=column_type, num_bits=num_bits, op=op)
    rop = op or lib().covt_oracle_resolve_op(stream_type, encoding, column_type, flags)
    mult = 2 if rop in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON) else 1
    out = np.zeros(num_values * mult + 4, dtype=abi.op_dtype(rop))
    rc = lib().covt_oracle_decode_stream(b.ctypes.data, len(b), C.byref(d), flags, out.ctypes.data, out.nbytes)
    if rc != 0:
        raise RuntimeError("covt_oracle_decode_stream failed: %d" % rc)
    return out[: d.out_count].copy(), d.status, d.bytes_consumed


def _copy_from(ptr, nbytes, dtype):
    """numpy copy of `nbytes` at a ctypes pointer (C.string_at stops at 2 GiB; the 1 M-tile result buffers are larger)."""
    addr = C.cast(ptr, C.c_void_p).value
    return np.frombuffer((C.c_uint8 * nbytes).from_address(addr), dtype=dtype).copy()


class BatchResult:
    """Owns a covt_oracle_result; exposes numpy views with the same layout as the product's result."""

    def __init__(self, ptr):
        self._ptr = ptr
        r = ptr.contents
        self.n_tiles, self.n_layers = r.n_tiles, r.n_layers
        self.payload_bytes, self.vertices = r.payload_bytes, r.vertices
        self.tile_status = np.ctypeslib.as_array(r.tile_status, shape=(max(1, r.n_tiles),))[: r.n_tiles].copy()
        self.first_layer = np.ctypeslib.as_array(r.first_layer, shape=(r.n_tiles + 1,)).copy()
        if r.n_layers:
            self.layers = _copy_from(r.layers, r.n_layers * C.sizeof(abi.Layer), abi.LAYER_DTYPE)
        else:
            self.layers = np.zeros(0, dtype=abi.LAYER_DTYPE)
        self.buffers = []
        for b in range(abi.NUM_BUFFERS):
            n = r.counts[b]
            dt = np.dtype(abi.BUF_DTYPES[b])
            if n and r.buffers[b]:
                arr = _copy_from(r.buffers[b], n * dt.itemsize, dt)
            else:
                arr = np.zeros(0, dtype=dt)
            self.buffers.append(arr)
        lib().covt_oracle_result_free(ptr)
        self._ptr = None

    def buffer(self, which):
        return self.buffers[which]


def decode_batch(blob, tile_offsets, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None, n_threads=0):
    b = _as_u8(blob)
    offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
    tj, keep = make_tilejson(n_fields)
    out = C.POINTER(OracleResult)()
    rc = lib().covt_oracle_decode_batch(b.ctypes.data, offs.ctypes.data, len(offs) - 1, container,
                                        C.byref(tj) if tj is not None else None, flags, n_threads, C.byref(out))
    if rc != 0:
        raise RuntimeError("covt_oracle_decode_batch failed: %d" % rc)
    return BatchResult(out)


def decode_batch_timed(blob, tile_offsets, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None, n_threads=0):
    """The CPU-baseline loop. Returns (status, payload_bytes, vertices, checksum)."""
    b = _as_u8(blob)
    offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
    tj, keep = make_tilejson(n_fields)
    pb, vx, cs = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib().covt_oracle_decode_batch_timed(b.ctypes.data, offs.ctypes.data, len(offs) - 1, container,
                                              C.byref(tj) if tj is not None else None, flags, n_threads,
                                              C.byref(pb), C.byref(vx), C.byref(cs))
    return rc, pb.value, vx.value, cs.value


def parse_tile(blob, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None, begin=0, end=None):
    b = _as_u8(blob)
    if end is None:
        end = len(b)
    cap = 4096
    layers = (abi.Layer * cap)()
    n = C.c_uint32()
    ep = C.c_uint64()
    tj, keep = make_tilejson(n_fields)
    rc = lib().covt_oracle_parse_tile(b.ctypes.data, begin, end, container, C.byref(tj) if tj is not None else None,
                                      flags, 0, layers, cap, C.byref(n), C.byref(ep))
    arr = np.frombuffer(C.string_at(layers, n.value * C.sizeof(abi.Layer)), dtype=abi.LAYER_DTYPE).copy()
    return rc, arr, ep.value


def decode_morton(code, num_bits, no_shift=False):
    x, y = C.c_int32(), C.c_int32()
    lib().covt_oracle_decode_morton(code, num_bits, int(no_shift), C.byref(x), C.byref(y))
    return x.value, y.value


# ---- property columns (covt_oracle_decode_properties: SURVEY 8 f1) -----------------------------------------------------------
PropColumn, PropDictionary = abi.PropColumn, abi.PropDictionary


class OracleProps(C.Structure):
    _fields_ = [("n_tiles", C.c_uint32), ("n_columns", C.c_uint32), ("n_dictionaries", C.c_uint32), ("reserved", C.c_uint32),
                ("tile_status", C.POINTER(C.c_uint32)), ("columns", C.POINTER(PropColumn)), ("dictionaries", C.POINTER(PropDictionary)),
                ("validity", C.POINTER(C.c_uint8)), ("validity_bytes", C.c_uint64), ("i64", C.POINTER(C.c_int64)), ("n_i64", C.c_uint64),
                ("f32", C.POINTER(C.c_float)), ("n_f32", C.c_uint64), ("f64", C.POINTER(C.c_double)), ("n_f64", C.c_uint64),
                ("bools", C.POINTER(C.c_uint8)), ("bool_bytes", C.c_uint64), ("dict_index", C.POINTER(C.c_int32)),
                ("n_dict_index", C.c_uint64), ("dict_offsets", C.POINTER(C.c_int32)), ("n_dict_offsets", C.c_uint64),
                ("payload_bytes", C.c_uint64)]


PV_NONE, PV_I64, PV_F32, PV_F64, PV_BOOL, PV_DICT_INDEX = range(6)


class PropsResult:
    """Host copy of a covt_oracle_props: `columns` / `dictionaries` structured arrays + the value buffers (indexed by abi.PBUF_*)."""

    def __init__(self, ptr):
        r = ptr.contents

        def arr(p, n, dt):
            return _copy_from(p, n * np.dtype(dt).itemsize, dt) if n else np.zeros(0, dtype=dt)
        self.tile_status = arr(r.tile_status, r.n_tiles, np.uint32)
        self.payload_bytes = int(r.payload_bytes)
        self.columns = (np.frombuffer(C.string_at(r.columns, r.n_columns * C.sizeof(PropColumn)), dtype=abi.PROP_COLUMN_DTYPE).copy()
                        if r.n_columns else np.zeros(0, dtype=abi.PROP_COLUMN_DTYPE))
        self.dictionaries = (np.frombuffer(C.string_at(r.dictionaries, r.n_dictionaries * C.sizeof(PropDictionary)), dtype=abi.PROP_DICTIONARY_DTYPE).copy()
                             if r.n_dictionaries else np.zeros(0, dtype=abi.PROP_DICTIONARY_DTYPE))
        self.validity = arr(r.validity, r.validity_bytes, np.uint8)
        self.i64 = arr(r.i64, r.n_i64, np.int64)
        self.f32 = arr(r.f32, r.n_f32, np.float32)
        self.f64 = arr(r.f64, r.n_f64, np.float64)
        self.bools = arr(r.bools, r.bool_bytes, np.uint8)
        self.dict_index = arr(r.dict_index, r.n_dict_index, np.int32)
        self.dict_offsets = arr(r.dict_offsets, r.n_dict_offsets, np.int32)
        self.buffers = [self.validity, self.i64, self.f32, self.f64, self.bools, self.dict_index, self.dict_offsets]
        lib().covt_oracle_props_free(ptr)

    def values_buffer(self, c):
        return {PV_I64: self.i64, PV_F32: self.f32, PV_F64: self.f64, PV_BOOL: self.bools, PV_DICT_INDEX: self.dict_index}.get(int(c["value_kind"]), self.i64)

    def column_values(self, blob, c):
        """List<Optional> view of one column (CovtParser.decodePropertyColumn): value or None per feature."""
        return abi.prop_column_values(blob, c, self.validity, self.values_buffer(c), self.dict_offsets, self.dictionaries)


def decode_properties(blob, tile_offsets, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None):
    b = _as_u8(blob)
    offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
    out = C.POINTER(OracleProps)()
    L = lib()
    L.covt_oracle_decode_properties.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(OracleProps))]
    L.covt_oracle_decode_properties.restype = C.c_int32
    L.covt_oracle_props_free.argtypes = [C.POINTER(OracleProps)]
    L.covt_oracle_props_free.restype = None
    tj, keep = make_tilejson(n_fields)
    rc = L.covt_oracle_decode_properties(b.ctypes.data, offs.ctypes.data, len(offs) - 1, container,
                                         C.cast(C.byref(tj), C.c_void_p) if tj is not None else None, flags, C.byref(out))
    if rc != 0:
        raise RuntimeError("covt_oracle_decode_properties failed: %d" % rc)
    return PropsResult(out)
