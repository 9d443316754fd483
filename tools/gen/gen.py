"""ctypes binding of tools/gen/libcovt_gen.so: COVT encoders + synthetic tile generator (input synthesis)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ALLOW_PFOR_TOPOLOGY = 0x01
ALLOW_PFOR_VERTEX = 0x02
ICE_MORTON = 0x04
ID_DELTA_VARINT = 0x08
FORCE_VARINT_VERTEX = 0x10
FORCE_RLE_TOPOLOGY = 0x20
OPTIMIZED_METADATA = 0x40


def build(force=False):
    so = os.path.join(_HERE, "libcovt_gen.so")
    src = [os.path.join(_HERE, f) for f in ("covt_gen.c", "covt_gen.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libcovt_gen.so"], stdout=subprocess.DEVNULL)
    return so


class GenBuf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t)]


class GenLayer(C.Structure):
    _fields_ = [("name", C.c_char_p), ("extent", C.c_uint32), ("n_features", C.c_uint32),
                ("types", C.c_void_p),
                ("geom_counts", C.c_void_p), ("n_geom", C.c_uint32),
                ("part_counts", C.c_void_p), ("n_part", C.c_uint32),
                ("ring_counts", C.c_void_p), ("n_ring", C.c_uint32),
                ("xy", C.c_void_p), ("n_vertices", C.c_uint32),
                ("ids", C.c_void_p),
                ("index_buffer", C.c_void_p), ("n_index", C.c_uint32)]


class GenParams(C.Structure):
    _fields_ = [("layers_per_tile", C.c_uint32), ("mean_features", C.c_double),
                ("p_point", C.c_double), ("p_line", C.c_double), ("p_polygon", C.c_double),
                ("p_multiline", C.c_double), ("p_multipolygon", C.c_double),
                ("mean_line_extra", C.c_double), ("mean_ring_extra", C.c_double), ("p_second_ring", C.c_double),
                ("extent", C.c_uint32), ("container", C.c_uint32), ("with_ids", C.c_uint32),
                ("with_index_buffer", C.c_uint32), ("max_step", C.c_uint32)]


class GenTruth(C.Structure):
    _fields_ = [("features", C.c_uint64), ("vertices", C.c_uint64), ("parts", C.c_uint64), ("rings", C.c_uint64),
                ("polygon_rings", C.c_uint64), ("sum_x", C.c_int64), ("sum_y", C.c_int64),
                ("sum_x_closed", C.c_int64), ("sum_y_closed", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.covt_enc_varints.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _lib.covt_enc_varints.restype = C.c_size_t
        _lib.covt_enc_rle.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        _lib.covt_enc_rle.restype = C.c_size_t
        _lib.covt_enc_byte_rle.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _lib.covt_enc_byte_rle.restype = C.c_size_t
        _lib.covt_enc_fastpfor.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _lib.covt_enc_fastpfor.restype = C.c_size_t
        _lib.covt_enc_zigzag_delta_coordinates.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib.covt_enc_zigzag_delta_coordinates.restype = None
        _lib.covt_enc_morton.argtypes = [C.c_int32, C.c_int32, C.c_uint32]
        _lib.covt_enc_morton.restype = C.c_int32
        _lib.covt_gen_varint_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib.covt_gen_varint_stream.restype = C.c_uint64
        _lib.covt_gen_default_params.argtypes = [C.POINTER(GenParams)]
        _lib.covt_gen_default_params.restype = None
        _lib.covt_gen_tiles.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(GenParams), C.c_uint32,
                                        C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64), C.c_void_p,
                                        C.POINTER(GenTruth)]
        _lib.covt_gen_tiles.restype = C.c_int32
        _lib.covt_gen_free.argtypes = [C.c_void_p]
        _lib.covt_gen_free.restype = None
        _lib.covt_gen_begin_tile.argtypes = [C.POINTER(GenBuf), C.c_uint32, C.c_uint32]
        _lib.covt_gen_append_layer.argtypes = [C.POINTER(GenBuf), C.POINTER(GenLayer), C.c_uint32, C.c_uint32]
        _lib.covt_gen_append_layer.restype = C.c_int32
        _lib.covt_gen_buf_free.argtypes = [C.POINTER(GenBuf)]
    return _lib


def _enc(fn, arr, cap, *args):
    out = np.empty(cap, dtype=np.uint8)
    n = fn(arr.ctypes.data, len(arr), *args, out.ctypes.data, cap)
    if n == C.c_size_t(-1).value:
        raise RuntimeError("encoder buffer too small")
    return out[:n].copy()


def encode_varints(values, zigzag=False, delta=False):
    """EncodingUtils.encodeVarints (EncodingUtils.java:39-55); values as int64."""
    a = np.ascontiguousarray(values, dtype=np.int64)
    return _enc(lib().covt_enc_varints, a, len(a) * 10 + 16, int(zigzag), int(delta))


def encode_rle(values, signed=False):
    """EncodingUtils.encodeRle (EncodingUtils.java:123-134)."""
    a = np.ascontiguousarray(values, dtype=np.int64)
    return _enc(lib().covt_enc_rle, a, len(a) * 11 + 16, int(signed))


def encode_byte_rle(values):
    """EncodingUtils.encodeByteRle (EncodingUtils.java:136-147)."""
    a = np.ascontiguousarray(values, dtype=np.uint8)
    return _enc(lib().covt_enc_byte_rle, a, len(a) * 2 + 16)


def encode_fastpfor(values, zigzag=False, delta=False):
    """EncodingUtils.encodeFastPfor128 (EncodingUtils.java:149-188); values as int32."""
    a = np.ascontiguousarray(values, dtype=np.int32)
    return _enc(lib().covt_enc_fastpfor, a, len(a) * 5 + 8192, int(zigzag), int(delta))


def encode_zigzag_delta_coordinates(xy):
    a = np.ascontiguousarray(xy, dtype=np.int32)
    out = np.empty_like(a)
    lib().covt_enc_zigzag_delta_coordinates(a.ctypes.data, len(a), out.ctypes.data)
    return out


def encode_morton(x, y, num_bits):
    return lib().covt_enc_morton(int(x), int(y), int(num_bits))


def varint_stream(target_bytes, seed=0xC0717):
    """Config 3 stream. Returns (bytes ndarray, n_values)."""
    out = np.empty(target_bytes + 16, dtype=np.uint8)
    n = lib().covt_gen_varint_stream(out.ctypes.data, target_bytes, seed)
    return out[:target_bytes], int(n)


def default_params(**kw):
    p = GenParams()
    lib().covt_gen_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def tiles(first_tile, n_tiles, params=None, n_threads=0):
    """Config 5 tiles. Returns (blob uint8 ndarray, tile_offsets uint64 ndarray, truth dict)."""
    if params is None:
        params = default_params()
    offs = np.zeros(n_tiles + 1, dtype=np.uint64)
    blob = C.POINTER(C.c_uint8)()
    blen = C.c_uint64()
    truth = GenTruth()
    rc = lib().covt_gen_tiles(first_tile, n_tiles, C.byref(params), n_threads, C.byref(blob), C.byref(blen),
                              offs.ctypes.data, C.byref(truth))
    if rc != 0:
        raise MemoryError("covt_gen_tiles failed")
    arr = np.ctypeslib.as_array(blob, shape=(max(1, blen.value),))[: blen.value].copy()
    lib().covt_gen_free(blob)
    return arr, offs, truth.as_dict()


def make_tile(layers, container=0, options=ALLOW_PFOR_TOPOLOGY | ALLOW_PFOR_VERTEX):
    """Encode one tile from explicit layers. Each layer is a dict with keys name, extent, types, geom, part, ring,
    xy (flat x,y list, no closing vertices), optional ids, index_buffer, options."""
    buf = GenBuf()
    L = lib()
    L.covt_gen_begin_tile(C.byref(buf), container, len(layers))
    keep = []
    for ly in layers:
        types = np.ascontiguousarray(ly["types"], dtype=np.uint8)
        geom = np.ascontiguousarray(ly.get("geom", []), dtype=np.int32)
        part = np.ascontiguousarray(ly.get("part", []), dtype=np.int32)
        ring = np.ascontiguousarray(ly.get("ring", []), dtype=np.int32)
        xy = np.ascontiguousarray(ly["xy"], dtype=np.int32)
        ids = np.ascontiguousarray(ly["ids"], dtype=np.int64) if ly.get("ids") is not None else None
        idx = np.ascontiguousarray(ly["index_buffer"], dtype=np.int32) if ly.get("index_buffer") is not None else None
        keep += [types, geom, part, ring, xy, ids, idx]
        g = GenLayer(name=ly.get("name", "layer").encode(), extent=ly.get("extent", 4096), n_features=len(types),
                     types=types.ctypes.data, geom_counts=geom.ctypes.data, n_geom=len(geom),
                     part_counts=part.ctypes.data, n_part=len(part), ring_counts=ring.ctypes.data, n_ring=len(ring),
                     xy=xy.ctypes.data, n_vertices=len(xy) // 2, ids=ids.ctypes.data if ids is not None else None,
                     index_buffer=idx.ctypes.data if idx is not None else None, n_index=len(idx) if idx is not None else 0)
        rc = L.covt_gen_append_layer(C.byref(buf), C.byref(g), container, ly.get("options", options))
        if rc != 0:
            raise RuntimeError("covt_gen_append_layer failed")
    out = np.ctypeslib.as_array(buf.data, shape=(max(1, buf.len),))[: buf.len].copy()
    L.covt_gen_buf_free(C.byref(buf))
    return out
