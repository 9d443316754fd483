"""cov-tiles_b200 — B200-native COVT tile-batch decoder (host-side binding of libcovt_b200.so).

This package is plumbing over the C ABI in include/covt_b200.h: it loads the in-tree shared library
(hand-written sm_100a kernels) with ctypes and mirrors the reference decoder's interface for the
tile-decode path:

    CovtParser.decode_covt(...)      <- CovtParser.decodeCovt(byte[], TileJson)   J/decoder/CovtParser.java:53
    DecodingUtils.decode_*(...)      <- the static codecs of DecodingUtils        J/decoder/DecodingUtils.java:35-444

There is no CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi  # noqa: F401
from .abi import *  # noqa: F401,F403

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("COVT_LIB") or os.path.join(_HERE, "libcovt_b200.so")  # COVT_LIB: an experiment build of the same library
_lib = None


class CovtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("covt error %d (%s): %s" % (code, abi.STATUS_NAMES[code] if 0 <= code < len(abi.STATUS_NAMES) else "?", message))
        self.code = code


def build(force=False, verbose=False):
    """Compile libcovt_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "covt_b200.h"))
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", csrc, "-j2"], stdout=None if verbose else subprocess.DEVNULL)
    return _LIB_PATH


# every symbol include/covt_b200.h declares
ABI_SYMBOLS = [
    "covt_abi_version", "covt_create", "covt_destroy", "covt_last_error", "covt_trim", "covt_decode_batch", "covt_decode_batch_to_host",
    "covt_batch_upload",
    "covt_batch_decode", "covt_batch_free", "covt_decode_streams", "covt_batch_decode_streams", "covt_encode_streams", "covt_resolve_op",
    "covt_result_num_tiles", "covt_result_num_layers", "covt_result_layers", "covt_result_tile_status",
    "covt_result_buffer", "covt_result_read", "covt_result_timing", "covt_result_kernel_times", "covt_result_free",
    "covt_result_prop_columns", "covt_result_prop_dictionaries", "covt_result_prop_buffer", "covt_result_prop_read",
    "covt_host_register", "covt_host_unregister", "covt_partition_tiles",
    "covt_create_multi", "covt_destroy_multi", "covt_multi_last_error", "covt_multi_device_count", "covt_multi_context",
    "covt_decode_batch_multi", "covt_multi_result_parts", "covt_multi_result_part", "covt_multi_result_timing",
    "covt_multi_result_read", "covt_multi_result_free",
]


def lib():
    """Loads libcovt_b200.so (must have been built: see build() / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise CovtError(abi.ERR_CUDA, "libcovt_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(_LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    L.covt_abi_version.restype = i32
    L.covt_create.argtypes = [i32, C.POINTER(vp)]
    L.covt_destroy.argtypes = [vp]
    L.covt_destroy.restype = None
    L.covt_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.covt_trim.argtypes = [vp]
    L.covt_decode_batch.argtypes = [vp, vp, vp, u32, u32, C.POINTER(abi.TileJson), u32, C.POINTER(vp)]
    L.covt_decode_batch_to_host.argtypes = [vp, vp, vp, u32, u32, C.POINTER(abi.TileJson), u32, C.POINTER(abi.HostSink), C.POINTER(vp)]
    L.covt_batch_upload.argtypes = [vp, vp, vp, u32, C.POINTER(vp)]
    L.covt_batch_decode.argtypes = [vp, vp, u32, C.POINTER(abi.TileJson), u32, C.POINTER(vp)]
    L.covt_batch_free.argtypes = [vp]
    L.covt_batch_free.restype = None
    L.covt_decode_streams.argtypes = [vp, vp, u64, C.POINTER(abi.StreamDesc), u32, u32, C.POINTER(vp)]
    L.covt_batch_decode_streams.argtypes = [vp, vp, C.POINTER(abi.StreamDesc), u32, u32, C.POINTER(vp)]
    L.covt_encode_streams.argtypes = [vp, vp, u64, C.POINTER(abi.EncodeDesc), u32, u32, C.POINTER(vp)]
    L.covt_resolve_op.argtypes = [u32, u32, u32, u32]
    L.covt_result_num_tiles.argtypes = [vp]
    L.covt_result_num_tiles.restype = u32
    L.covt_result_num_layers.argtypes = [vp]
    L.covt_result_num_layers.restype = u32
    L.covt_result_layers.argtypes = [vp, C.POINTER(C.POINTER(abi.Layer))]
    L.covt_result_tile_status.argtypes = [vp, C.POINTER(C.POINTER(u32)), C.POINTER(C.POINTER(u32))]
    L.covt_result_buffer.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64), C.POINTER(u32)]
    L.covt_result_read.argtypes = [vp, u32, u64, u64, vp]
    L.covt_result_prop_columns.argtypes = [vp, C.POINTER(C.POINTER(abi.PropColumn)), C.POINTER(u32)]
    L.covt_result_prop_dictionaries.argtypes = [vp, C.POINTER(C.POINTER(abi.PropDictionary)), C.POINTER(u32)]
    L.covt_result_prop_buffer.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64), C.POINTER(u32)]
    L.covt_result_prop_read.argtypes = [vp, u32, u64, u64, vp]
    L.covt_result_timing.argtypes = [vp, C.POINTER(abi.Timing)]
    L.covt_result_kernel_times.argtypes = [vp, C.POINTER(abi.KernelTime), u32, C.POINTER(u32)]
    L.covt_result_free.argtypes = [vp]
    L.covt_result_free.restype = None
    L.covt_host_register.argtypes = [vp, vp, C.c_size_t]
    L.covt_host_unregister.argtypes = [vp, vp]
    L.covt_partition_tiles.argtypes = [vp, u32, u32, vp]
    L.covt_create_multi.argtypes = [u32, C.POINTER(i32), C.POINTER(vp)]
    L.covt_destroy_multi.argtypes = [vp]
    L.covt_destroy_multi.restype = None
    L.covt_multi_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.covt_multi_device_count.argtypes = [vp]
    L.covt_multi_device_count.restype = u32
    L.covt_multi_context.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(i32)]
    L.covt_decode_batch_multi.argtypes = [vp, vp, vp, u32, u32, C.POINTER(abi.TileJson), u32, C.POINTER(vp)]
    L.covt_multi_result_parts.argtypes = [vp]
    L.covt_multi_result_parts.restype = u32
    L.covt_multi_result_part.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(u32), C.POINTER(i32)]
    L.covt_multi_result_timing.argtypes = [vp, C.POINTER(abi.Timing)]
    L.covt_multi_result_read.argtypes = [vp, u32, C.POINTER(vp)]
    L.covt_multi_result_free.argtypes = [vp]
    L.covt_multi_result_free.restype = None
    for name in ABI_SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:
            fn.restype = i32
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def partition_tiles(tile_offsets, n_parts):
    """Batch scheduler: contiguous tile ranges balanced by payload bytes (no collective; tiles share nothing)."""
    offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
    starts = np.zeros(n_parts + 1, dtype=np.uint32)
    rc = lib().covt_partition_tiles(offs.ctypes.data, len(offs) - 1, n_parts, starts.ctypes.data)
    if rc:
        raise CovtError(rc, "covt_partition_tiles")
    return starts


class Result:
    """Owns a covt_result (device-resident buffers + layer table)."""

    def __init__(self, dec, handle):
        self._dec = dec
        self._h = handle
        self._layers = None

    def free(self):
        # (a result outlives its context only by accident — e.g. an exception between decode and free followed by Decoder.close();
        # the context owned the device blocks, so there is nothing left to give back and the handle must not be touched)
        if self._h and getattr(self._dec, "_h", None):
            lib().covt_result_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def n_tiles(self):
        return lib().covt_result_num_tiles(self._h)

    @property
    def n_layers(self):
        return lib().covt_result_num_layers(self._h)

    @property
    def layers(self):
        """numpy structured array of covt_layer (device->host on first use)."""
        if self._layers is None:
            p = C.POINTER(abi.Layer)()
            self._dec._check(lib().covt_result_layers(self._h, C.byref(p)))
            n = self.n_layers
            if n:
                self._layers = np.frombuffer(C.string_at(p, n * C.sizeof(abi.Layer)), dtype=abi.LAYER_DTYPE).copy()
            else:
                self._layers = np.zeros(0, dtype=abi.LAYER_DTYPE)
        return self._layers

    def tile_status(self):
        s = C.POINTER(C.c_uint32)()
        f = C.POINTER(C.c_uint32)()
        self._dec._check(lib().covt_result_tile_status(self._h, C.byref(s), C.byref(f)))
        n = self.n_tiles
        status = np.ctypeslib.as_array(s, shape=(max(n, 1),))[:n].copy()
        first = np.ctypeslib.as_array(f, shape=(n + 1,)).copy()
        return status, first

    def touch_tile_status(self):
        """Device->host copy of the per-tile status + layer index without materialising numpy copies."""
        s = C.POINTER(C.c_uint32)()
        f = C.POINTER(C.c_uint32)()
        self._dec._check(lib().covt_result_tile_status(self._h, C.byref(s), C.byref(f)))

    def device_buffer(self, which):
        """(device pointer, element count, element size) of one result buffer."""
        p, n, es = C.c_void_p(), C.c_uint64(), C.c_uint32()
        self._dec._check(lib().covt_result_buffer(self._h, which, C.byref(p), C.byref(n), C.byref(es)))
        return p.value or 0, n.value, es.value

    def buffer(self, which, offset=0, count=None):
        """Host copy (numpy) of a result buffer or a slice of it."""
        _, n, _ = self.device_buffer(which)
        if count is None:
            count = n - offset
        out = np.empty(count, dtype=abi.BUF_DTYPES[which])
        if count:
            self._dec._check(lib().covt_result_read(self._h, which, offset, count, out.ctypes.data))
        return out

    def read_into(self, which, offset, count, host_ptr):
        self._dec._check(lib().covt_result_read(self._h, which, offset, count, host_ptr))

    def to_arrow(self, blob, with_props=True):
        """[(tile index, layer name, pyarrow.Table)]: id, geometry_type, the GeoArrow-nested geometry and (with FLAG_DECODE_PROPERTIES)
        the property columns of every decoded layer, wrapped around host copies of the result buffers (see arrow.py)."""
        from . import arrow

        class _Lazy(dict):
            def __missing__(s, which):
                s[which] = self.buffer(which)
                return s[which]

        props = None
        if with_props and len(self.prop_columns()):
            class _P:
                pass
            props = _P()
            props.columns, props.dictionaries = self.prop_columns(), self.prop_dictionaries()
            props.buffers = [self.prop_buffer(b) for b in range(abi.NUM_PROP_BUFFERS)]
            props.validity, props.dict_offsets = props.buffers[abi.PBUF_VALIDITY], props.buffers[abi.PBUF_DICT_OFFSETS]
        return arrow.layer_tables(np.ascontiguousarray(blob, dtype=np.uint8), self.layers, _Lazy(), props)

    # ---- property columns (FLAG_DECODE_PROPERTIES) ----
    def prop_columns(self):
        """numpy structured array of covt_prop_column (one record per decoded property column / localized sub-column)."""
        p, n = C.POINTER(abi.PropColumn)(), C.c_uint32()
        self._dec._check(lib().covt_result_prop_columns(self._h, C.byref(p), C.byref(n)))
        if not n.value:
            return np.zeros(0, dtype=abi.PROP_COLUMN_DTYPE)
        return np.frombuffer(C.string_at(p, n.value * C.sizeof(abi.PropColumn)), dtype=abi.PROP_COLUMN_DTYPE).copy()

    def prop_dictionaries(self):
        p, n = C.POINTER(abi.PropDictionary)(), C.c_uint32()
        self._dec._check(lib().covt_result_prop_dictionaries(self._h, C.byref(p), C.byref(n)))
        if not n.value:
            return np.zeros(0, dtype=abi.PROP_DICTIONARY_DTYPE)
        return np.frombuffer(C.string_at(p, n.value * C.sizeof(abi.PropDictionary)), dtype=abi.PROP_DICTIONARY_DTYPE).copy()

    def prop_device_buffer(self, which):
        p, n, es = C.c_void_p(), C.c_uint64(), C.c_uint32()
        self._dec._check(lib().covt_result_prop_buffer(self._h, which, C.byref(p), C.byref(n), C.byref(es)))
        return p.value or 0, n.value, es.value

    def prop_buffer(self, which, offset=0, count=None):
        """Host copy (numpy) of a property value buffer (abi.PBUF_*) or a slice of it."""
        _, n, _ = self.prop_device_buffer(which)
        if count is None:
            count = n - offset
        out = np.empty(count, dtype=abi.PBUF_DTYPES[which])
        if count:
            self._dec._check(lib().covt_result_prop_read(self._h, which, offset, count, out.ctypes.data))
        return out

    def timing(self):
        t = abi.Timing()
        self._dec._check(lib().covt_result_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in abi.Timing._fields_}

    def kernel_times(self):
        arr = (abi.KernelTime * 32)()
        n = C.c_uint32()
        self._dec._check(lib().covt_result_kernel_times(self._h, arr, 32, C.byref(n)))
        return [{"name": arr[i].name.decode(), "ms": arr[i].ms, "launches": arr[i].launches,
                 "algorithmic_bytes": arr[i].algorithmic_bytes} for i in range(min(n.value, 32))]


class Batch:
    """A blob of tiles resident in HBM (covt_batch)."""

    def __init__(self, dec, handle, n_tiles, nbytes):
        self._dec, self._h, self.n_tiles, self.nbytes = dec, handle, n_tiles, nbytes

    def free(self):
        if self._h:
            lib().covt_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Decoder:
    """One decoder context per GPU (covt_ctx). Raises CovtError if no B200 / library is available."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().covt_create(device, C.byref(self._h))
        if rc:
            buf = C.create_string_buffer(512)
            lib().covt_last_error(None, buf, 512)
            self._h = None
            raise CovtError(rc, buf.value.decode())
        self.device = device

    def close(self):
        if self._h:
            lib().covt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """Return the device blocks parked by finished decodes to the driver (covt_trim)."""
        self._check(lib().covt_trim(self._h))

    def _check(self, rc):
        if rc:
            buf = C.create_string_buffer(512)
            lib().covt_last_error(self._h, buf, 512)
            raise CovtError(rc, buf.value.decode())

    @staticmethod
    def _tilejson(n_fields):
        if n_fields is None:
            return None, None
        arr = (C.c_uint32 * max(1, len(n_fields)))(*n_fields)
        return abi.TileJson(len(n_fields), C.cast(arr, C.POINTER(C.c_uint32))), arr

    # ---- batch path -------------------------------------------------------------------------
    def upload(self, blob, tile_offsets):
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
        h = C.c_void_p()
        self._check(lib().covt_batch_upload(self._h, _ptr(b), offs.ctypes.data, len(offs) - 1, C.byref(h)))
        return Batch(self, h, len(offs) - 1, int(offs[-1]))

    def upload_raw(self, blob_ptr, tile_offsets_ptr, n_tiles, nbytes):
        """Upload from raw host pointers (pinned memory owned by the caller)."""
        h = C.c_void_p()
        self._check(lib().covt_batch_upload(self._h, blob_ptr, tile_offsets_ptr, n_tiles, C.byref(h)))
        return Batch(self, h, n_tiles, nbytes)

    def decode(self, batch, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None):
        tj, keep = self._tilejson(n_fields)
        h = C.c_void_p()
        self._check(lib().covt_batch_decode(self._h, batch._h, container, C.byref(tj) if tj is not None else None, flags, C.byref(h)))
        return Result(self, h)

    def decode_batch(self, blob, tile_offsets, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None):
        """The reference-facing call with HOST buffers: upload + decode (covt_decode_batch)."""
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
        tj, keep = self._tilejson(n_fields)
        h = C.c_void_p()
        self._check(lib().covt_decode_batch(self._h, _ptr(b), offs.ctypes.data, len(offs) - 1, container,
                                            C.byref(tj) if tj is not None else None, flags, C.byref(h)))
        return Result(self, h)

    def decode_batch_raw(self, blob_ptr, tile_offsets_ptr, n_tiles, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT):
        h = C.c_void_p()
        self._check(lib().covt_decode_batch(self._h, blob_ptr, tile_offsets_ptr, n_tiles, container, None, flags, C.byref(h)))
        return Result(self, h)

    def decode_batch_to_host_raw(self, blob_ptr, tile_offsets_ptr, n_tiles, sink, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT):
        """covt_decode_batch_to_host: like decode_batch_raw, and every buffer the abi.HostSink names lands in its (page-locked) host
        memory, copied back segment by segment while later segments are uploaded and decoded."""
        h = C.c_void_p()
        self._check(lib().covt_decode_batch_to_host(self._h, blob_ptr, tile_offsets_ptr, n_tiles, container, None, flags, C.byref(sink), C.byref(h)))
        return Result(self, h)

    def decode_batch_to_host(self, blob, tile_offsets, host_arrays, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT):
        """host_arrays: {buffer id: numpy array of that buffer's dtype} (page-locked memory for overlap; any memory works)."""
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
        sink = abi.HostSink()
        for which, a in host_arrays.items():
            assert a.dtype == np.dtype(abi.BUF_DTYPES[which]) and a.flags["C_CONTIGUOUS"]
            sink.ptr[which] = a.ctypes.data
            sink.capacity[which] = a.size
        return self.decode_batch_to_host_raw(_ptr(b), offs.ctypes.data, len(offs) - 1, sink, container, flags)

    # ---- stream path ------------------------------------------------------------------------
    def decode_streams(self, blob_or_batch, descs, flags=abi.FLAG_DEFAULT):
        """descs: ctypes array of abi.StreamDesc (filled in place). Returns a Result holding BUF_STREAM_ARENA."""
        h = C.c_void_p()
        n = len(descs)
        if isinstance(blob_or_batch, Batch):
            self._check(lib().covt_batch_decode_streams(self._h, blob_or_batch._h, descs, n, flags, C.byref(h)))
        else:
            b = np.ascontiguousarray(blob_or_batch, dtype=np.uint8)
            self._check(lib().covt_decode_streams(self._h, _ptr(b), b.size, descs, n, flags, C.byref(h)))
        return Result(self, h)

    def encode_streams(self, values, descs, flags=abi.FLAG_DEFAULT):
        """values: one host buffer (any dtype, viewed as bytes); descs: ctypes array of abi.EncodeDesc (filled in place).
        Returns a Result whose BUF_STREAM_ARENA holds the encoded streams."""
        h = C.c_void_p()
        v = np.ascontiguousarray(values).view(np.uint8).reshape(-1)
        self._check(lib().covt_encode_streams(self._h, _ptr(v), v.size, descs, len(descs), flags, C.byref(h)))
        return Result(self, h)

    def encode_stream(self, values, op, num_bits=0, flags=abi.FLAG_DEFAULT):
        """One EncodingUtils call: the values a stream of `op` decodes to -> (its bytes, status). Morton ops take an (n, 2) int32 array."""
        v = np.ascontiguousarray(values, dtype=np.int32 if op in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON) else abi.op_dtype(op))
        n = len(v) if op in (abi.OP_VARINT_DELTA_MORTON, abi.OP_PFOR_DELTA_MORTON) and v.ndim == 2 else v.size
        descs = (abi.EncodeDesc * 1)()
        descs[0] = abi.EncodeDesc(value_offset=0, num_values=n, op=op, num_bits=num_bits)
        res = self.encode_streams(v, descs, flags)
        d = descs[0]
        raw = res.buffer(abi.BUF_STREAM_ARENA, d.out_offset, d.byte_length) if d.byte_length else np.zeros(0, np.uint8)
        res.free()
        return raw.copy(), d.status

    def decode_stream(self, blob, op=0, *, byte_offset=0, byte_length=None, num_values, stream_type=0, encoding=0,
                      column_type=0, num_bits=0, flags=abi.FLAG_DEFAULT):
        """One DecodingUtils call. Returns (values ndarray, status, bytes_consumed)."""
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        if byte_length is None:
            byte_length = b.size - byte_offset
        descs = (abi.StreamDesc * 1)()
        descs[0] = abi.StreamDesc(byte_offset=byte_offset, byte_length=byte_length, num_values=num_values,
                                  stream_type=stream_type, encoding=encoding, column_type=column_type,
                                  num_bits=num_bits, op=op)
        res = self.decode_streams(b, descs, flags)
        d = descs[0]
        rop = op or lib().covt_resolve_op(stream_type, encoding, column_type, flags)
        dt = np.dtype(abi.op_dtype(rop))
        raw = res.buffer(abi.BUF_STREAM_ARENA, d.out_offset, d.out_count * dt.itemsize) if d.out_count else np.zeros(0, np.uint8)
        res.free()
        return raw.view(dt).copy(), d.status, d.bytes_consumed


class MultiResult:
    """Owns a covt_multi_result: one Result per GPU (device-resident), tile ranges in batch order."""

    def __init__(self, owner, handle):
        self._owner, self._h = owner, handle
        self.parts = []
        for p in range(lib().covt_multi_result_parts(handle)):
            r, t0, n, dev = C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_int32()
            owner._check(lib().covt_multi_result_part(handle, p, C.byref(r), C.byref(t0), C.byref(n), C.byref(dev)))
            res = Result(owner._part_checker(p), r)
            res.free = lambda: None  # owned by the multi result
            self.parts.append({"result": res, "first_tile": t0.value, "n_tiles": n.value, "device": dev.value})

    def timing(self):
        t = abi.Timing()
        self._owner._check(lib().covt_multi_result_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in abi.Timing._fields_}

    def read_into(self, which, host_ptrs):
        """Device->host copy of buffer `which` of every part at once (all GPUs copy side by side): part p -> host_ptrs[p]."""
        arr = (C.c_void_p * len(host_ptrs))(*host_ptrs)
        self._owner._check(lib().covt_multi_result_read(self._h, which, arr))

    def free(self):
        if self._h:
            for p in self.parts:
                p["result"]._h = None
            lib().covt_multi_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _PartChecker:
    """Error lookup of one per-GPU context of a MultiDecoder (Result only needs _check)."""

    def __init__(self, ctx):
        self._h = ctx

    def _check(self, rc):
        if rc:
            buf = C.create_string_buffer(512)
            lib().covt_last_error(self._h, buf, 512)
            raise CovtError(rc, buf.value.decode())


class MultiDecoder:
    """The batch scheduler of the library (covt_create_multi): ONE call decodes a host batch on several GPUs of this box, one
    persistent host thread + context per GPU, contiguous tile ranges balanced by payload bytes, no collective."""

    def __init__(self, device_ids=None):
        self._h = C.c_void_p()
        n = 0 if device_ids is None else len(device_ids)
        ids = (C.c_int32 * max(n, 1))(*(device_ids or [0]))
        rc = lib().covt_create_multi(n, ids if device_ids is not None else None, C.byref(self._h))
        if rc:
            buf = C.create_string_buffer(512)
            lib().covt_multi_last_error(None, buf, 512)
            self._h = None
            raise CovtError(rc, buf.value.decode())
        self.n_devices = lib().covt_multi_device_count(self._h)

    def _check(self, rc):
        if rc:
            buf = C.create_string_buffer(768)
            lib().covt_multi_last_error(self._h, buf, 768)
            raise CovtError(rc, buf.value.decode())

    def _part_checker(self, part):
        ctx, dev = C.c_void_p(), C.c_int32()
        self._check(lib().covt_multi_context(self._h, part, C.byref(ctx), C.byref(dev)))
        return _PartChecker(ctx)

    def decode_batch(self, blob, tile_offsets, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT, n_fields=None):
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        offs = np.ascontiguousarray(tile_offsets, dtype=np.uint64)
        tj, keep = Decoder._tilejson(n_fields)
        h = C.c_void_p()
        self._check(lib().covt_decode_batch_multi(self._h, _ptr(b), offs.ctypes.data, len(offs) - 1, container,
                                                  C.byref(tj) if tj is not None else None, flags, C.byref(h)))
        return MultiResult(self, h)

    def decode_batch_raw(self, blob_ptr, tile_offsets_ptr, n_tiles, container=abi.CONTAINER_GEN2B, flags=abi.FLAG_DEFAULT):
        h = C.c_void_p()
        self._check(lib().covt_decode_batch_multi(self._h, blob_ptr, tile_offsets_ptr, n_tiles, container, None, flags, C.byref(h)))
        return MultiResult(self, h)

    def close(self):
        if self._h:
            lib().covt_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- mirrors of the reference's static interface -----------------------------------------------------
_default = {}


def default_decoder(device=0):
    if device not in _default:
        _default[device] = Decoder(device)
    return _default[device]


class EncodingUtils:
    """Same names and argument meaning as J/converter/EncodingUtils.java (the static encoders), on the GPU."""

    @staticmethod
    def _run(op, values, num_bits=0):
        raw, st = default_decoder().encode_stream(values, op, num_bits=num_bits)
        if st != abi.OK:
            raise CovtError(st, "stream encode failed")
        return raw

    @staticmethod
    def encodeVarints(values, zigZag, delta):  # EncodingUtils.java:39 (long[] in; 64-bit LEB128)
        v = np.asarray(values, dtype=np.int64)
        if zigZag and delta:
            return EncodingUtils._run(abi.OP_VARINT_ZZ_DELTA_64, v)
        if not zigZag and not delta:
            return EncodingUtils._run(abi.OP_VARINT_U64, v)
        if zigZag:  # zigzag without delta: 32-bit values only (INT_64 property data, CovtParser.java:303-306)
            return EncodingUtils._run(abi.OP_VARINT_ZZ, v.astype(np.int32))
        raise CovtError(abi.ERR_UNSUPPORTED_ENCODING, "delta without zigzag is only written for Morton codes (encodeMortonDeltaVarints)")

    @staticmethod
    def encodeZigZagDeltaCoordinates(vertices):  # :190-211 + encodeVarints(.., false, false): the PLAIN vertex buffer
        return EncodingUtils._run(abi.OP_VARINT_ZZ_DELTA_XY, np.asarray(vertices, dtype=np.int32).reshape(-1))

    @staticmethod
    def encodeRle(values, signed):  # :123
        return EncodingUtils._run(abi.OP_RLE_S64 if signed else abi.OP_RLE_U64, np.asarray(values, dtype=np.int64))

    @staticmethod
    def encodeByteRle(values):  # :136
        return EncodingUtils._run(abi.OP_BYTE_RLE, np.asarray(values, dtype=np.uint8))

    @staticmethod
    def encodeFastPfor128(values, zigZag=True, delta=True):  # :149 (topology streams: zigzag + delta)
        if not (zigZag and delta):
            raise CovtError(abi.ERR_UNSUPPORTED_ENCODING, "the decode path reads FastPFOR streams as zigzag deltas (coordinates: "
                            "encodeFastPfor128Coordinates, Morton codes: encodeMortonFastPfor128)")
        return EncodingUtils._run(abi.OP_PFOR_ZZ_DELTA, np.asarray(values, dtype=np.int32))

    @staticmethod
    def encodeFastPfor128Coordinates(vertices):  # encodeZigZagDeltaCoordinates + encodeFastPfor128(.., false, false)
        return EncodingUtils._run(abi.OP_PFOR_ZZ_DELTA_XY, np.asarray(vertices, dtype=np.int32).reshape(-1))

    @staticmethod
    def encodeMortonDeltaVarints(vertices, numBits):  # GeometryUtils.encodeMorton + encodeVarints(codes, false, true)
        return EncodingUtils._run(abi.OP_VARINT_DELTA_MORTON, np.asarray(vertices, dtype=np.int32).reshape(-1, 2), num_bits=numBits)

    @staticmethod
    def encodeMortonFastPfor128(vertices, numBits):  # GeometryUtils.encodeMorton + encodeFastPfor128(codes, false, true)
        return EncodingUtils._run(abi.OP_PFOR_DELTA_MORTON, np.asarray(vertices, dtype=np.int32).reshape(-1, 2), num_bits=numBits)


class DecodingUtils:
    """Same names and argument meaning as J/decoder/DecodingUtils.java; `pos` is a one-element list used like
    me.lemire.integercompression.IntWrapper (advanced by the call)."""

    @staticmethod
    def _run(op, buf, pos, n, byte_length=None, num_bits=0, advance_by_length=False):
        vals, st, consumed = default_decoder().decode_stream(buf, op, byte_offset=pos[0], byte_length=byte_length,
                                                             num_values=n, num_bits=num_bits)
        if st not in (abi.OK,):
            raise CovtError(st, "stream decode failed")
        pos[0] += byte_length if advance_by_length else consumed
        return vals

    @staticmethod
    def decodeVarint(src, pos, numValues):  # DecodingUtils.java:35
        return DecodingUtils._run(abi.OP_VARINT_U32, src, pos, numValues)

    @staticmethod
    def decodeZigZagVarint(buf, pos, numValues):  # :46
        return DecodingUtils._run(abi.OP_VARINT_ZZ, buf, pos, numValues)

    @staticmethod
    def decodeZigZagDeltaVarint(buf, pos, numValues):  # :55
        return DecodingUtils._run(abi.OP_VARINT_ZZ_DELTA, buf, pos, numValues)

    @staticmethod
    def decodeZigZagDeltaVarintCoordinates(buf, pos, numValues):  # :95
        return DecodingUtils._run(abi.OP_VARINT_ZZ_DELTA_XY, buf, pos, numValues)

    @staticmethod
    def decodeRle(buf, numValues, pos, signed):  # :257
        return DecodingUtils._run(abi.OP_RLE_S64 if signed else abi.OP_RLE_U64, buf, pos, numValues)

    @staticmethod
    def decodeByteRle(buf, numValues, pos, byteLength=None):  # :275 / :290
        return DecodingUtils._run(abi.OP_BYTE_RLE, buf, pos, numValues, byte_length=byteLength,
                                  advance_by_length=byteLength is not None)

    @staticmethod
    def decodeFastPfor128ZigZagDelta(buf, numValues, byteLength, pos):  # :316
        return DecodingUtils._run(abi.OP_PFOR_ZZ_DELTA, buf, pos, numValues, byte_length=byteLength, advance_by_length=True)

    @staticmethod
    def decodeFastPfor128DeltaCoordinates(buf, numValues, byteLength, pos):  # :349
        return DecodingUtils._run(abi.OP_PFOR_ZZ_DELTA_XY, buf, pos, numValues, byte_length=byteLength, advance_by_length=True)

    @staticmethod
    def decodeDeltaVarintMortonCodes(buf, pos, numVertices, numBits):  # :394
        return DecodingUtils._run(abi.OP_VARINT_DELTA_MORTON, buf, pos, numVertices, num_bits=numBits)

    @staticmethod
    def decodeFastPfor128DeltaMortonCodes(buf, numVertices, byteLength, pos, numBits):  # :411
        return DecodingUtils._run(abi.OP_PFOR_DELTA_MORTON, buf, pos, numVertices, byte_length=byteLength,
                                  num_bits=numBits, advance_by_length=True)


class CovtParser:
    """CovtParser.decodeCovt (J/decoder/CovtParser.java:53) for one tile; returns flat buffers per layer instead of
    JTS objects (SURVEY §8b output contract)."""

    @staticmethod
    def decodeCovt(covtBuffer, tileJson=None, container=abi.CONTAINER_GEN3, flags=abi.FLAG_DEFAULT):
        b = np.ascontiguousarray(np.frombuffer(covtBuffer, dtype=np.uint8) if not isinstance(covtBuffer, np.ndarray) else covtBuffer)
        dec = default_decoder()
        res = dec.decode_batch(b, np.array([0, b.size], dtype=np.uint64), container, flags, n_fields=tileJson)
        status, _ = res.tile_status()
        if status[0]:
            res.free()
            raise CovtError(int(status[0]), "tile decode failed")
        layers = []
        bufs = [res.buffer(i) for i in range(abi.NUM_BUFFERS - 1)]
        for L in res.layers:
            F = int(L["streams"][abi.SLOT_TYPES]["num_values"])
            o = L["out"]
            name = bytes(b[int(L["name_offset"]):int(L["name_offset"]) + int(L["name_length"])]).decode("utf-8", "replace")
            layers.append({
                "name": name if L["name_length"] else int(L["name_offset"]),
                "extent": int(L["extent"]),
                "geometry_types": bufs[abi.BUF_S_GEOMETRY_TYPES][int(o[abi.BUF_S_GEOMETRY_TYPES]):][:F],
                "ids": bufs[abi.BUF_S_IDS][int(o[abi.BUF_S_IDS]):][:F] if L["has_id"] else None,
                "geom_offsets": bufs[abi.BUF_A_GEOM_OFFSETS][int(o[abi.BUF_A_GEOM_OFFSETS]):][:F + 1],
                "part_offsets": bufs[abi.BUF_A_PART_OFFSETS][int(o[abi.BUF_A_PART_OFFSETS]):][:int(L["n_parts"]) + 1],
                "ring_offsets": bufs[abi.BUF_A_RING_OFFSETS][int(o[abi.BUF_A_RING_OFFSETS]):][:int(L["n_rings"]) + 1],
                "coords": bufs[abi.BUF_A_COORDS][int(o[abi.BUF_A_COORDS]):][:2 * int(L["n_coords"])],
            })
        res.free()
        return layers
from . import scheduler  # noqa: E402,F401  (one-process-per-GPU helpers over partition_tiles)
from .converter import CovtConverter  # noqa: E402,F401  (the tile writer over the GPU stream encoders)
