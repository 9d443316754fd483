"""GeoArrow / Arrow materialiser: one pyarrow Table per decoded layer, WITHOUT copying the decoded buffers.

The reference materialises `List<Layer>` of JTS geometries and `Map<String, List<Optional>>` properties on the host
(CovtParser.java:87-131). The library's results are already in Arrow layout — nested offsets + interleaved coordinates for the
geometry, a validity bitmap + one value slot per feature (or dictionary indices + utf8 offsets into the tile's own bytes) for
every property column — so materialising them for an Arrow consumer is a matter of wrapping slices of the host copies:

  id             int64
  geometry_type  uint8 (GeometryType ordinals, CovtParser.java:20-27)
  geometry       list<list<list<fixed_size_list<int32>[2]>>>   feature -> parts -> rings -> vertices (x, y): the GeoArrow
                 multipolygon nesting for every geometry type (a point is one part of one ring of one vertex)
  <property>     int64 / float / double / bool / dictionary<int32, utf8>, nulls where the feature has no value

Works on the product's Result (host copies of its buffers) and, because the oracle emits the same layout, on the test
infrastructure's results too — that is how tests/test_arrow.py checks it without a GPU. Needs pyarrow (not a dependency of the decode path).
"""
import numpy as np

from . import abi


def _pa():
    import pyarrow as pa
    return pa


def geometry_array(types, geom_off, part_off, ring_off, coords):
    """The nested GeoArrow array of one layer from its slices (see tests/canon.layer_slices for the slicing rule)."""
    pa = _pa()
    xy = pa.FixedSizeListArray.from_arrays(pa.array(np.ascontiguousarray(coords, dtype=np.int32)), 2)
    rings = pa.ListArray.from_arrays(pa.array(np.ascontiguousarray(ring_off, dtype=np.int32)), xy)
    parts = pa.ListArray.from_arrays(pa.array(np.ascontiguousarray(part_off, dtype=np.int32)), rings)
    return pa.ListArray.from_arrays(pa.array(np.ascontiguousarray(geom_off, dtype=np.int32)), parts)


def property_array(blob, c, validity, values, dict_offsets, dictionaries):
    """One decoded property column (a covt_prop_column record with status OK) as an Arrow array over the SAME bytes:
    validity = COVT_PBUF_VALIDITY, values = the buffer of the column's value_kind, blob = the batch blob (dictionary bytes)."""
    pa = _pa()
    F = int(c["num_features"])
    vo = int(c["validity_offset"])
    vbuf = pa.py_buffer(validity[vo:vo + (F + 7) // 8])
    kind, o = int(c["value_kind"]), int(c["values_offset"])
    if kind == abi.PV_BOOL:
        return pa.Array.from_buffers(pa.bool_(), F, [vbuf, pa.py_buffer(values[o:o + (F + 7) // 8])])
    if kind in (abi.PV_I64, abi.PV_F32, abi.PV_F64):
        t = {abi.PV_I64: pa.int64(), abi.PV_F32: pa.float32(), abi.PV_F64: pa.float64()}[kind]
        return pa.Array.from_buffers(t, F, [vbuf, pa.py_buffer(values[o:o + F])])
    if kind == abi.PV_DICT_INDEX:
        d = dictionaries[int(c["dictionary"])]
        oo, ne, bo, nb = int(d["offsets_offset"]), int(d["n_entries"]), int(d["bytes_offset"]), int(d["n_bytes"])
        words = pa.Array.from_buffers(pa.utf8(), ne, [None, pa.py_buffer(dict_offsets[oo:oo + ne + 1]), pa.py_buffer(blob[bo:bo + nb])])
        idx = pa.Array.from_buffers(pa.int32(), F, [vbuf, pa.py_buffer(values[o:o + F])])
        return pa.DictionaryArray.from_arrays(idx, words)
    raise ValueError("column of value kind %d has no Arrow form" % kind)


def _name(blob, off, n):
    return bytes(blob[int(off):int(off) + int(n)]).decode("utf-8", "replace")


def layer_tables(blob, layers, buffers, props=None, with_geometry=True):
    """blob: the batch blob (names and dictionary bytes live there); layers: covt_layer records (numpy structured array);
    buffers: indexable by abi.BUF_* -> host numpy array of that result buffer; props: an object with columns / dictionaries /
    validity / buffers[abi.PBUF_*] / dict_offsets (the product's property result or the oracle's), or None.
    Returns [(tile index, layer name, pyarrow.Table)] for every layer whose status is OK."""
    pa = _pa()
    by_layer = {}
    if props is not None:
        for i, c in enumerate(props.columns):
            if int(c["status"]) == abi.OK:
                by_layer.setdefault((int(c["tile"]), int(c["layer"])), []).append(i)
    out = []
    for L in layers:
        if int(L["status"]) != abi.OK:
            continue
        F = int(L["streams"][abi.SLOT_TYPES]["num_values"])
        o = L["out"]
        cols, names = [], []
        if int(L["has_id"]):
            ids = buffers[abi.BUF_S_IDS][int(o[abi.BUF_S_IDS]):][:F]
            cols.append(pa.array(np.ascontiguousarray(ids, dtype=np.int64)))
            names.append("id")
        types = buffers[abi.BUF_S_GEOMETRY_TYPES][int(o[abi.BUF_S_GEOMETRY_TYPES]):][:F]
        cols.append(pa.array(np.ascontiguousarray(types, dtype=np.uint8)))
        names.append("geometry_type")
        if with_geometry:
            g = buffers[abi.BUF_A_GEOM_OFFSETS][int(o[abi.BUF_A_GEOM_OFFSETS]):][:F + 1]
            p = buffers[abi.BUF_A_PART_OFFSETS][int(o[abi.BUF_A_PART_OFFSETS]):][:int(L["n_parts"]) + 1]
            r = buffers[abi.BUF_A_RING_OFFSETS][int(o[abi.BUF_A_RING_OFFSETS]):][:int(L["n_rings"]) + 1]
            c = buffers[abi.BUF_A_COORDS][int(o[abi.BUF_A_COORDS]):][:2 * int(L["n_coords"])]
            cols.append(geometry_array(types, g, p, r, c))
            names.append("geometry")
        for i in by_layer.get((int(L["tile"]), int(L["layer_index"])), []):
            c = props.columns[i]
            kind = int(c["value_kind"])
            values = {abi.PV_I64: abi.PBUF_I64, abi.PV_F32: abi.PBUF_F32, abi.PV_F64: abi.PBUF_F64, abi.PV_BOOL: abi.PBUF_BOOL,
                      abi.PV_DICT_INDEX: abi.PBUF_DICT_INDEX}.get(kind)
            if values is None:
                continue
            key = _name(blob, c["name_offset"], c["name_length"]) if int(c["name_length"]) else "field%d" % int(c["name_offset"])  # optimised gen-3: TileJSON field index
            if int(c["sub_length"]):  # localized dictionary (gen-2b fixtures): one column per sub-key
                sub = _name(blob, c["sub_offset"], c["sub_length"])
                key = key if sub == key else key + ":" + sub
            cols.append(property_array(blob, c, props.validity, props.buffers[values], props.dict_offsets, props.dictionaries))
            names.append(key if key not in names else key + "#%d" % i)
        lname = _name(blob, L["name_offset"], L["name_length"]) if int(L["name_length"]) else str(int(L["name_offset"]))
        out.append((int(L["tile"]), lname, pa.Table.from_arrays(cols, names=names)))
    return out
