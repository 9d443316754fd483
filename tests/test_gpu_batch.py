"""GPU parity tests of the batch path (covt_decode_batch / covt_batch_decode) against the CPU oracle,
the reference's own fixture tiles and the MVT ground-truth digests. All calls go through the C ABI."""
import numpy as np
import pytest

import canon
import util

pytestmark = pytest.mark.gpu


def _decode_both(covt, oracle, decoder, blob, offs, container=0, flags=None, n_fields=None):
    abi = covt.abi
    if flags is None:
        flags = abi.FLAG_DEFAULT
    res = decoder.decode_batch(blob, offs, container, flags, n_fields=n_fields)
    ref = oracle.decode_batch(blob, offs, container, flags, n_fields=n_fields)
    st, first = res.tile_status()
    assert np.array_equal(first, ref.first_layer), "first_layer differs"
    assert np.array_equal(st == 0, ref.tile_status == 0), "tile status OK-ness differs: %s vs %s" % (st, ref.tile_status)
    return res, ref


def test_config1_zoom5_transportation(covt, oracle, decoder, fixtures):
    """BASELINE config 1: the zoom-5 OMT transportation layer (and its three z5 siblings), every stream bit-exact."""
    abi = covt.abi
    tiles = [(n, b) for n, b in fixtures if n in ("omt/5_16_20", "omt/5_16_21", "omt/5_17_20", "omt/5_17_21")]
    assert len(tiles) == 4
    blob, offs = util.concat_tiles([b for _, b in tiles])
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, flags=flags)
    n = util.compare_results(abi, res, ref)
    names = [util.layer_name(blob, L) for L in res.layers]
    assert names.count("transportation") == 4
    # the layer BASELINE.md names: 45 232 features, 92 378 assembled vertices
    for L in res.layers:
        if util.layer_name(blob, L) == "transportation" and tiles[L["tile"]][0] == "omt/5_16_21":
            assert L["num_features"] == 45232 and L["n_vertices"] == 92378 and L["status"] == 0
    assert n == len(res.layers)
    res.free()


def test_config2_fixture_sweep_vs_oracle(covt, oracle, decoder, fixtures):
    """BASELINE config 2: all 129 gen-2b fixture tiles (omt z2-z14, bing, amazon) in ONE batched launch."""
    abi = covt.abi
    blob, offs = util.concat_tiles([b for _, b in fixtures])
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, flags=flags)
    n = util.compare_results(abi, res, ref)
    assert n >= 1375
    t = res.timing()
    assert t["payload_bytes"] == ref.payload_bytes and t["vertices"] == ref.vertices
    res.free()


@pytest.mark.parametrize("flags_extra", [0, 0x0004, 0x0008, 0x0010, 0x0020])
def test_fixture_sweep_flag_variants(covt, oracle, decoder, fixtures, flags_extra):
    """Quirk switches: no ring closing, MORTON_NO_SHIFT, ID_WIDTH_32, ICE_VB_COUNT_IS_INTS, SKIP_ASSEMBLY."""
    abi = covt.abi
    sub = [b for n, b in fixtures if n.startswith(("omt/8_", "omt/4_", "amazon/", "omt/14_"))]
    blob, offs = util.concat_tiles(sub)
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, flags=flags_extra | abi.FLAG_ID_DVZZ_IS_RLE)
    util.compare_results(abi, res, ref, check_assembled=not (flags_extra & abi.FLAG_SKIP_ASSEMBLY))
    res.free()


def test_fixture_geometry_equals_mvt(covt, decoder, fixtures):
    """Assembled geometry from the GPU equals the partner .mvt/.pbf of the reference fixtures (committed digests)."""
    abi = covt.abi
    digests = util.mvt_digests()
    checked = 0
    for z8 in (False, True):
        tiles = [(n, b) for n, b in fixtures if n.startswith("omt/8_") == z8 and not n.startswith("bing/")]
        blob, offs = util.concat_tiles([b for _, b in tiles])
        flags = abi.FLAG_ID_DVZZ_IS_RLE | (abi.FLAG_MORTON_NO_SHIFT if z8 else 0)
        res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
        bufs = [res.buffer(b) for b in range(abi.NUM_BUFFERS - 1)]
        for L in res.layers:
            key = "%s/%s" % (tiles[L["tile"]][0], util.layer_name(blob, L))
            if key in util.KNOWN_MISLABELLED:
                assert L["status"] != 0
                continue
            assert L["status"] == 0, key
            if key in util.KNOWN_MVT_MISMATCH or key not in digests:
                continue
            c = canon.canonical_from_assembled(*canon.layer_slices(L, bufs, abi))
            d = digests[key]
            assert (len(c[0]), len(c[1]), len(c[2]) // 2) == (d["features"], d["rings"], d["vertices"]), key
            assert canon.digest(c) == d["digest"], key
            checked += 1
        res.free()
    assert checked >= 1100


@pytest.mark.parametrize("props,id_last", [(False, False), (True, False), (True, True)])
def test_gen3_rewrapped_fixtures(covt, oracle, gen, decoder, fixtures, props, id_last):
    """gen-3 (HEAD CovtParser grammar) inputs made by re-wrapping the gen-2b fixtures decode to the same streams and geometry as
    the gen-2b originals; optimised metadata goes through the TileJSON side-car. props: the tiles keep their property columns
    as the HEAD converter writes them (unlisted Byte-RLE present streams, up to 40+ columns per layer: the device walker hops over
    every one of them); id_last: the id column follows the geometry column (columns are consumed in metadata order)."""
    abi = covt.abi
    names = ["omt/5_16_21", "omt/2_2_2", "omt/7_66_84", "omt/14_8298_10748", "amazon/5_5_11", "bing/4-8-5", "omt/12_2132_2734"]
    have = dict(fixtures)
    tiles = [have[n] for n in names if n in have]
    assert len(tiles) >= 4
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    for optimized in (False, True):
        wrapped = [util.rewrap_gen3(abi, oracle, t, optimized, props=props, gen=gen, id_last=id_last) for t in tiles]
        # one TileJSON for the batch: vector layer i carries the largest field count any tile needs for its layer i
        nf = None
        if optimized:
            nf = [0] * max(len(w[1]) for w in wrapped)
            for w in wrapped:
                for i, n in enumerate(w[1]):
                    nf[i] = max(nf[i], n)
        blob3, offs3 = util.concat_tiles([w[0] for w in wrapped])
        res3, ref3 = _decode_both(covt, oracle, decoder, blob3, offs3, container=abi.CONTAINER_GEN3, flags=flags, n_fields=nf)
        assert np.all(ref3.tile_status == 0)
        if props:
            assert int(ref3.layers["num_columns"].max()) > 10
        util.compare_results(abi, res3, ref3)
        # and against the gen-2b decode of the same tiles: identical decoded buffers
        blob2, offs2 = util.concat_tiles(tiles)
        res2 = decoder.decode_batch(blob2, offs2, abi.CONTAINER_GEN2B, flags)
        util.compare_results(abi, res3, res2, same_container=False)
        res2.free()
        res3.free()


def test_id_column_after_geometry_gen2b(covt, oracle, gen, decoder):
    """Columns are consumed in METADATA order (CovtParser.java:64-85): a gen-2b tile whose id column follows the geometry column
    carries the id payload behind the geometry payload."""
    abi = covt.abi
    blob, offs, truth = gen.tiles(77, 40, gen.default_params())
    from oracle import properties as P
    tiles = []
    for i in range(40):
        t = bytes(blob[int(offs[i]):int(offs[i + 1])])
        tiles.append(util.swap_id_and_geometry_gen2b(P, t))
    blob2, offs2 = util.concat_tiles(tiles)
    res, ref = _decode_both(covt, oracle, decoder, blob2, offs2)
    assert np.all(ref.tile_status == 0) and ref.layers["has_id"].all()
    L = ref.layers[0]
    assert L["streams"][abi.SLOT_ID]["byte_offset"] > L["streams"][abi.SLOT_VBUF]["byte_offset"]
    util.compare_results(abi, res, ref)
    ref0 = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    util.compare_results(abi, res, ref0, same_container=False)
    res.free()


@pytest.mark.parametrize("container", [0, 1, 2])
def test_synthetic_tiles_vs_oracle_and_truth(covt, oracle, gen, decoder, container):
    """Config-5 style synthetic mixed-geometry tiles: GPU == oracle bit for bit, and both == what was encoded."""
    abi = covt.abi
    n_tiles = 3000
    p = gen.default_params(container=container)
    blob, offs, truth = gen.tiles(1000, n_tiles, p)
    cont = abi.CONTAINER_GEN2B if container == 0 else abi.CONTAINER_GEN3
    nf = [0] * 16 if container == 2 else None
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, container=cont, n_fields=nf)
    assert np.all(ref.tile_status == 0)
    util.compare_results(abi, res, ref)
    L = res.layers
    assert int(L["n_vertices"].sum()) == truth["vertices"]
    assert int(L["n_parts"].sum()) == truth["parts"] and int(L["n_rings"].sum()) == truth["rings"]
    assert int(L["num_features"].sum()) == truth["features"]
    # order-insensitive checksum of all assembled coordinates (closing vertices included)
    coords = res.buffer(abi.BUF_A_COORDS)
    sx = sy = 0
    for row in L:
        o, n = int(row["out"][abi.BUF_A_COORDS]), int(row["n_coords"])
        c = coords[o:o + 2 * n].astype(np.int64)
        sx += int(c[0::2].sum())
        sy += int(c[1::2].sum())
    assert (sx, sy) == (truth["sum_x_closed"], truth["sum_y_closed"])
    res.free()


def test_synthetic_index_buffer_config4(covt, oracle, gen, decoder):
    """Config 4: polygon-heavy tiles carrying the INDEX_BUFFER extension stream (FAST_PFOR_DELTA_ZIG_ZAG); parity
    is against the oracle only (the reference defines IndexBuffer in prose, README.md:114-121)."""
    abi = covt.abi
    p = gen.default_params(with_index_buffer=1, p_point=0.05, p_line=0.15, p_polygon=0.6, p_multiline=0.05,
                           p_multipolygon=0.15, mean_features=120, mean_ring_extra=12)
    blob, offs, truth = gen.tiles(7, 600, p)
    res, ref = _decode_both(covt, oracle, decoder, blob, offs)
    util.compare_results(abi, res, ref)
    assert res.device_buffer(abi.BUF_S_INDEX_BUFFER)[1] > 0
    res.free()


def test_large_layers(covt, oracle, gen, decoder):
    """Layers far larger than one warp chunk / FastPFOR page (multi-page streams, > 65 536 values)."""
    abi = covt.abi
    p = gen.default_params(mean_features=30000, layers_per_tile=2)
    blob, offs, truth = gen.tiles(5, 6, p)
    res, ref = _decode_both(covt, oracle, decoder, blob, offs)
    util.compare_results(abi, res, ref)
    assert int(res.layers["n_vertices"].sum()) == truth["vertices"]
    res.free()


def test_malformed_tiles_do_not_poison_the_batch(covt, oracle, gen, decoder, fixtures):
    """Truncated tiles, garbage, an empty tile and a gen-2a tile between good ones: per-tile status, neighbours intact."""
    abi = covt.abi
    good = dict(fixtures)["omt/5_17_21"]
    rng = np.random.default_rng(5)
    tiles = [good, good[: len(good) // 2], bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), b"", good[:37], good,
             good[:-1], bytes(200), good + b"\x00"]
    blob, offs = util.concat_tiles(tiles)
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, flags=flags)
    st, _ = res.tile_status()
    assert st[0] == 0 and st[5] == 0 and st[1] != 0 and st[3] != 0
    util.compare_results(abi, res, ref)
    res.free()
    # corrupt single bytes inside stream payloads: status may be anything, but the call must survive and
    # every tile the oracle accepts must match
    for seed in range(6):
        r = np.random.default_rng(seed)
        b = bytearray(good)
        for _ in range(4):
            b[int(r.integers(200, len(b)))] ^= int(r.integers(1, 256))
        blob, offs = util.concat_tiles([good, bytes(b), good])
        res = decoder.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
        ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
        st, _ = res.tile_status()
        assert st[0] == 0 and st[2] == 0
        ok = (res.layers["status"] == 0) & (ref.layers["status"] == 0)
        assert ok.sum() >= 2 * (len(ok) // 3)
        res.free()


def test_empty_batch_and_split_api(covt, oracle, decoder, fixtures):
    abi = covt.abi
    res = decoder.decode_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert res.n_tiles == 0 and res.n_layers == 0
    res.free()
    # upload once, decode twice (device-resident timing path): identical results
    blob, offs = util.concat_tiles([b for n, b in fixtures if n.startswith("omt/6_")])
    batch = decoder.upload(blob, offs)
    r1 = decoder.decode(batch, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE | abi.FLAG_PROFILE_KERNELS)
    r2 = decoder.decode(batch, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE)
    util.compare_results(abi, r1, r2)
    kt = {k["name"]: k for k in r1.kernel_times()}
    assert "k_assemble_layers" in kt and kt["k_assemble_layers"]["ms"] > 0 and "k_decode_varint32" in kt
    t = r1.timing()
    assert t["decode_ms"] > 0 and t["payload_bytes"] > 0 and t["vertices"] > 0
    r1.free()
    r2.free()
    batch.free()


def test_pipelined_segments_equal_one_shot(covt, oracle, gen, fixtures, monkeypatch):
    """covt_decode_batch with host input uploads and decodes in segments (upload of segment i+1 overlaps the decode of segment i;
    result capacities are extrapolated from segment 0). Whatever the segmentation, the result is bit-identical to the oracle's."""
    abi = covt.abi
    monkeypatch.setenv("COVT_SEG_BYTES", str(1 << 20))
    monkeypatch.setenv("COVT_MAX_SEGMENTS", "64")
    monkeypatch.setenv("COVT_SEG_MIN_TILES", "16")
    dec = covt.Decoder(0)
    try:
        blob, offs, truth = gen.tiles(31, 4000, gen.default_params())
        res = dec.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
        t = res.timing()
        assert t["segments"] >= 4 and t["capacity_retries"] == 0
        ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
        st, first = res.tile_status()
        assert np.array_equal(first, ref.first_layer) and np.all(st == 0)
        util.compare_results(abi, res, ref)
        assert int(res.layers["n_vertices"].sum()) == truth["vertices"] and t["vertices"] == truth["vertices"]
        res.free()
        # a batch whose first segment is NOT representative (small tiles first, the big fixture tiles last): the extrapolated
        # capacities overflow on the device, and the call transparently decodes again with exact sizes
        small, soffs, _ = gen.tiles(5, 1500, gen.default_params(mean_features=4))
        big = [b for n, b in fixtures if n.startswith(("omt/5_", "omt/6_", "omt/7_"))]
        blob2, offs2 = util.concat_tiles([bytes(small[int(soffs[i]):int(soffs[i + 1])]) for i in range(1500)] + big)
        flags = abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE
        res2 = dec.decode_batch(blob2, offs2, abi.CONTAINER_GEN2B, flags)
        assert res2.timing()["capacity_retries"] == 1
        ref2 = oracle.decode_batch(blob2, offs2, abi.CONTAINER_GEN2B, flags)
        util.compare_results(abi, res2, ref2)
        res2.free()
    finally:
        dec.close()


def test_partitioned_decode_and_trim(covt, oracle, gen):
    """The batch scheduler's per-rank slices (cov-tiles_b200/scheduler.py), decoded one after the other on this GPU, cover the
    batch exactly; covt_trim hands the parked device blocks back without disturbing later calls."""
    abi = covt.abi
    from cov_tiles_b200 import scheduler
    dec = covt.Decoder(0)
    try:
        blob, offs, truth = gen.tiles(900, 1500, gen.default_params())
        whole = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
        verts = 0
        for rank in range(3):
            res, t0 = scheduler.decode_partitioned(dec, blob, offs, rank, 3)
            sub, sub_offs, t0b = scheduler.rank_slice(blob, offs, rank, 3)
            assert t0 == t0b
            ref = oracle.decode_batch(sub, sub_offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
            util.compare_results(abi, res, ref)
            rows = whole.layers[(whole.layers["tile"] >= t0) & (whole.layers["tile"] < t0 + len(sub_offs) - 1)]
            assert np.array_equal(rows["n_vertices"], res.layers["n_vertices"])
            verts += int(res.layers["n_vertices"].sum())
            res.free()
            if rank == 1:
                dec.trim()
        assert verts == truth["vertices"]
    finally:
        dec.close()


def _mutants(rng, base, n_mutants, good_every):
    tiles, is_good = [], []
    for k in range(n_mutants):
        data = base[k % len(base)]
        b = bytearray(data)
        kind = k % 4
        if kind == 0:      # anywhere
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= int(rng.integers(1, 256))
        elif kind == 1:    # the first layer's metadata
            for _ in range(int(rng.integers(1, 3))):
                b[int(rng.integers(0, min(len(b), 160)))] ^= int(rng.integers(1, 256))
        elif kind == 2:    # truncated
            b = b[: int(rng.integers(0, len(b)))]
        else:              # junk appended / a byte dropped
            if rng.integers(0, 2):
                b += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            else:
                del b[int(rng.integers(0, len(b)))]
        tiles.append(bytes(b))
        is_good.append(False)
        if k % good_every == 0:
            tiles.append(data)
            is_good.append(True)
    return tiles, np.array(is_good)


@pytest.mark.gpu
@pytest.mark.parametrize("corpus", ["small_gen2b", "large_gen2b", "small_gen3", "small_gen3_props"])
def test_mutation_fuzz_against_oracle(covt, oracle, gen, decoder, fixtures, corpus):
    """Mutants of fixture tiles (byte flips in metadata and payload, truncations, appended junk, dropped bytes) in ONE batch
    between good tiles: the call survives, tile / layer / stream statuses agree with the oracle on OK-ness, the result layout is
    identical, every stream and layer both sides accept is bit-exact, and the good tiles are untouched. small = 1 200 mutants
    of the five smallest OMT tiles (lane decoders, shared-memory FastPFOR), large = 160 mutants of the zoom-5 tiles (the
    second-pass warp decoders of 10^4..10^5-value streams), gen3 = the small tiles re-wrapped in the HEAD container."""
    abi = covt.abi
    flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
    clean = [(n, b) for n, b in fixtures if n.startswith("omt/") and not n.startswith("omt/8_")
             and not any(k.startswith(n + "/") for k in util.KNOWN_MISLABELLED)]
    container = abi.CONTAINER_GEN2B
    if corpus == "large_gen2b":
        base = [b for n, b in clean if n.startswith("omt/5_")]
        n_mutants, good_every = 160, 20
    else:
        base = [b for _, b in sorted(clean, key=lambda t: len(t[1]))[:5]]
        n_mutants, good_every = 1200, 50
        if corpus.startswith("small_gen3"):
            # (_props: the property columns stay in, so that mutations also hit the unlisted present streams and their hop-over)
            base = [util.rewrap_gen3(abi, oracle, b, props=corpus.endswith("_props"), gen=gen)[0] for b in base]
            container = abi.CONTAINER_GEN3
    assert len(base) >= 3
    tiles, good = _mutants(np.random.default_rng(2026), base, n_mutants, good_every)
    blob, offs = util.concat_tiles(tiles)
    res, ref = _decode_both(covt, oracle, decoder, blob, offs, container=container, flags=flags)
    st, _ = res.tile_status()
    assert not st[good].any(), "a good tile was poisoned by its neighbours"
    assert (st[~good] != 0).sum() > n_mutants // 5 and (st[~good] == 0).sum() > n_mutants // 20  # both outcomes are exercised
    n = util.compare_results(abi, res, ref)
    assert n > n_mutants // 2
    res.free()


@pytest.mark.gpu
def test_config2_rle_topology_variant(covt, oracle, gen, decoder, fixtures):
    """BASELINE config 2 names "RLE topology streams": the 91 OMT tiles with their 273 FastPFOR topology streams transcoded
    to ORC RLE (tools/gen/rewrite.py; the FastPFOR side decoded by the product itself, as bench.py --rle-topology does).
    GPU == oracle on the rewritten batch, and every result buffer equals the one of the original batch."""
    from tools.gen import rewrite
    abi = covt.abi
    names = [n for n, _ in fixtures if n.startswith("omt/")]
    tiles = [b for n, b in fixtures if n.startswith("omt/")]
    new_tiles, n_streams = [], 0
    for data in tiles:
        arr = np.frombuffer(data + bytes(64), dtype=np.uint8)
        todo = rewrite.topology_pfor_streams(data)
        decoded = {}
        if todo:
            descs = (abi.StreamDesc * len(todo))()
            for i, (off, bl, nv) in enumerate(todo):
                descs[i] = abi.StreamDesc(byte_offset=off, byte_length=bl, num_values=nv, op=abi.OP_PFOR_ZZ_DELTA)
            r = decoder.decode_streams(arr, descs, abi.FLAG_DEFAULT)
            arena = r.buffer(abi.BUF_STREAM_ARENA)
            for i, (off, bl, nv) in enumerate(todo):
                assert descs[i].status == 0 and descs[i].out_count == nv
                decoded[off] = arena[descs[i].out_offset:descs[i].out_offset + 4 * nv].view(np.int32).copy()
            r.free()
        new_tiles.append(rewrite.transcode_topology_to_rle(data, decoded))
        n_streams += len(todo)
    assert n_streams >= 270
    for z8 in (False, True):
        sel = [i for i, n in enumerate(names) if n.startswith("omt/8_") == z8]
        flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE | (abi.FLAG_MORTON_NO_SHIFT if z8 else 0)
        blob, offs = util.concat_tiles([new_tiles[i] for i in sel])
        res, ref = _decode_both(covt, oracle, decoder, blob, offs, flags=flags)
        util.compare_results(abi, res, ref)
        enc = res.layers["streams"]["encoding"][:, [abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING]]
        assert not (enc == abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG).any()
        blob0, offs0 = util.concat_tiles([tiles[i] for i in sel])
        res0 = decoder.decode_batch(blob0, offs0, abi.CONTAINER_GEN2B, flags)
        assert np.array_equal(res.layers["out"], res0.layers["out"]) and np.array_equal(res.layers["status"], res0.layers["status"])
        # valid slices only (padding between slices is unspecified): the bulk comparison with the original batch as reference
        util.compare_results_bulk(abi, res, res0, chunk_layers=4096, same_container=False)
        res.free()
        res0.free()


@pytest.mark.gpu
@pytest.mark.parametrize("two_gpus", [False, True])
def test_multi_gpu_scheduler_one_call(covt, oracle, gen, fixtures, two_gpus):
    """covt_decode_batch_multi: ONE call, the library cuts the batch into contiguous tile ranges balanced by payload bytes and
    decodes them side by side, one context + host thread per part. two_gpus=False: three contexts on GPU 0 (what a 1-GPU box
    can run); True: contexts on GPUs 0 and 1 in ONE process (skipped below 2 GPUs). Every part equals the oracle's decode of
    the same tile range, the parts cover the batch exactly, and a concurrent device->host read-back of a buffer works."""
    import torch
    abi = covt.abi
    if two_gpus and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ids = [0, 1] if two_gpus else [0, 0, 0]
    md = covt.MultiDecoder(ids)
    try:
        assert md.n_devices == len(ids)
        small, soffs, truth = gen.tiles(123, 2500, gen.default_params())
        big = [b for n, b in fixtures if n.startswith(("omt/5_", "omt/9_"))]
        tiles = [bytes(small[int(soffs[i]):int(soffs[i + 1])]) for i in range(2500)] + big
        blob, offs = util.concat_tiles(tiles)
        flags = abi.FLAG_DEFAULT | abi.FLAG_ID_DVZZ_IS_RLE
        for rep in range(2):  # the second call reuses every context's parked blocks
            mr = md.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
            assert len(mr.parts) == len(ids)
            starts = covt.partition_tiles(offs, len(ids))
            nxt = 0
            verts = 0
            for k, part in enumerate(mr.parts):
                assert part["first_tile"] == nxt == int(starts[k]) and part["device"] == ids[k]
                t0, n = part["first_tile"], part["n_tiles"]
                nxt = t0 + n
                b0 = int(offs[t0])
                ref = oracle.decode_batch(blob[b0:int(offs[t0 + n])], offs[t0:t0 + n + 1] - np.uint64(b0), abi.CONTAINER_GEN2B, flags)
                res = part["result"]
                st, first = res.tile_status()
                assert np.array_equal(first, ref.first_layer) and np.array_equal(st == 0, ref.tile_status == 0)
                util.compare_results(abi, res, ref)
                verts += int(res.layers["n_vertices"].sum())
            assert nxt == len(tiles)
            t = mr.timing()
            assert t["vertices"] == verts and t["payload_bytes"] > 0 and t["decode_ms"] > 0
            # all parts' coordinates back to the host at once
            counts = [p["result"].device_buffer(abi.BUF_A_COORDS)[1] for p in mr.parts]
            hosts = [np.empty(max(c, 1), np.int32) for c in counts]
            mr.read_into(abi.BUF_A_COORDS, [h.ctypes.data for h in hosts])
            for p, h, c in zip(mr.parts, hosts, counts):
                assert np.array_equal(h[:c], p["result"].buffer(abi.BUF_A_COORDS))
            mr.free()
    finally:
        md.close()


@pytest.mark.gpu
def test_one_context_per_thread(covt, oracle, gen):
    """The threading contract of include/covt_b200.h: a context serves one thread at a time, different contexts run concurrently."""
    import threading
    abi = covt.abi
    blob, offs, truth = gen.tiles(500, 1200, gen.default_params())
    ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
    errors = []

    def work(k):
        try:
            dec = covt.Decoder(0)
            for _ in range(3):
                res = dec.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_DEFAULT)
                util.compare_results(abi, res, ref)
                res.free()
            dec.close()
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))
    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_decode_batch_to_host_delivers_the_same_buffers(covt, oracle, fixtures, gen, monkeypatch):
    """covt_decode_batch_to_host: the buffers named by the sink land in host memory, read back segment by segment while later
    segments are uploaded and decoded — same bytes as covt_result_read of the device-resident result, in one and in many segments,
    and through the capacity retry; a sink that is too small fails the call."""
    abi = covt.abi
    synth_blob, synth_offs, _ = gen.tiles(7000, 3000, gen.default_params())
    fx = util.concat_tiles([b for n, b in fixtures if n.startswith("omt/")] * 2)
    for blob, offs, seg_bytes in ((synth_blob, synth_offs, None), (synth_blob, synth_offs, 1 << 19), (fx[0], fx[1], 1 << 20)):
        if seg_bytes:
            monkeypatch.setenv("COVT_SEG_BYTES", str(seg_bytes))
            monkeypatch.setenv("COVT_MAX_SEGMENTS", "64")
            monkeypatch.setenv("COVT_SEG_MIN_TILES", "1")
        else:
            monkeypatch.delenv("COVT_SEG_BYTES", raising=False)
        dec = covt.Decoder(0)
        try:
            flags = abi.FLAG_CLOSE_RINGS | abi.FLAG_ID_DVZZ_IS_RLE
            ref = dec.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
            wanted = [abi.BUF_S_IDS, abi.BUF_S_GEOMETRY_TYPES, abi.BUF_A_GEOM_OFFSETS, abi.BUF_A_PART_OFFSETS, abi.BUF_A_RING_OFFSETS, abi.BUF_A_COORDS]
            host = {b: np.full(ref.device_buffer(b)[1] + 5, 0x5A, dtype=abi.BUF_DTYPES[b]) for b in wanted}
            res = dec.decode_batch_to_host(blob, offs, host, abi.CONTAINER_GEN2B, flags)
            t = res.timing()
            if seg_bytes:
                assert t["segments"] >= 4 or t["capacity_retries"] == 1
            for b in wanted:
                n = ref.device_buffer(b)[1]
                assert res.device_buffer(b)[1] == n
                # (byte-identical to the device-resident buffer of the SAME result: the alignment gaps between layer slices are never written)
                assert np.array_equal(host[b][:n], res.buffer(b)), abi.BUF_NAMES[b]
                assert (host[b][n:] == 0x5A).all()  # nothing written behind the buffer's end
            util.compare_results(abi, res, oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags))
            res.free()
            small = dict(host)
            small[abi.BUF_A_COORDS] = np.zeros(max(ref.device_buffer(abi.BUF_A_COORDS)[1] - 1, 0), dtype=np.int32)
            with pytest.raises(covt.CovtError):
                dec.decode_batch_to_host(blob, offs, small, abi.CONTAINER_GEN2B, flags)
            ref.free()
        finally:
            dec.close()


def test_decreasing_tile_offsets_are_rejected(covt, gen, monkeypatch):
    """tile_offsets are checked segment by segment while the upload is already running: a batch whose offsets go backwards anywhere
    fails with INVALID_ARG (one segment and many), and the context decodes the next batch as if nothing had happened."""
    abi = covt.abi
    blob, offs, _ = gen.tiles(100, 4000, gen.default_params())
    for seg_bytes in (None, 1 << 18):
        if seg_bytes:
            monkeypatch.setenv("COVT_SEG_BYTES", str(seg_bytes))
            monkeypatch.setenv("COVT_MAX_SEGMENTS", "64")
            monkeypatch.setenv("COVT_SEG_MIN_TILES", "1")
        dec = covt.Decoder(0)
        try:
            for at in (1, 2000, 3999):
                bad = offs.copy()
                bad[at] = bad[at + 1] + 7 if at + 1 < len(bad) - 1 else bad[at - 1] - 1
                with pytest.raises(covt.CovtError) as ei:
                    dec.decode_batch(blob, bad, abi.CONTAINER_GEN2B, abi.FLAG_CLOSE_RINGS)
                assert ei.value.code == abi.ERR_INVALID_ARG
            res = dec.decode_batch(blob, offs, abi.CONTAINER_GEN2B, abi.FLAG_CLOSE_RINGS)
            st, _ = res.tile_status()
            assert not st.any()
            res.free()
        finally:
            dec.close()
