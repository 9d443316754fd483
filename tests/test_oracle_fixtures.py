"""CPU: pins the oracle (oracle/covt_oracle.c, the checker of every GPU parity test) against the reference's own data.

The JVM is absent, so the Java decoder cannot run here; what the reference ships instead are 129 gen-2b `.covt`
fixture tiles with `.mvt`/`.pbf` partners (test/fixtures/{omt,amazon,bing}) — committed under tests/golden/ by
tests/golden/make_golden.py. The oracle must
  * walk every tile's container EOF-exactly,
  * consume exactly the declared byteLength of every stream (17 758 RLE, 14 879 Byte-RLE, 1 045 FastPFOR, ... streams),
  * assemble geometry equal to the partner MVT (blake2b digests of the canonical form, tests/canon.py),
  * and its decoded values must re-encode (tools/gen = restatement of EncodingUtils.java) to the very bytes in the fixture.
"""
import numpy as np

import canon
import util


def _flags(abi, name):
    return util.fixture_flags(abi, name)


def test_container_walk_is_eof_exact(oracle, fixtures):
    abi = oracle.abi
    n_layers = 0
    for name, data in fixtures:
        rc, layers, end_pos = oracle.parse_tile(np.frombuffer(data, np.uint8), abi.CONTAINER_GEN2B, flags=_flags(abi, name))
        assert rc == 0, name
        assert end_pos == len(data), "%s: walk ended at %d of %d" % (name, end_pos, len(data))
        n_layers += len(layers)
    assert len(fixtures) == 129 and n_layers >= 1375


def test_every_stream_consumes_its_declared_length_and_reencodes(oracle, gen, fixtures):
    abi = oracle.abi
    seen = {}
    identical = {}
    for name, data in fixtures:
        blob = np.frombuffer(data, np.uint8)
        flags = _flags(abi, name)
        rc, layers, _ = oracle.parse_tile(blob, abi.CONTAINER_GEN2B, flags=flags)
        for L in layers:
            key = "%s/%s" % (name, util.layer_name(blob, L))
            for s in range(abi.NUM_SLOTS):
                S = L["streams"][s]
                if S["encoding"] == abi.ENC_ABSENT or S["op"] == abi.OP_NONE:
                    continue
                op, off, bl, nv = int(S["op"]), int(S["byte_offset"]), int(S["byte_length"]), int(S["num_values"])
                if s == abi.SLOT_VBUF and L["geom_column_type"] == abi.CT_ICE:
                    nv *= 2  # encoder wrote #vertices (SURVEY §8a dispatch table)
                vals, st, cons = oracle.decode_stream(blob, op, byte_offset=off, byte_length=bl, num_values=nv,
                                                      num_bits=int(L["num_bits"]), flags=flags)
                cname = abi.OP_NAMES[op].upper()
                if key in util.KNOWN_MISLABELLED and s == abi.SLOT_VBUF:
                    assert st != 0, key
                    continue
                assert st == 0, "%s slot %d op %s status %d" % (key, s, cname, st)
                assert cons == bl, "%s slot %d op %s consumed %d of %d" % (key, s, cname, cons, bl)
                seen[cname] = seen.get(cname, 0) + 1
                # re-encode with the restated EncodingUtils encoders: must give back the fixture bytes
                payload = blob[off:off + bl]
                enc = None
                if op == abi.OP_BYTE_RLE:
                    enc = gen.encode_byte_rle(vals)
                elif op in (abi.OP_RLE_U32, abi.OP_RLE_U64):
                    enc = gen.encode_rle(vals.astype(np.int64), signed=False)
                elif op == abi.OP_VARINT_ZZ_DELTA:
                    enc = gen.encode_varints(vals.astype(np.int64), zigzag=True, delta=True)
                elif op == abi.OP_VARINT_ZZ_DELTA_XY:
                    zz = gen.encode_zigzag_delta_coordinates(vals.astype(np.int32)).astype(np.int64) & 0xFFFFFFFF
                    enc = gen.encode_varints(zz)
                elif op == abi.OP_PFOR_ZZ_DELTA:
                    enc = gen.encode_fastpfor(vals.astype(np.int32), zigzag=True, delta=True)
                elif op == abi.OP_PFOR_ZZ_DELTA_XY:
                    enc = gen.encode_fastpfor(gen.encode_zigzag_delta_coordinates(vals.astype(np.int32)), zigzag=False, delta=False)
                if enc is None:
                    continue
                same = len(enc) == bl and np.array_equal(enc, payload)
                tot, ok = identical.get(cname, (0, 0))
                identical[cname] = (tot + 1, ok + int(same))
                if not same and not (cname.startswith("PFOR") and nv > 65536):
                    # JavaFastPFOR leaks stale exception-array contents into padding bits of pages >= 2
                    raise AssertionError("%s slot %d op %s: re-encoding differs (%d vs %d bytes)" % (key, s, cname, len(enc), bl))
    # the fixture corpus exercises every codec family of the path
    for fam in ("BYTE_RLE", "RLE_U32", "RLE_U64", "VARINT_ZZ_DELTA", "VARINT_DELTA_MORTON", "PFOR_ZZ_DELTA"):
        assert seen.get(fam, 0) > 0, (fam, seen)
    assert seen["BYTE_RLE"] >= 1375 and seen["RLE_U32"] + seen["RLE_U64"] >= 1000
    pf_tot = sum(v[0] for k, v in identical.items() if k.startswith("PFOR"))
    pf_ok = sum(v[1] for k, v in identical.items() if k.startswith("PFOR"))
    assert pf_tot >= 900 and pf_tot - pf_ok <= 30, identical
    for k, (tot, ok) in identical.items():
        if not k.startswith("PFOR"):
            assert tot == ok, (k, tot, ok)


def test_assembled_geometry_equals_partner_mvt(oracle, fixtures):
    abi = oracle.abi
    digests = util.mvt_digests()
    checked = 0
    for z8 in (False, True):
        tiles = [(n, b) for n, b in fixtures if n.startswith("omt/8_") == z8 and not n.startswith("bing/")]
        blob, offs = util.concat_tiles([b for _, b in tiles])
        flags = abi.FLAG_ID_DVZZ_IS_RLE | (abi.FLAG_MORTON_NO_SHIFT if z8 else 0)
        ref = oracle.decode_batch(blob, offs, abi.CONTAINER_GEN2B, flags)
        assert np.all(ref.tile_status[[i for i, (n, _) in enumerate(tiles)
                                        if not any(k.startswith(n + "/") for k in util.KNOWN_MISLABELLED)]] == 0)
        for L in ref.layers:
            key = "%s/%s" % (tiles[L["tile"]][0], util.layer_name(blob, L))
            if key in util.KNOWN_MISLABELLED:
                assert L["status"] != 0
                continue
            assert L["status"] == 0, key
            if key in util.KNOWN_MVT_MISMATCH or key not in digests:
                continue
            c = canon.canonical_from_assembled(*canon.layer_slices(L, ref.buffers, abi))
            d = digests[key]
            assert (len(c[0]), len(c[1]), len(c[2]) // 2) == (d["features"], d["rings"], d["vertices"]), key
            assert canon.digest(c) == d["digest"], key
            checked += 1
    assert checked >= 1100


def test_config1_layer_shape(oracle, fixtures):
    """BASELINE config 1 (SURVEY §8a sizes): transportation layer of omt/5_16_21."""
    abi = oracle.abi
    data = dict(fixtures)["omt/5_16_21"]
    blob = np.frombuffer(data, np.uint8)
    ref = oracle.decode_batch(blob, np.array([0, len(blob)], np.uint64), abi.CONTAINER_GEN2B, abi.FLAG_ID_DVZZ_IS_RLE)
    for L in ref.layers:
        if util.layer_name(blob, L) == "transportation":
            assert L["num_features"] == 45232 and L["n_vertices"] == 92378
            assert L["streams"][abi.SLOT_VOFF]["num_values"] == 92378 and L["streams"][abi.SLOT_VOFF]["byte_length"] == 121267
            assert L["streams"][abi.SLOT_VBUF]["num_values"] == 19704 and L["streams"][abi.SLOT_VBUF]["byte_length"] == 26176
            assert L["streams"][abi.SLOT_PART]["num_values"] == 45234 and L["streams"][abi.SLOT_PART]["byte_length"] == 3504
            assert L["streams"][abi.SLOT_TYPES]["byte_length"] == 702 and L["streams"][abi.SLOT_ID]["byte_length"] == 1656
            return
    raise AssertionError("transportation layer not found")


def test_rle_topology_variant_decodes_to_the_same_geometry(oracle, gen, fixtures):
    """BASELINE config 2 names "RLE topology streams": tools/gen/rewrite.py transcodes the 273 FastPFOR topology streams of the
    91 OMT fixtures to ORC RLE (metadata + payload re-serialised). The oracle must decode the rewritten tiles to exactly the
    buffers of the originals (which test_assembled_geometry_equals_partner_mvt pins on the MVT), and a tile without FastPFOR
    topology must come back byte for byte."""
    from tools.gen import rewrite
    abi = oracle.abi
    n_streams = n_identical = 0
    for name, data in fixtures:
        if not name.startswith("omt/"):
            continue
        arr = np.frombuffer(data + bytes(64), np.uint8)
        todo = rewrite.topology_pfor_streams(data)
        decoded = {}
        for off, bl, nv in todo:
            vals, st, cons = oracle.decode_stream(arr, abi.OP_PFOR_ZZ_DELTA, byte_offset=off, byte_length=bl, num_values=nv)
            assert st == 0 and cons == bl
            decoded[off] = vals
        new = rewrite.transcode_topology_to_rle(data, decoded)
        n_streams += len(todo)
        if not todo:
            assert new == data
            n_identical += 1
            continue
        flags = _flags(abi, name)
        a = oracle.decode_batch(np.frombuffer(data, np.uint8), np.array([0, len(data)], np.uint64), abi.CONTAINER_GEN2B, flags)
        b = oracle.decode_batch(np.frombuffer(new, np.uint8), np.array([0, len(new)], np.uint64), abi.CONTAINER_GEN2B, flags)
        assert np.array_equal(a.tile_status, b.tile_status) and np.array_equal(a.layers["status"], b.layers["status"])
        assert np.array_equal(a.layers["out"], b.layers["out"])
        for k in range(abi.NUM_BUFFERS - 1):
            assert np.array_equal(a.buffer(k), b.buffer(k)), (name, abi.BUF_NAMES[k])
        enc = b.layers["streams"]["encoding"][:, [abi.SLOT_GEOM, abi.SLOT_PART, abi.SLOT_RING]]
        assert not (enc == abi.ENC_FAST_PFOR_DELTA_ZIG_ZAG).any()
    assert n_streams >= 270 and n_identical >= 1


def test_gen2b_grammar_reserialises_every_fixture_byte_for_byte(fixtures):
    """No gen-2b reader or writer exists in the reference at HEAD: the container grammar was established from the
    fixture bytes. tools/gen/rewrite.py walks a tile with that grammar (every column, property columns included) and writes it
    again from the parsed fields + payload slices: all 129 tiles (OMT, Amazon, Bing) must come back byte for byte."""
    from tools.gen import rewrite
    n_streams = 0
    for name, data in fixtures:
        assert rewrite.transcode_topology_to_rle(data, {}, keep_pfor=True) == data, name
        n_streams += sum(len(c["streams"]) for L in rewrite.walk(data)[1] for c in L["columns"])
    assert len(fixtures) == 129 and n_streams == 38449  # the census of SURVEY §4.4


def test_gen3_walk_with_property_columns(oracle, gen, fixtures):
    """gen-3 (HEAD grammar, CovtParser.java:574-652) tiles that KEEP their property columns the way the HEAD converter writes them
    (unlisted Byte-RLE present streams, CovtConverter.java:434-436; up to 40+ columns per layer): the oracle's walk hops over
    every property payload, lands on EOF, and decodes the same geometry and ids as from the gen-2b original — with optimised and
    plain metadata, and with the id column before or after the geometry column (columns are consumed in metadata order,
    CovtParser.java:64-85)."""
    abi = oracle.abi
    pick = [(n, b) for n, b in fixtures if n in ("omt/5_16_21", "omt/2_2_2", "omt/14_8298_10748", "amazon/5_5_11", "amazon/8_136_89",
                                                  "bing/4-8-5", "bing/5-16-11", "omt/9_265_341", "omt/12_2132_2734")]
    assert len(pick) >= 6
    max_cols = 0
    for k, (name, data) in enumerate(pick):
        flags = abi.FLAG_CLOSE_RINGS | _flags(abi, name)
        blob2, offs2 = util.concat_tiles([data])
        ref2 = oracle.decode_batch(blob2, offs2, abi.CONTAINER_GEN2B, flags)
        for optimized in (False, True):
            for id_last in (False, True):
                w, nf = util.rewrap_gen3(abi, oracle, data, optimized, props=True, gen=gen, id_last=id_last)
                blob3, offs3 = util.concat_tiles([w])
                rc, layers, end_pos = oracle.parse_tile(blob3, abi.CONTAINER_GEN3, flags=flags, n_fields=nf)
                assert rc == 0 and end_pos == len(w), (name, optimized, id_last, rc, end_pos, len(w))
                max_cols = max(max_cols, int(layers["num_columns"].max()))
                ref3 = oracle.decode_batch(blob3, offs3, abi.CONTAINER_GEN3, flags, n_fields=nf)
                assert np.array_equal(ref3.tile_status, ref2.tile_status)
                util.compare_results(abi, ref3, ref2, same_container=False)
                if id_last and ref3.layers["has_id"].any():
                    L = ref3.layers[ref3.layers["has_id"] != 0][0]
                    assert L["streams"][abi.SLOT_ID]["byte_offset"] > L["streams"][abi.SLOT_VBUF]["byte_offset"]
    assert max_cols > 10  # more property columns than the 8 the round-1 device walker could hop over
