"""Builds the committed golden data from the reference checkout (run in the build container only).

  fixtures_covt.tar.xz      the reference's own gen-2b fixture tiles test/fixtures/{omt,bing,amazon}/covt/*.covt
                            (data files, not sources; the GPU box has no /root/reference)
  mvt_geometry_digests.json per (source, tile, layer): feature / ring / vertex counts + a blake2b digest of the
                            canonical geometry read from the partner .mvt/.pbf (tests/mvt.py, tests/canon.py)
  mvt_property_digests.json per (source, tile, layer, property key): features carrying the key + a digest of the value column

Usage: python tests/golden/make_golden.py [/root/reference]
"""
import glob
import io
import json
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import canon  # noqa: E402
import mvt  # noqa: E402


def main(ref):
    fx = os.path.join(ref, "test", "fixtures")
    members = []
    for src in ("omt", "bing", "amazon"):
        for f in sorted(glob.glob(os.path.join(fx, src, "covt", "*.covt"))):
            if src == "omt" and os.path.basename(f) == "3_4_5.covt":
                continue  # older gen-2a container
            members.append((src + "/" + os.path.basename(f), f))
    out = os.path.join(HERE, "fixtures_covt.tar.xz")
    with tarfile.open(out, "w:xz", preset=9) as tf:
        for name, path in members:
            data = open(path, "rb").read()
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            ti.mtime = 0
            tf.addfile(ti, io.BytesIO(data))
    print("wrote", out, os.path.getsize(out), "bytes,", len(members), "tiles")

    digests = {}
    for src, ext in (("omt", "mvt"), ("amazon", "pbf")):
        for f in sorted(glob.glob(os.path.join(fx, src, "mvt", "*." + ext))):
            tile = os.path.basename(f)[: -len(ext) - 1]
            if not os.path.exists(os.path.join(fx, src, "covt", tile + ".covt")):
                continue
            for layer in mvt.read_layers(open(f, "rb").read()):
                c = mvt.canonical_from_features(layer["features"])
                digests["%s/%s/%s" % (src, tile, layer["name"])] = {
                    "features": int(len(c[0])), "rings": int(len(c[1])), "vertices": int(len(c[2]) // 2),
                    "digest": canon.digest(c)}
    out = os.path.join(HERE, "mvt_geometry_digests.json")
    with open(out, "w") as fh:
        json.dump(digests, fh, indent=0, sort_keys=True)
    print("wrote", out, len(digests), "layers")

    # property columns: per (source, tile, layer, key) the number of features that carry the key and a digest of
    # the value column in feature order (tests/canon.property_digest); pins oracle/properties.py
    pdig = {}
    for src, ext in (("omt", "mvt"), ("amazon", "pbf")):
        for f in sorted(glob.glob(os.path.join(fx, src, "mvt", "*." + ext))):
            tile = os.path.basename(f)[: -len(ext) - 1]
            if not os.path.exists(os.path.join(fx, src, "covt", tile + ".covt")) or (src == "omt" and tile == "3_4_5"):
                continue
            for layer in mvt.read_layers(open(f, "rb").read(), with_properties=True):
                # the converter folds the two spellings of a localized key (`name:de`, `name_de`) into ONE sub-column; wherever
                # both exist on a feature of the fixtures their values agree, so the golden column is their union under the
                # `_` spelling
                keys = sorted(set(k.replace(":", "_") for pr in layer["properties"] for k in pr))
                for k in keys:
                    col = []
                    for pr in layer["properties"]:
                        vals = [v[1] for kk, v in pr.items() if kk.replace(":", "_") == k]
                        assert all(v == vals[0] for v in vals), (f, layer["name"], k, vals)
                        col.append(vals[0] if vals else None)
                    pdig["%s/%s/%s/%s" % (src, tile, layer["name"], k)] = {
                        "present": sum(v is not None for v in col), "digest": canon.property_digest(col)}
    out = os.path.join(HERE, "mvt_property_digests.json")
    with open(out, "w") as fh:
        json.dump(pdig, fh, indent=0, sort_keys=True)
    print("wrote", out, len(pdig), "property columns")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
