"""Regenerates tests/golden/ (run in the authoring container, where /root/reference exists).

1. Copies the reference's own small fixtures (src/data/test.txt, test.lz4.1-9, test.z.0-9, test.z.go — the
   known-answer inputs of lz4.rs:647-659 and flate.rs:528-542) under a `ref_` prefix.
2. Records sha256 of the big fixture pair (test.large / test.large.z.5, flate.rs:544-548) in manifest.json; the
   files themselves (6.1 MB / 3.4 MB) are read in place when present.
3. Records oracle-derived vectors for the codecs whose encode side the reference does not pin (bwt, dc, ari):
   they agree with the independently derived SURVEY.md Appendix C values.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

SRC = "/root/reference/src/data"
names = ["test.txt", "test.z.go"] + ["test.lz4.%d" % i for i in range(1, 10)] + ["test.z.%d" % i for i in range(10)]
man = {"files": {}, "large": {}, "vectors": {}}
for n in names:
    dst = os.path.join(HERE, "ref_" + n)
    shutil.copyfile(os.path.join(SRC, n), dst)
    man["files"]["ref_" + n] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
for n in ["test.large", "test.large.z.5"]:
    b = open(os.path.join(SRC, n), "rb").read()
    man["large"][n] = {"sha256": hashlib.sha256(b).hexdigest(), "size": len(b)}
txt = open(os.path.join(SRC, "test.txt"), "rb").read()
v = man["vectors"]
st, l, org = o.bwt_encode(b"abracadabra")
v["bwt_abracadabra"] = {"L": l.decode(), "origin": org, "sa": [int(x) for x in o.bwt_suffixes(b"abracadabra")],
                        "table": [int(x) for x in o.bwt_inversion_table(l, org)[1]]}
st, s = o.bwt_stream_encode(b"abracadabra", 1024)
v["bwt_stream_abracadabra_hex"] = s.hex()
st, s = o.bwt_stream_encode(txt, 1024)
v["bwt_stream_txt"] = {"len": len(s), "sha256": hashlib.sha256(s).hexdigest()}
v["ari_abracadabra_hex"] = o.ari_encode(b"abracadabra").hex()
v["ari_empty_hex"] = o.ari_encode(b"").hex()
e = o.ari_encode(txt)
v["ari_txt"] = {"len": len(e), "sha256": hashlib.sha256(e).hexdigest()}
st, init, dist, ctx = o.dc_encode(b"teeesst_dc", True)
v["dc_teeesst_dc"] = {"dist": [int(x) for x in dist], "init": {chr(i): int(x) for i, x in enumerate(init) if x < 10},
                      "ctx": [[int(a), int(b), int(c)] for a, b, c in zip(*ctx)]}
st, init, dist = o.dc_encode(b"abracadabra")
v["dc_abracadabra_dist"] = [int(x) for x in dist]
enc = o.lz4_encode_block(txt)
v["lz4_encode_block_txt"] = {"len": len(enc), "sha256": hashlib.sha256(enc).hexdigest()}
v["rle_txt_len"] = len(o.rle_encode(txt))
v["adler32_txt"] = "%08x" % o.adler32(txt)
json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(names), "fixtures + manifest")
