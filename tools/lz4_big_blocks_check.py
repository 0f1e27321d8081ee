import sys, importlib, numpy as np, torch
sys.path.insert(0,'.')
from tools import gen
rcz = importlib.import_module("rust-compress_b200")
ctx = rcz.Context(device=0); ctx.set_stream(torch.cuda.current_stream())
for kind, unit in (("lzsyn", 96<<20), ("hextext", 48<<20), ("runs", 128<<20), ("random", 40<<20)):
    raw = gen.units(kind, 12345, unit, 2)
    packed, off, lens = gen.lz4_compress_units(raw, unit, 2)
    d_in = torch.from_numpy(packed).cuda(); d_out = torch.zeros(unit*2, dtype=torch.uint8, device="cuda")
    oo = np.arange(2, dtype=np.uint64)*unit
    ol, st = ctx.lz4_decode_blocks(d_in, off, lens, d_out, oo, np.full(2, unit, dtype=np.uint64))
    ok = torch.equal(d_out, torch.from_numpy(raw).cuda())
    print(kind, unit, "windows", int((lens[0]+8191)//8192), "status", st.tolist(), "ok", ok, "ms", ctx.last_stage_ms())
    assert ok and (st == 0).all()
print("big blocks ok")
