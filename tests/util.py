"""Shared helpers for the parity tests: pack units into one buffer, call the batch ABI, unpack results."""
import numpy as np


def pack(units, pad_front=0, gap=0, align=1):
    """Concatenate byte strings; returns (uint8 array with 64 B slack, offsets u64, lengths u64)."""
    offs, buf = [], bytearray(b"\xAA" * pad_front)
    for u in units:
        while len(buf) % align:
            buf += b"\x55"
        offs.append(len(buf))
        buf += u
        buf += b"\x55" * gap
    arr = np.frombuffer(bytes(buf) + b"\0" * 64, dtype=np.uint8).copy()
    return arr, np.array(offs, dtype=np.uint64), np.array([len(u) for u in units], dtype=np.uint64)


def out_layout(caps, gap=0):
    caps = [int(c) for c in caps]
    off = np.zeros(len(caps), dtype=np.uint64)
    if len(caps) > 1:
        off[1:] = np.cumsum([c + gap for c in caps[:-1]])
    total = int(sum(caps)) + gap * len(caps) + 64
    return off, np.array(caps, dtype=np.uint64), total


def to_dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def run_batch(ctx, method, units, caps, device=False, **kw):
    """Runs ctx.<method>(in, in_off, in_len, out, out_off, out_cap) and returns [(status, bytes)] (+ extras)."""
    inb, in_off, in_len = pack(units, **kw)
    out_off, out_cap, total = out_layout(caps, gap=3)
    outb = np.zeros(total, dtype=np.uint8)
    if device:
        import torch
        d_in, d_out = to_dev(inb), torch.zeros(total, dtype=torch.uint8, device="cuda")
        res = getattr(ctx, method)(d_in, in_off, in_len, d_out, out_off, out_cap)
        outb = d_out.cpu().numpy()
    else:
        res = getattr(ctx, method)(inb, in_off, in_len, outb, out_off, out_cap)
    out_len, status = res[0], res[1]
    outs = [(int(s), outb[int(o): int(o) + int(l)].tobytes()) for s, o, l in zip(status, out_off, out_len)]
    return outs, res[2:]
