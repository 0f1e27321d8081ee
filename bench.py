#!/usr/bin/env python
"""bench.py — throughput of the hot path on B200 (BASELINE.json metric: uncompressed GB/s, % of HBM roofline, CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload lz4|bwt_decode]

Default workload = BASELINE.json configs[1]: lz4::Decoder over 256 independent 4 MiB synthetic blocks on one B200
("lzsyn" generator of SURVEY.md §8d, compressed with liblz4).  One step = one pass of the batch through the C ABI.
  value     device-resident: compressed blocks already in HBM, output stays in HBM (CUDA events, max over ranks)
  e2e       the same call with HOST (pinned) buffers: H2D of the compressed blocks + kernel + D2H of the output
  roofline  algorithmic bytes (compressed in + decoded out) / average step duration, against MEASURED_PEAKS.json
With N > 1 (torchrun, one rank per GPU) every rank decodes its own 256 blocks (weak scaling, no data-path collective).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = 4 << 20
COUNT = 256


def traffic_for(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json), else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU with nvidia-smi while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- workloads
def make_lz4(rank, nthreads):
    from tools import gen
    raw = gen.units("lzsyn", gen.unit_seed(2, rank * COUNT), UNIT, COUNT, nthreads=nthreads)
    packed, off, lens = gen.lz4_compress_units(raw, UNIT, COUNT, nthreads=nthreads)
    out_off = np.arange(COUNT, dtype=np.uint64) * UNIT
    caps = np.full(COUNT, UNIT, dtype=np.uint64)
    return {"raw": raw, "packed": packed, "in_off": off, "in_len": lens, "out_off": out_off, "out_cap": caps,
            "U": UNIT * COUNT, "C": int(lens.sum()),
            "name": "lz4::Decoder over 256 independent 4 MiB synthetic blocks (lzsyn generator, liblz4 LZ4_compress_default)"}


def make_bwt_decode(rank, nthreads, count=COUNT):
    """BASELINE configs[2] decode leg; the L columns come from the device encoder when available, else the oracle (slow)."""
    from oracle import oracle
    from tools import gen
    raw = gen.units("random", gen.unit_seed(3, rank * count), UNIT, count, nthreads=nthreads)
    off = np.arange(count, dtype=np.uint64) * UNIT
    n = np.full(count, UNIT, dtype=np.uint64)
    l_buf = np.zeros(UNIT * count + 64, dtype=np.uint8)
    origin, st = oracle.bwt_encode_blocks_mt(raw, off, n, l_buf, nthreads)
    assert (st == 0).all()
    return {"raw": raw, "L": l_buf, "off": off, "n": n, "origin": origin, "U": UNIT * count, "C": UNIT * count + 8 * count,
            "name": "bwt::Decoder over %d x 4 MiB random blocks" % count}


def config_for(w):
    """The same `config` object on both arms (ours / --impl reference)."""
    nb = len(w.get("in_off", w.get("off")))
    return {"workload": w["name"], "blocks_per_gpu": nb, "block_bytes": UNIT, "compressed_bytes_per_gpu": w["C"],
            "l2": "inputs larger than L2: %.2f GiB touched per step vs 126 MB L2, no flush needed" % ((w["C"] + w["U"]) / 2**30),
            "parallelism": "independent blocks, %d per GPU, no data-path collective" % nb}


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path = oracle/ (C++ restatement; no rustc in this image), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    w = make_lz4(0, cores)
    out = np.zeros(w["U"] + 64, dtype=np.uint8)
    for _ in range(args.warmup):
        oracle.lz4_decode_blocks_mt(w["packed"], w["in_off"], w["in_len"], out, w["out_off"], w["out_cap"], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ln, st = oracle.lz4_decode_blocks_mt(w["packed"], w["in_off"], w["in_len"], out, w["out_off"], w["out_cap"], cores)
    dt = (time.perf_counter() - t0) / args.steps
    assert (st == 0).all() and bytes(out[: w["U"]]) == w["raw"].tobytes()
    gbs = w["U"] / dt / 1e9
    line = {"impl": "reference", "metric": "lz4_decode_uncompressed_GBps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config_for(w),
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "all 256 blocks per step, C++ restatement of lz4.rs BlockDecoder::decode, one block per thread"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rcz = importlib.import_module("rust-compress_b200")
    ctx = rcz.Context(device=local)
    ctx.set_stream(torch.cuda.current_stream())
    nthreads = max(1, (os.cpu_count() or 1) // world)
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "lz4":
        w = make_lz4(rank, nthreads)
        d_in = torch.from_numpy(w["packed"]).cuda()
        d_out = torch.zeros(w["U"], dtype=torch.uint8, device="cuda")
        d_len = None

        def step_dev():
            return ctx.lz4_decode_blocks(d_in, w["in_off"], w["in_len"], d_out, w["out_off"], w["out_cap"], async_=True)

        h_in = torch.from_numpy(w["packed"]).pin_memory()
        h_out = torch.zeros(w["U"] + 64, dtype=torch.uint8).pin_memory()
        h_in_np, h_out_np = h_in.numpy(), h_out.numpy()

        def step_host():
            return ctx.lz4_decode_blocks(h_in_np, w["in_off"], w["in_len"], h_out_np, w["out_off"], w["out_cap"])

        h2d, d2h = int(w["in_off"][-1] + w["in_len"][-1] - w["in_off"][0]), w["U"]
        metric = "lz4_decode_uncompressed_GBps"
        kernel_name = "lz4_mat_kernel"
    elif args.workload == "bwt_decode":
        w = make_bwt_decode(rank, nthreads, count=args.blocks or 64)
        d_in = torch.from_numpy(w["L"]).cuda()
        d_out = torch.zeros(w["U"], dtype=torch.uint8, device="cuda")

        def step_dev():
            return ctx.bwt_decode_blocks(d_in, w["off"], w["n"], w["origin"], d_out, w["off"], async_=True)

        h_in = torch.from_numpy(w["L"]).pin_memory()
        h_out = torch.zeros(w["U"] + 64, dtype=torch.uint8).pin_memory()
        h_in_np, h_out_np = h_in.numpy(), h_out.numpy()

        def step_host():
            return ctx.bwt_decode_blocks(h_in_np, w["off"], w["n"], w["origin"], h_out_np, w["off"])

        h2d, d2h = w["U"], w["U"]
        metric = "bwt_decode_uncompressed_GBps"
        kernel_name = "ibwt_walk_kernel"
    else:
        raise SystemExit("unknown workload " + args.workload)

    # ---- device-resident timing (value, roofline)
    for _ in range(max(args.warmup, 3)):
        res = step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in evs:
        a.record()
        res = step_dev()
        b.record()
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms_per_step = max_over_ranks(total_ms / args.steps)
    # duration of the kernel(s) alone: CUDA events recorded by librcz on the context's stream around its launches
    kms, stages = [], []
    for _ in range(3):
        step_dev()
        kms.append(ctx.last_kernel_ms())
        stages.append(ctx.last_stage_ms())
    kern_ms = float(np.mean(kms))
    # ops that launch several kernels per call (lz4: parse, scan, materialise) report every kernel; the roofline is that of the
    # dominant one, and `op_frac` is the same ratio for the whole call
    stage_ms = [float(np.mean([st[i] for st in stages])) for i in range(len(stages[0]))] if stages and stages[0] else []
    step_ms_mean = float(np.mean(step_ms))
    # final gather of the decoded shards (N > 1): NCCL all-gather over NVLink, timed separately from the decode
    gather = None
    if world > 1:
        flat = torch.empty(w["U"] * world, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            dist.all_gather_into_tensor(flat, d_out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            dist.all_gather_into_tensor(flat, d_out)
        g1.record()
        torch.cuda.synchronize()
        gms = max_over_ranks(g0.elapsed_time(g1) / 3)
        gather = {"collective": "ncclAllGather (uniform shards)", "bytes_per_rank": w["U"], "ms": gms,
                  "busbw_GBps": w["U"] * (world - 1) / (gms * 1e-3) / 1e9}
        del flat
    out_len, status = res[0], res[1]
    assert int((status != 0).sum().item() if hasattr(status, "sum") else 0) == 0, "decode reported errors"
    ok = torch.equal(d_out, torch.from_numpy(w["raw"]).cuda())
    assert ok, "device output differs from the generator's bytes"

    # ---- end-to-end timing through the host-buffer ABI (e2e)
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol, st = step_host()[:2]
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    assert (st == 0).all() and bytes(h_out_np[: w["U"]]) == w["raw"].tobytes()

    # ---- CPU baseline on rank 0 (N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and args.workload == "lz4":
        from oracle import oracle
        nb = 64
        out = np.zeros(w["U"] + 64, dtype=np.uint8)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 12.0 and reps < 12:
            oracle.lz4_decode_blocks_mt(w["packed"], w["in_off"][:nb], w["in_len"][:nb], out, w["out_off"][:nb], w["out_cap"][:nb], 1)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": nb * UNIT / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": "first %d of the 256 blocks, %d repetitions, single thread (the reference is single-threaded); "
                         "C++ restatement of lz4.rs BlockDecoder::decode (no rustc in this image)" % (nb, reps)}

    if rank == 0:
        value = world * w["U"] / (ms_per_step * 1e-3) / 1e9
        alg_bytes = w["C"] + w["U"]
        dom_ms = max(stage_ms) if stage_ms else kern_ms
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        line = {"metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_for(w),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic_for(kernel_name), "kernel": kernel_name, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms, "op_ms": kern_ms,
                             "op_frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / peak, "stage_ms": stage_ms, "step_ms_events": step_ms_mean},
                "cpu_baseline": cpu,
                "e2e": {"value": world * w["U"] / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
                "gpu_launches": launches, "clocks": clocks, "host_cores": os.cpu_count()}
        if gather:
            line["gather"] = gather
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lz4")
    ap.add_argument("--blocks", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
