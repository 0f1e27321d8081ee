#!/usr/bin/env python
"""bench.py — throughput of the hot path on B200 (BASELINE.json metric: uncompressed GB/s per codec, % of HBM roofline, CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--codecs lz4,bwt,flate,pipeline]

Headline keys = BASELINE.json configs[1]: lz4::Decoder over 256 independent 4 MiB synthetic blocks on one B200 ("lzsyn" generator of
SURVEY.md §8d, compressed with liblz4).  One step = one pass of the batch through the C ABI.
  value     device-resident: compressed blocks already in HBM, output stays in HBM (CUDA events, max over ranks)
  e2e       the same call with HOST (pinned) buffers: H2D of the compressed blocks + kernels + D2H of the output
  roofline  algorithmic bytes (compressed in + decoded out) of the dominant kernel / its duration, against MEASURED_PEAKS.json
With N > 1 (torchrun, one rank per GPU) every rank decodes its own 256 blocks (weak scaling, no data-path collective);
`value_with_gather` adds the final gather of the decoded shards over NVLink.

`per_codec` carries the other BASELINE configs, measured in the same run, each with value / roofline / cpu_baseline / e2e:
  bwt_encode, bwt_decode      configs[2]  1 GiB random bytes in 256 blocks of 4 MiB, sharded over the ranks (strong scaling)
  flate_decode                configs[3]  131,072 raw-DEFLATE streams of 64 KiB (8 GiB), sharded over the ranks (strong scaling)
  bwt_dc_ari_encode/_decode   configs[4]  256 text blocks of 4 MiB per GPU through the chained pipeline (weak scaling)
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on the same workloads (bounded samples).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import traceback
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = 4 << 20
COUNT = 256
FL_UNIT = 65536
FL_TOTAL = 131072          # configs[3]: 8 GiB of 64 KiB streams
FL_UNIQUE = 16384          # generated and compressed once (1 GiB), tiled to FL_TOTAL
ARI_CHUNK = int(os.environ.get("RCZ_BENCH_ARI_CHUNK", 16384))   # configs[4]: ByteEncoder streams of 16 KiB of serialised dc output (65,536 streams per GiB)


def traffic_for(kernel):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/traffic.json), else None."""
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


def bind_numa(local):
    """Run this rank (and therefore first-touch its pinned buffers) on the CPUs of its GPU's NUMA node.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return "numa node unknown (single node)"
        cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # pragma: no cover
        return "unavailable: %s" % type(e).__name__
    return "unchanged"


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


def config_for(w):
    """The same `config` object on both arms (ours / --impl reference)."""
    nb = len(w.get("in_off", w.get("off")))
    return {"workload": w["name"], "blocks_per_gpu": nb, "block_bytes": UNIT, "compressed_bytes_per_gpu": w["C"],
            "l2": "inputs larger than L2: %.2f GiB touched per step vs 126 MB L2, no flush needed" % ((w["C"] + w["U"]) / 2**30),
            "parallelism": "independent blocks, %d per GPU, no data-path collective" % nb}


def make_flate_unique(nthreads, count=FL_UNIQUE):
    """`count` hexdump-text streams of 64 KiB, each compressed on its own with zlib level 6 as raw DEFLATE."""
    from concurrent.futures import ThreadPoolExecutor
    from tools import gen
    raw = gen.units("hextext", gen.unit_seed(4, 0), FL_UNIT, count, nthreads=nthreads)

    def comp(i):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        return c.compress(raw[i * FL_UNIT: (i + 1) * FL_UNIT].tobytes()) + c.flush()
    with ThreadPoolExecutor(max(1, nthreads)) as ex:
        cs = list(ex.map(comp, range(count), chunksize=64))
    lens = np.array([len(c) for c in cs], dtype=np.uint64)
    al = (lens + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(count, dtype=np.uint64)
    off[1:] = np.cumsum(al)[:-1]
    span = int(al.sum())
    packed = np.zeros(span, dtype=np.uint8)
    for i, c in enumerate(cs):
        packed[int(off[i]): int(off[i]) + len(c)] = np.frombuffer(c, dtype=np.uint8)
    return raw, packed, off, lens, span


# ------------------------------------------------------------------------------------------------- reference arm
def _timed_cpu(fn, budget_s=8.0, max_reps=6):
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        fn()
        reps += 1
    return (time.perf_counter() - t0) / reps, reps


def cpu_legs(cores, which, single):
    """CPU restatement (oracle/) of every per-codec workload on a bounded sample; `cores` threads (1 = the reference as written)."""
    from oracle import oracle
    from tools import gen
    out = {}
    kind = {"kind": "port", "cores": cores, "unit": "GB/s"}
    nb = max(1, cores) * (1 if single else 2)
    if "bwt" in which:
        nb_b = min(COUNT, max(2, nb))
        raw = gen.units("random", gen.unit_seed(3, 0), UNIT, nb_b)
        off = np.arange(nb_b, dtype=np.uint64) * UNIT
        n = np.full(nb_b, UNIT, dtype=np.uint64)
        l_buf = np.zeros(UNIT * nb_b + 64, dtype=np.uint8)
        res = {}

        def enc():
            res["o"] = oracle.bwt_encode_blocks_mt(raw, off, n, l_buf, cores)
        dt, reps = _timed_cpu(enc, 6.0, 3)
        origin, st = res["o"]
        assert (st == 0).all()
        out["bwt_encode"] = dict(kind, value=nb_b * UNIT / dt / 1e9, sample="%d of the 256 random 4 MiB blocks, %d repetitions; oracle compute_suffixes + TransformIterator" % (nb_b, reps))
        back = np.zeros(UNIT * nb_b + 64, dtype=np.uint8)
        dt, reps = _timed_cpu(lambda: oracle.bwt_decode_blocks_mt(l_buf, off, n, origin, back, cores), 4.0, 4)
        assert bytes(back[: UNIT * nb_b]) == raw.tobytes()
        out["bwt_decode"] = dict(kind, value=nb_b * UNIT / dt / 1e9, sample="%d of the 256 blocks, %d repetitions; oracle compute_inversion_table + InverseIterator" % (nb_b, reps))
    if "flate" in which:
        ns = 256 * max(1, cores)
        raw, packed, off, lens, span = make_flate_unique(max(cores, 4), ns)
        outb = np.zeros(FL_UNIT * ns + 64, dtype=np.uint8)
        ooff = np.arange(ns, dtype=np.uint64) * FL_UNIT
        caps = np.full(ns, FL_UNIT, dtype=np.uint64)
        res = {}

        def dec():
            res["o"] = oracle.flate_decode_streams_mt(packed, off, lens, outb, ooff, caps, cores)
        dt, reps = _timed_cpu(dec, 4.0, 6)
        assert (res["o"][1] == 0).all() and bytes(outb[: FL_UNIT * ns]) == raw.tobytes()
        out["flate_decode"] = dict(kind, value=ns * FL_UNIT / dt / 1e9, sample="%d of the 131,072 streams, %d repetitions; oracle flate::Decoder" % (ns, reps))
    if "pipeline" in which:
        nb_p = min(COUNT, max(1, cores) if single else max(2, cores))
        raw = gen.units("hextext", gen.unit_seed(5, 0), UNIT, nb_p)
        off = np.arange(nb_p, dtype=np.uint64) * UNIT
        n = np.full(nb_p, UNIT, dtype=np.uint64)
        cap = 24 + 4 * 260 + 2 * 4 * (256 + UNIT) + 64 * 260
        cont = np.zeros(cap * nb_p + 64, dtype=np.uint8)
        coff = np.arange(nb_p, dtype=np.uint64) * cap
        res = {}

        def enc():
            res["o"] = oracle.bda_encode_blocks_mt(raw, off, n, ARI_CHUNK, cont, coff, np.full(nb_p, cap, np.uint64), cores)
        dt, reps = _timed_cpu(enc, 4.0, 2)
        clen, org, st = res["o"]
        assert (st == 0).all()
        out["bwt_dc_ari_encode"] = dict(kind, value=nb_p * UNIT / dt / 1e9, sample="%d text blocks of 4 MiB, %d repetitions; oracle bwt + dc + ByteEncoder (%d KiB streams)" % (nb_p, reps, ARI_CHUNK >> 10))
        back = np.zeros(UNIT * nb_p + 64, dtype=np.uint8)
        dt, reps = _timed_cpu(lambda: oracle.bda_decode_blocks_mt(cont, coff, clen, ARI_CHUNK, back, off, n, cores), 4.0, 3)
        assert bytes(back[: UNIT * nb_p]) == raw.tobytes()
        out["bwt_dc_ari_decode"] = dict(kind, value=nb_p * UNIT / dt / 1e9, sample="%d text blocks of 4 MiB, %d repetitions; oracle ByteDecoder + dc + inverse bwt" % (nb_p, reps))
    return out


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
    which = [c for c in args.codecs.split(",") if c != "lz4"]
    try:
        line["per_codec"] = cpu_legs(cores, which, single=False)
    except Exception as e:  # pragma: no cover
        line["per_codec"] = {"error": repr(e)}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.numa = bind_numa(self.local) if self.world > 1 else "single rank: not bound"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.rcz = importlib.import_module("rust-compress_b200")
        self.shard = importlib.import_module("rust-compress_b200.shard")
        self.ctx = self.rcz.Context(device=self.local)
        self.ctx.set_stream(torch.cuda.current_stream())
        self.nthreads = max(1, len(os.sched_getaffinity(0)) // (1 if self.numa.startswith("node") else self.world))
        self.peak, self.peak_src = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def time_dev(self, step, steps, warmup):
        """W warm-ups, then exactly `steps` steps between barrier + synchronize on both sides; CUDA events; max over ranks (ms per step)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        self.barrier()
        return self.max_over_ranks(ms)

    def time_host(self, step, steps, warmup=1):
        for _ in range(warmup):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        return self.max_over_ranks((time.perf_counter() - t0) / steps)

    def reps_for(self, est_ms, steps, budget_ms=2500.0):
        return int(max(2, min(steps, budget_ms / max(est_ms, 1e-3))))

    def pinned(self, arr):
        t = self.torch.from_numpy(arr).pin_memory()
        return t, t.numpy()

    def gather_time(self, decode_step, shard_tensor, steps):
        """decode + final gather of the uniform shards (ncclAllGather over NVLink) in one timed region."""
        torch, dist = self.torch, self.dist
        flat = torch.empty(shard_tensor.numel() * self.world, dtype=torch.uint8, device="cuda")

        def step():
            decode_step()
            dist.all_gather_into_tensor(flat, shard_tensor)
        ms = self.time_dev(step, steps, 2)
        del flat
        return ms


def roofline(env, kernel, alg_bytes, kernel_ms, op_ms, extra=None, traffic_scale=None):
    """traffic: DRAM bytes of the kernel from the committed ncu capture; the captures of the per_codec kernels were taken on a fraction of
    the bench workload (profiles/traffic.json _note), so they are scaled by units(bench) / units(capture) and flagged as such."""
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    tr = traffic_for(kernel)
    if tr is not None and traffic_scale is not None:
        tr = tr * traffic_scale
    r = {"bound": "hbm", "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak, "traffic": tr,
         "kernel": kernel, "peak_source": env.peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms, "op_ms": op_ms,
         "op_frac": alg_bytes / (op_ms * 1e-3) / 1e9 / env.peak}
    r.update(extra or {})
    return r


def launches_of(ctx, step):
    """librcz kernel launches of one call"""
    l0 = ctx.launches
    step()
    return ctx.launches - l0


def stage_avg(ctx, step, reps=3):
    kms, stages = [], []
    for _ in range(reps):
        step()
        kms.append(ctx.last_kernel_ms())
        stages.append(ctx.last_stage_ms())
    n = min(len(s) for s in stages) if stages else 0
    return float(np.mean(kms)), [float(np.mean([s[i] for s in stages])) for i in range(n)]


def leg_lz4(env, args):
    torch, dist, ctx, world, rank = env.torch, env.dist, env.ctx, env.world, env.rank
    w = make_lz4(rank, env.nthreads)
    d_in = torch.from_numpy(w["packed"]).cuda()
    d_out = torch.zeros(w["U"], dtype=torch.uint8, device="cuda")
    res = {}

    def step_dev():
        res["o"] = ctx.lz4_decode_blocks(d_in, w["in_off"], w["in_len"], d_out, w["out_off"], w["out_cap"], async_=True)

    # ---- device-resident timing (value, roofline)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    env.barrier()
    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in evs:
        a.record()
        step_dev()
        b.record()
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    env.barrier()
    ms_per_step = env.max_over_ranks(total_ms / args.steps)
    # duration of the kernels alone: CUDA events recorded by librcz on the context's stream around its launches (parse, scan, materialise)
    kern_ms, stage_ms = stage_avg(ctx, step_dev)
    out_len, status = res["o"]
    assert int((status != 0).sum().item()) == 0, "decode reported errors"
    assert torch.equal(d_out, torch.from_numpy(w["raw"]).cuda()), "device output differs from the generator's bytes"

    # ---- decode + final gather (N > 1): (a) ncclAllGather after the decode; (b) fused: the materialise kernel stores every output chunk
    # into the peers' gathered buffers as well (symmetric memory over NVLink / NVSwitch), no separate collective
    gather = None
    if world > 1:
        flat = torch.empty(w["U"] * world, dtype=torch.uint8, device="cuda")
        gms = env.time_dev(lambda: dist.all_gather_into_tensor(flat, d_out), 3, 2)
        both = env.gather_time(step_dev, d_out, max(3, min(args.steps, 10)))
        gather = {"collective": "ncclAllGather of the uniform 1 GiB shards (torch.distributed, NCCL over NVLink), issued after the decode", "bytes_per_rank": w["U"],
                  "gather_alone_ms": gms, "busbw_GBps": w["U"] * (world - 1) / (gms * 1e-3) / 1e9, "decode_plus_gather_ms": both,
                  "value_with_gather": world * w["U"] / (both * 1e-3) / 1e9}
        try:
            import torch.distributed._symmetric_memory as symm
            U = w["U"]
            buf = symm.empty(world * U, dtype=torch.uint8, device="cuda")
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            mine = buf[rank * U: (rank + 1) * U]
            peers = [ptrs[p] + rank * U for p in range(world) if p != rank]
            buf.zero_()
            env.barrier()

            def step_fused():
                res["f"] = ctx.lz4_decode_blocks_gather(d_in, w["in_off"], w["in_len"], mine, w["out_off"], w["out_cap"], peers)
            fms = env.time_dev(step_fused, max(3, min(args.steps, 10)), 2)
            env.barrier()
            assert int((res["f"][1] != 0).sum().item()) == 0
            # every rank's shard must have arrived: the gathered buffer equals what ncclAllGather delivered above, byte for byte
            assert torch.equal(buf, flat) and torch.equal(mine, d_out), "fused gather: a peer's shard did not arrive intact"
            gather["fused"] = {"what": "rcz_lz4_decode_blocks_gather: lz4_mat_kernel writes its output to the local buffer and pushes every finished tile from its "
                                       "shared-memory ring to the same offset of the %d peers' gathered buffers with TMA bulk stores (torch symmetric memory, "
                                       "P2P over NVLink / NVSwitch); no collective call" % (world - 1),
                               "decode_plus_gather_ms": fms, "value_with_gather": world * U / (fms * 1e-3) / 1e9}
            gather["nccl_value_with_gather"] = gather["value_with_gather"]
            gather["value_with_gather"] = max(gather["value_with_gather"], gather["fused"]["value_with_gather"])
            del buf
            del flat
        except Exception as e:  # symmetric memory unavailable on this box / torch build: the NCCL figure stands
            traceback.print_exc(file=sys.stderr)
            gather["fused"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---- end-to-end timing through the host-buffer ABI (e2e), and the platform's copy ceiling for the same bytes
    h_in, h_in_np = env.pinned(w["packed"])
    h_out = torch.zeros(w["U"] + 64, dtype=torch.uint8).pin_memory()
    h_out_np = h_out.numpy()
    r2 = {}

    def step_host():
        r2["o"] = ctx.lz4_decode_blocks(h_in_np, w["in_off"], w["in_len"], h_out_np, w["out_off"], w["out_cap"])
    e2e_s = env.time_host(step_host, args.steps, 2)
    assert (r2["o"][1] == 0).all() and bytes(h_out_np[: w["U"]]) == w["raw"].tobytes()
    h2d, d2h = int(w["in_off"][-1] + w["in_len"][-1] - w["in_off"][0]), w["U"]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def step_copy():
        with torch.cuda.stream(s1):
            d_in[:h2d].copy_(h_in[:h2d], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[: w["U"]].copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    ceil_s = env.time_host(step_copy, 5, 2)

    # ---- CPU baseline on rank 0 (N == 1 only)
    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle
        nb = 64
        out = np.zeros(w["U"] + 64, dtype=np.uint8)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0 and reps < 12:
            oracle.lz4_decode_blocks_mt(w["packed"], w["in_off"][:nb], w["in_len"][:nb], out, w["out_off"][:nb], w["out_cap"][:nb], 1)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": nb * UNIT / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": "first %d of the 256 blocks, %d repetitions, single thread (the reference is single-threaded); "
                         "C++ restatement of lz4.rs BlockDecoder::decode (no rustc in this image)" % (nb, reps)}

    alg_bytes = w["C"] + w["U"]
    dom_ms = max(stage_ms) if stage_ms else kern_ms
    line = {"metric": "lz4_decode_uncompressed_GBps", "value": world * w["U"] / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_for(w),
            "roofline": roofline(env, "lz4_mat_kernel", alg_bytes, dom_ms, kern_ms, {"stage_ms": stage_ms, "stages": ["lz4_parse_kernel", "lz4_scan_kernel", "lz4_mat_kernel"],
                                                                                    "step_ms_events": float(np.mean(step_ms))}),
            "cpu_baseline": cpu,
            "e2e": {"value": world * w["U"] / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "ceiling": {"what": "plain pinned cudaMemcpyAsync of the same bytes, H2D and D2H on two streams at once, no kernels; max over ranks",
                                "ms": ceil_s * 1e3, "GBps": world * w["U"] / ceil_s / 1e9}},
            "gpu_launches": launches, "clocks": clocks, "host_cores": os.cpu_count()}
    if gather:
        line["gather"] = gather
        line["value_with_gather"] = gather["value_with_gather"]
    return line


def _shard_range(env, total):
    lo, hi = env.shard.my_range(np.ones(total), env.world, env.rank)
    return lo, hi


def leg_bwt(env, args):
    """configs[2]: 1 GiB random bytes, 256 blocks of 4 MiB, sharded contiguously over the ranks; encode and decode timed separately."""
    from tools import gen
    torch, ctx, world = env.torch, env.ctx, env.world
    lo, hi = _shard_range(env, COUNT)
    nb = hi - lo
    raw = gen.units("random", gen.unit_seed(3, lo), UNIT, nb, nthreads=env.nthreads)
    off = np.arange(nb, dtype=np.uint64) * UNIT
    n = np.full(nb, UNIT, dtype=np.uint64)
    d_raw = torch.from_numpy(raw).cuda()
    d_l = torch.zeros(UNIT * nb + 64, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(UNIT * nb, dtype=torch.uint8, device="cuda")
    origin, st = ctx.bwt_encode_blocks(d_raw, off, n, d_l, off)
    assert (st == 0).all()
    # RCZ_MEM_DEVICE (not _ASYNC): the call reads one counter per doubling round and stops as soon as every block is sorted; the blind
    # _ASYNC form would enqueue all ~20 rounds (1,950 launches, most of them empty) for data that resolves in round 0
    enc = lambda: ctx.bwt_encode_blocks(d_raw, off, n, d_l, off)                       # noqa: E731
    dec = lambda: ctx.bwt_decode_blocks(d_l, off, n, origin, d_back, off, async_=True)  # noqa: E731
    enc(); torch.cuda.synchronize()
    t0 = time.perf_counter(); enc(); torch.cuda.synchronize(); est_e = (time.perf_counter() - t0) * 1e3
    reps_e = env.reps_for(est_e, args.steps)
    ms_e = env.time_dev(enc, reps_e, 1)
    per_call_e = launches_of(ctx, enc)
    dec(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dec(); torch.cuda.synchronize(); est_d = (time.perf_counter() - t0) * 1e3
    reps_d = env.reps_for(est_d, args.steps)
    ms_d = env.time_dev(dec, reps_d, 3)
    per_call_d = launches_of(ctx, dec)
    assert torch.equal(d_back, d_raw), "decode(encode(x)) != x"
    kd_ms, kd_stage = stage_avg(ctx, dec)
    total_U = COUNT * UNIT
    alg = (2 * UNIT + 8) * nb
    # e2e: host buffers through the ABI
    h_raw, h_raw_np = env.pinned(raw)
    h_l = torch.zeros(UNIT * nb + 64, dtype=torch.uint8).pin_memory()
    h_back = torch.zeros(UNIT * nb + 64, dtype=torch.uint8).pin_memory()
    r = {}

    def enc_host():
        r["e"] = ctx.bwt_encode_blocks(h_raw_np, off, n, h_l.numpy(), off)
    e2e_e = env.time_host(enc_host, 2, 1)
    assert (r["e"][1] == 0).all() and (r["e"][0] == origin).all() and torch.equal(h_l[: UNIT * nb], d_l[: UNIT * nb].cpu())

    def dec_host():
        r["d"] = ctx.bwt_decode_blocks(h_l.numpy(), off, n, origin, h_back.numpy(), off)
    e2e_d = env.time_host(dec_host, 3, 1)
    assert (r["d"][1] == 0).all() and bytes(h_back.numpy()[: UNIT * nb]) == raw.tobytes()
    gather = None
    if world > 1:
        both = env.gather_time(dec, d_back, max(2, reps_d))
        gather = {"collective": "ncclAllGather of the decoded shards after the decode", "decode_plus_gather_ms": both, "value_with_gather": total_U / (both * 1e-3) / 1e9}
    cfg = {"workload": "bwt::Encoder / Decoder, 1 GiB random bytes (splitmix64) in 256 blocks of 4 MiB", "blocks_total": COUNT, "blocks_per_gpu": nb,
           "block_bytes": UNIT, "l2": "inputs larger than L2 (%.2f GiB per GPU per step)" % (2 * UNIT * nb / 2**30), "parallelism": "contiguous block ranges per rank (shard.partition), no data-path collective"}
    out = {}
    out["bwt_encode"] = {"config": cfg, "value": total_U / (ms_e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e, "steps": reps_e, "scaling": "strong", "dtype": "u8",
                         "roofline": roofline(env, "bwte::scatter_kernel (no capture: traffic null)", alg, ms_e, ms_e, {"note": "op-level figure: round 0 sorts 45-bit keys in 6 radix passes (hist, scan, scatter launches each), later rounds only the still-unresolved suffixes; kernel_ms = the whole call"}),
                         "e2e": {"value": total_U / e2e_e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": UNIT * nb, "d2h_bytes_per_step": (UNIT + 4) * nb, "ms_per_step": e2e_e * 1e3},
                         "gpu_launches_per_step": per_call_e}
    walk_ms = max(kd_stage) if kd_stage else kd_ms
    out["bwt_decode"] = {"config": cfg, "value": total_U / (ms_d * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_d, "steps": reps_d, "scaling": "strong", "dtype": "u8",
                         "roofline": roofline(env, "ibwt_walk_kernel", alg, walk_ms, kd_ms, {"stage_ms": kd_stage, "stages": ["hist+scan+scatter", "walk", "heads+headrank+place"],
                                                                                             "traffic_note": "ncu capture on 64 blocks, scaled to this rank's block count"}, traffic_scale=nb / 64.0),
                         "e2e": {"value": total_U / e2e_d / 1e9, "unit": "GB/s", "h2d_bytes_per_step": (UNIT + 4) * nb, "d2h_bytes_per_step": UNIT * nb, "ms_per_step": e2e_d * 1e3},
                         "gpu_launches_per_step": per_call_d}
    if gather:
        out["bwt_decode"]["gather"] = gather
        out["bwt_decode"]["value_with_gather"] = gather["value_with_gather"]
    return out


def leg_flate(env, args):
    """configs[3]: 131,072 raw-DEFLATE streams of 64 KiB (8 GiB), sharded over the ranks.  1 GiB (16,384 streams) is generated and
    compressed, and tiled to the rank's share."""
    torch, ctx, world = env.torch, env.ctx, env.world
    raw, packed, off, lens, span = make_flate_unique(env.nthreads)
    per_rank = FL_TOTAL // world
    tiles = max(1, per_rank // FL_UNIQUE)
    ns = FL_UNIQUE * tiles if per_rank >= FL_UNIQUE else per_rank
    uniq = min(FL_UNIQUE, ns)
    if ns < FL_UNIQUE:                                                   # more than 8 ranks: a prefix of the unique set
        span = int(off[uniq - 1] + ((lens[uniq - 1] + np.uint64(15)) & ~np.uint64(15)))
    in_off = np.concatenate([off[:uniq] + np.uint64(t * span) for t in range(tiles)])
    in_len = np.tile(lens[:uniq], tiles)
    out_off = np.arange(ns, dtype=np.uint64) * FL_UNIT
    caps = np.full(ns, FL_UNIT, dtype=np.uint64)
    d_in = torch.from_numpy(packed[:span]).cuda().repeat(tiles)
    d_in = torch.cat([d_in, torch.zeros(64, dtype=torch.uint8, device="cuda")])
    d_out = torch.zeros(FL_UNIT * ns, dtype=torch.uint8, device="cuda")
    C = int(in_len.sum())
    res = {}

    def dec():
        res["o"] = ctx.flate_decode_streams(d_in, in_off, in_len, d_out, out_off, caps, async_=True)
    dec(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dec(); torch.cuda.synchronize(); est = (time.perf_counter() - t0) * 1e3
    reps = env.reps_for(est, args.steps)
    ms = env.time_dev(dec, reps, 3)
    per_call = launches_of(ctx, dec)
    assert int((res["o"][1] != 0).sum().item()) == 0
    d_raw = torch.from_numpy(raw[: FL_UNIT * uniq]).cuda()
    for t in range(tiles):
        assert torch.equal(d_out[t * FL_UNIT * uniq: (t + 1) * FL_UNIT * uniq], d_raw), "inflate output differs from the generator's bytes"
    del d_raw
    k_ms, _ = stage_avg(ctx, dec, 2)
    total_U = FL_TOTAL * FL_UNIT
    # e2e with pinned host buffers of the rank's whole share
    h_in = torch.from_numpy(packed[:span]).repeat(tiles)
    h_in = torch.cat([h_in, torch.zeros(64, dtype=torch.uint8)]).pin_memory()
    h_out = torch.zeros(FL_UNIT * ns + 64, dtype=torch.uint8).pin_memory()
    r = {}

    def dec_host():
        r["o"] = ctx.flate_decode_streams(h_in.numpy(), in_off, in_len, h_out.numpy(), out_off, caps)
    e2e = env.time_host(dec_host, 2, 1)
    assert (r["o"][1] == 0).all() and bytes(h_out.numpy()[: FL_UNIT * uniq]) == raw[: FL_UNIT * uniq].tobytes()
    del h_in, h_out
    gather = None
    if world > 1:
        both = env.gather_time(dec, d_out, max(2, reps))
        gather = {"collective": "ncclAllGather of the decoded shards after the decode", "decode_plus_gather_ms": both, "value_with_gather": total_U / (both * 1e-3) / 1e9}
    cfg = {"workload": "flate::Decoder, 131,072 raw-DEFLATE streams of 64 KiB hexdump text (zlib level 6, wbits -15); 16,384 streams (1 GiB) generated, tiled %dx per GPU" % tiles,
           "streams_total": FL_TOTAL, "streams_per_gpu": ns, "stream_bytes": FL_UNIT, "compressed_bytes_per_gpu": C,
           "l2": "inputs larger than L2 (%.2f GiB per GPU per step)" % ((C + FL_UNIT * ns) / 2**30), "parallelism": "contiguous stream ranges per rank, no data-path collective"}
    leg = {"config": cfg, "value": total_U / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "steps": reps, "scaling": "strong", "dtype": "u8",
           "roofline": roofline(env, "inflate_kernel", C + FL_UNIT * ns, k_ms, k_ms, {"traffic_note": "ncu capture on 4,096 streams, scaled to this rank's stream count"}, traffic_scale=ns / 4096.0),
           "e2e": {"value": total_U / e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": tiles * span, "d2h_bytes_per_step": FL_UNIT * ns, "ms_per_step": e2e * 1e3},
           "gpu_launches_per_step": per_call}
    if gather:
        leg["gather"] = gather
        leg["value_with_gather"] = gather["value_with_gather"]
    return {"flate_decode": leg}


def leg_pipeline(env, args):
    """configs[4]: bwt -> dc -> entropy::ari, 256 text blocks of 4 MiB per GPU, chained on the device."""
    from tools import gen
    torch, ctx, world, rank = env.torch, env.ctx, env.world, env.rank
    nb = int(os.environ.get("RCZ_BENCH_C5_BLOCKS", COUNT))
    raw = gen.units("hextext", gen.unit_seed(5, rank * nb), UNIT, nb, nthreads=env.nthreads)
    off = np.arange(nb, dtype=np.uint64) * UNIT
    n = np.full(nb, UNIT, dtype=np.uint64)
    cap = 6 * UNIT                                                       # containers of this text come out at ~0.6 x the block
    coff = np.arange(nb, dtype=np.uint64) * cap
    caps = np.full(nb, cap, dtype=np.uint64)
    d_raw = torch.from_numpy(raw).cuda()
    d_cont = torch.zeros(cap * nb + 64, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(UNIT * nb, dtype=torch.uint8, device="cuda")
    clen, org, st = ctx.bwt_dc_ari_encode_blocks(d_raw, off, n, d_cont, coff, caps, ari_chunk=ARI_CHUNK)
    assert (st == 0).all()
    # (encode in RCZ_MEM_DEVICE, not _ASYNC: the suffix sort may then read its round counter and stop after the last live round
    #  instead of enqueueing all ~21 rounds; the call returns when the container is complete)
    enc = lambda: ctx.bwt_dc_ari_encode_blocks(d_raw, off, n, d_cont, coff, caps, ari_chunk=ARI_CHUNK)   # noqa: E731
    dec = lambda: ctx.bwt_dc_ari_decode_blocks(d_cont, coff, clen, d_back, off, n, ari_chunk=ARI_CHUNK, async_=True)  # noqa: E731
    t0 = time.perf_counter(); enc(); torch.cuda.synchronize(); est_e = (time.perf_counter() - t0) * 1e3
    reps_e = env.reps_for(est_e, args.steps, 4000.0)
    ms_e = env.time_dev(enc, reps_e, 1)
    per_call_e = launches_of(ctx, enc)
    ke_ms, ke_stage = stage_avg(ctx, enc, 1)
    dec(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dec(); torch.cuda.synchronize(); est_d = (time.perf_counter() - t0) * 1e3
    reps_d = env.reps_for(est_d, args.steps, 4000.0)
    ms_d = env.time_dev(dec, reps_d, 1)
    per_call_d = launches_of(ctx, dec)
    assert torch.equal(d_back, d_raw), "decode(encode(x)) != x"
    kd_ms, kd_stage = stage_avg(ctx, dec, 1)
    Cfin = int(clen.sum())
    # e2e on host buffers
    h_raw, h_raw_np = env.pinned(raw)
    h_cont = torch.zeros(cap * nb + 64, dtype=torch.uint8).pin_memory()
    h_back = torch.zeros(UNIT * nb + 64, dtype=torch.uint8).pin_memory()
    r = {}

    def enc_host():
        r["e"] = ctx.bwt_dc_ari_encode_blocks(h_raw_np, off, n, h_cont.numpy(), coff, caps, ari_chunk=ARI_CHUNK)
    e2e_e = env.time_host(enc_host, 2, 1)
    assert (r["e"][2] == 0).all() and (r["e"][0] == clen).all()

    def dec_host():
        r["d"] = ctx.bwt_dc_ari_decode_blocks(h_cont.numpy(), coff, clen, h_back.numpy(), off, n, ari_chunk=ARI_CHUNK)
    e2e_d = env.time_host(dec_host, 2, 1)
    assert (r["d"][1] == 0).all() and bytes(h_back.numpy()[: UNIT * nb]) == raw.tobytes()
    total_U = world * nb * UNIT
    cfg = {"workload": "bwt -> dc -> entropy::ari (bzip-style) on %d hexdump-text blocks of 4 MiB per GPU; dc output serialised as u32 LE, ByteEncoder streams of %d bytes" % (nb, ARI_CHUNK),
           "blocks_per_gpu": nb, "block_bytes": UNIT, "container_bytes_per_gpu": Cfin, "ari_chunk": ARI_CHUNK,
           "l2": "inputs larger than L2", "parallelism": "independent blocks, %d per GPU, no data-path collective" % nb,
           "parity": "ari / dc encode bytes are parity-unpinned (the reference tests them by round trip only); every stage equals the oracle's restatement"}
    e_dom = max(ke_stage) if ke_stage else ke_ms
    d_dom = max(kd_stage) if kd_stage else kd_ms
    names_e, names_d = ["bwt", "dc", "ari", "pack"], ["ari", "dc", "bwt"]
    return {
        "bwt_dc_ari_encode": {"config": cfg, "value": total_U / (ms_e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e, "steps": reps_e, "scaling": "weak", "dtype": "u8",
                              "roofline": roofline(env, "stage:" + (names_e[int(np.argmax(ke_stage))] if ke_stage else "?"), UNIT * nb + Cfin, e_dom, ke_ms, {"stage_ms": ke_stage, "stages": names_e}),
                              "e2e": {"value": total_U / e2e_e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": UNIT * nb, "d2h_bytes_per_step": Cfin, "ms_per_step": e2e_e * 1e3},
                              "gpu_launches_per_step": per_call_e},
        "bwt_dc_ari_decode": {"config": cfg, "value": total_U / (ms_d * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_d, "steps": reps_d, "scaling": "weak", "dtype": "u8",
                              "roofline": roofline(env, "stage:" + (names_d[int(np.argmax(kd_stage))] if kd_stage else "?"), UNIT * nb + Cfin, d_dom, kd_ms, {"stage_ms": kd_stage, "stages": names_d}),
                              "e2e": {"value": total_U / e2e_d / 1e9, "unit": "GB/s", "h2d_bytes_per_step": Cfin, "d2h_bytes_per_step": UNIT * nb, "ms_per_step": e2e_d * 1e3},
                              "gpu_launches_per_step": per_call_d},
    }


def run_ours(args):
    env = Env()
    torch = env.torch
    codecs = args.codecs.split(",")
    line = leg_lz4(env, args)
    line["config"]["numa"] = env.numa
    per = {}
    legs = [("bwt", leg_bwt), ("flate", leg_flate), ("pipeline", leg_pipeline)]
    for name, fn in legs:
        if name not in codecs:
            continue
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        try:
            got = fn(env, args)
            for k in got:
                got[k]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
            per.update(got)
            ok = 1.0
        except Exception as e:  # a failing secondary leg must not take the headline down; it is reported instead
            traceback.print_exc(file=sys.stderr)
            per[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            ok = 0.0
        if env.world > 1:                                                # keep the ranks in step whatever happened
            try:
                env.sum_over_ranks(ok)
            except Exception:
                pass
    if env.rank == 0 and env.world == 1 and per:
        try:
            cpu = cpu_legs(1, [c for c in codecs if c != "lz4"], single=True)
            for k, v in cpu.items():
                if k in per and "error" not in per[k]:
                    per[k]["cpu_baseline"] = v
        except Exception as e:  # pragma: no cover
            traceback.print_exc(file=sys.stderr)
            per["cpu_baseline_error"] = repr(e)
    if env.rank == 0:
        line["per_codec"] = per
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codecs", default="lz4,bwt,flate,pipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
