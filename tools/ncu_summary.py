#!/usr/bin/env python
"""Turn an ncu report into the text summary kept under profiles/ (the .ncu-rep itself stays in gpurun_out/, untracked).

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME.txt [--lines 30] [--note "..."]

Writes, per profiled launch: duration, DRAM bytes read/written (the `roofline.traffic` figure), throughput percentages,
occupancy, registers, shared-memory wavefronts / bank conflicts, issue utilisation; then the hottest CUDA source lines
(share of warp-stall samples and of executed instructions, top stall reasons) when the report carries source info."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__inst_executed.sum", "warp instructions executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem), blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 30
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    lines = ["# ncu summary of %s" % os.path.basename(rep)]
    if note:
        lines.append("# " + note)
    lines.append("# (cold-cache, serialised replays: compare shares and byte counts, not absolute times with bench.py)")
    traffic = {}
    for r in rows:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0]
        lines.append("")
        lines.append("kernel %s  (launch id %s)" % (name, d.get("ID", "?")))
        for k, label in KEYS:
            if k in d and d[k] != "":
                lines.append("  %-46s %s %s" % (label, d[k], u.get(k, "")))
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * scale.get(u["dram__bytes_read.sum"], 1)
            wr = float(d["dram__bytes_write.sum"]) * scale.get(u["dram__bytes_write.sum"], 1)
            lines.append("  %-46s %.0f bytes" % ("DRAM traffic (read + written)", rd + wr))
            traffic[name] = rd + wr
        except Exception:
            pass
    src = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    if "Line No" in src:
        tmp = out + ".src.csv"
        open(tmp, "w").write(src)
        top = subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), tmp, str(nlines)], stdout=subprocess.PIPE, text=True).stdout
        os.remove(tmp)
        lines += ["", "hottest source lines (first profiled launch with source):", top]
    open(out, "w").write("\n".join(lines) + "\n")
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
