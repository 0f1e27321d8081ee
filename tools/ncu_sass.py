"""Per-SASS-instruction stall breakdown of the hot region of a kernel in an ncu report (needs --import-source on / --set full):
   python tools/ncu_sass.py X.ncu-rep [min_share_percent]
Prints address, share of warp-stall samples, executed count, the instruction and its three largest stall reasons."""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    print("total samples", tot)
    for r in data:
        s = int(r[idx["# Samples"]] or 0)
        if s < tot * thr / 100:
            continue
        st = sorted(((int(r[idx[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
        print(r[idx["Address"]][-5:], "%5.1f%%" % (100 * s / tot), r[idx["Instructions Executed"]].rjust(10), r[idx["Source"]][:64].ljust(64),
              " ".join("%s:%d" % (k, v) for v, k in st if v))


if __name__ == "__main__":
    main()
