#!/usr/bin/env python
"""Phase-level reading of an `ncu --set full --import-source on` capture: the SASS of a kernel is cut at its
BAR.SYNC instructions (the fused kernels separate their stages with one barrier each) and the warp-stall samples,
executed warp instructions and shared-memory wavefronts are summed per stage.

    python tools/ncu_phases.py gpurun_out/prof.ncu-rep [kernel-substring]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def new_phase():
    return {"samples": 0, "inst": 0, "ops": Counter(), "stall": Counter(), "wav": 0, "ideal": 0}


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the page is a sequence of per-kernel tables, each introduced by a "Kernel Name" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    for b in blocks:
        if want not in b["name"]:
            continue
        h = {k: i for i, k in enumerate(b["hdr"])}
        stall_cols = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
        phases, ph = [], new_phase()
        for r in b["rows"]:
            src = r[h["Source"]].strip()
            toks = src.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.split(".")[0]
            ph["samples"] += int(r[h["# Samples"]] or 0)
            ph["inst"] += int(r[h["Instructions Executed"]] or 0)
            ph["ops"][op] += int(r[h["Instructions Executed"]] or 0)
            ph["wav"] += int(r[h["L1 Wavefronts Shared"]] or 0)
            ph["ideal"] += int(r[h["L1 Wavefronts Shared Ideal"]] or 0)
            for k in stall_cols:
                ph["stall"][k[6:]] += int(r[h[k]] or 0)
            if op == "BAR":
                phases.append(ph)
                ph = new_phase()
        phases.append(ph)
        tot = sum(p["samples"] for p in phases) or 1
        print(f"## {b['name'][:110]}\n")
        print("| stage | samples % | warp instr | smem wavefronts (ideal) | top stalls | top opcodes |")
        print("|---|---|---|---|---|---|")
        for i, p in enumerate(phases):
            st = ", ".join(f"{k} {100 * v / max(p['samples'], 1):.0f}%" for k, v in p["stall"].most_common(4))
            ops = ", ".join(f"{k} {v}" for k, v in p["ops"].most_common(6))
            print(f"| {i} | {100 * p['samples'] / tot:.1f} | {p['inst']} | {p['wav']} ({p['ideal']}) | {st} | {ops} |")
        allst = Counter()
        for p in phases:
            allst.update(p["stall"])
        print("\nall stages: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in allst.most_common(10)) + "\n")


if __name__ == "__main__":
    main()
