#!/usr/bin/env python
"""Turn gpurun_out/{launches_r1.csv, prof_r1.ncu-rep} into the tracked summaries under profiles/.
Run here (no GPU needed; needs `ncu` for reading the report)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof_r1.ncu-rep")
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches_r1.csv")
nelem = int(sys.argv[3]) if len(sys.argv) > 3 else 32768


def ncu_csv(page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


# ---- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("b200::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
lt = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    lt.append(f"| `{k[:90]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |")
fused = next((a for k, a in agg.items() if k.startswith("k_fused_apply<5, 5, 2, 1")), [1, 0.0])
per_step = {k: a[1] / a[0] for k, a in agg.items()}
open(os.path.join(ROOT, "profiles", "r1_launches.md"), "w").write(f"""# Round 1 — ncu launch list (`gpu__time_duration.sum`)

Command (gpurun, 1 GPU, after the same command exited 0 without ncu):

    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \\
        python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --box 32

Whole process (set-up + 3 warm-up + 2 timed MatMults + 13 kernel-only applies), box 32^3.  Per-launch times are
cold-cache and serialised; the SHARE is what matters.

{chr(10).join(lt)}

Per timed step (masked DM layout, the bench default) the launches are: one fill kernel (CeedOperatorApply zeroing its
output), `k_fused_apply<5,5,2,1>` (the whole CeedOperatorApply, reading X and accumulating into Y directly) and
`k_mask_zero` (Dirichlet rows of Y).  The fused kernel is {fused[1] / fused[0]:.0f} us of the ~{fused[1] / fused[0] + per_step.get('k_mask_zero', 0) + 9:.0f} us step here
(~{100 * (fused[1] / fused[0]) / (fused[1] / fused[0] + per_step.get('k_mask_zero', 0) + 9):.0f} %), matching bench.py at 64^3 (1.238 ms of 1.317 ms = 94 %).  `k_fused_apply<5,5,2,0>` is the residual evaluation that fills
gradu once; `k_restrict_strided`, `k_basis_apply`, `k_qfunction` are the one-off SetupGeo operator on the generic path;
`k_jcache_build<2>` builds the Jacobian cache once.
""")

# ---- full capture
raw = ncu_csv("raw")
h, units, r = raw[0], raw[1], raw[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
vals, tbl = {}, []
for k in keys:
    if k in h:
        i = h.index(k)
        vals[k] = (r[i], units[i])
        tbl.append(f"| `{k}` | {r[i]} | {units[i]} |")
sass = ncu_csv("source", ("--print-source", "sass"))
h2, data = sass[1], sass[2:]
ix = {k: i for i, k in enumerate(h2)}


def f(rr, k):
    try:
        return float(rr[ix[k]])
    except Exception:
        return 0.0


stalls = [k for k in h2 if k.startswith("stall_") and "Not Issued" not in k]
totst = {k: sum(f(rr, k) for rr in data) for k in stalls}
s = sum(totst.values())
st = ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in sorted(totst.items(), key=lambda x: -x[1])[:10])
ag = collections.defaultdict(lambda: [0, 0, 0, 0])
for rr in data:
    src = rr[ix["Source"]].split()
    if not src:
        continue
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    a = ag[op]
    a[0] += f(rr, "Instructions Executed"); a[1] += f(rr, "L1 Wavefronts Shared"); a[2] += f(rr, "L1 Wavefronts Shared Ideal")
    a[3] += f(rr, "L2 Theoretical Sectors Global")
ops = ["| opcode | warp instr | smem wavefronts | ideal | L2 sectors (theoretical) |", "|---|---|---|---|---|"]
for k, v in sorted(ag.items(), key=lambda x: -x[1][0])[:12]:
    ops.append(f"| {k} | {v[0]:.0f} | {v[1]:.0f} | {v[2]:.0f} | {v[3]:.0f} |")
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
dr = float(vals["dram__bytes_read.sum"][0]) * mult[vals["dram__bytes_read.sum"][1]]
dw = float(vals["dram__bytes_write.sum"][0]) * mult[vals["dram__bytes_write.sum"][1]]
json.dump({"kernel": "k_fused_apply<5,5,hyperFS,Jacobian>", "elements": nelem, "dram_bytes_per_launch": dr + dw,
           "dram_bytes_per_element": (dr + dw) / nelem,
           "source": "profiles/r1_k_fused_apply_full.md (ncu --set full, box 32^3)"},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
g = lambda k: float(vals[k][0])
open(os.path.join(ROOT, "profiles", "r1_k_fused_apply_full.md"), "w").write(f"""# Round 1 — `ncu --set full` capture of the dominant kernel

Command (gpurun, 1 GPU, after the same command exited 0 without ncu):

    ncu --set full --clock-control none --import-source on -k regex:k_fused_apply -s 5 -c 1 \\
        python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --box 32

Kernel: `k_fused_apply<P=5,Q=5,hyperFS,Jacobian>` on a 32^3 box ({nelem} elements, 6.44 M DoFs; the per-point
stream is 557 MB, larger than L2).  Numbers under ncu are cold-cache and serialised: read them as ratios.

| metric | value | unit |
|---|---|---|
{chr(10).join(tbl)}

DRAM traffic per launch: {dr / 1e6:.1f} MB read + {dw / 1e6:.1f} MB written = {(dr + dw) / nelem:.0f} B/element
(algorithmic figure of SURVEY.md 8(d): 22 572 B/element with the reference's 19 doubles per point; this kernel
streams 17 doubles per point from the Jacobian cache: 17 000 + 500 + 3 072 = 20 572 B/element).

Warp stall samples (all): {st}

{chr(10).join(ops)}

Reading: shared-memory bank conflicts are gone except in the scatter loop (wavefronts vs ideal above); the FP64 pipe is
~{g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.0f} % busy, the L1TEX data pipe ~{g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} %, DRAM ~{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} % of ncu's nominal peak, issue slots
~{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} %.  No unit is saturated: the kernel is latency-bound (long scoreboard = waits on the per-point stream, the two
dependent loads of the gather and the scatter's offset loads; `wait` = fixed-latency DFMA chains).  DRAM traffic is within
~9 % of the 20.6 kB/element the kernel must move, i.e. no wasted re-reads.  Next steps: DESIGN.md "What limits the kernel".
""")
print("wrote profiles/r1_launches.md, r1_k_fused_apply_full.md, traffic.json")
