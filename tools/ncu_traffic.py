#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of the fused Jacobian kernel:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch and per element, plus the pipe utilisations
quoted in DESIGN.md.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep NELEM "source note" > profiles/traffic.json
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, nelem, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[2:] if "k_fused_apply" in r[col["Kernel Name"]]]
    r = data[-1]

    def val(name, scale=True):
        v = float(r[col[name]].replace(",", ""))
        return v * UNIT.get(units[col[name]], 1.0) if scale else v
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    rec = {"kernel": r[col["Kernel Name"]], "elements": nelem, "launches_in_capture": len(data),
           "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_launch": rd + wr,
           "dram_bytes_per_element": (rd + wr) / nelem,
           "gpu_time_under_ncu": f'{val("gpu__time_duration.sum", False)} {units[col["gpu__time_duration.sum"]]}',
           "l1tex_lsu_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
           "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
           "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
           "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
           "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active", False),
           "registers_per_thread": val("launch__registers_per_thread", False),
           "smem_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False),
           "source": note}
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
