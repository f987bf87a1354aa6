"""Turns gpurun_out/*.ncu-rep captures into the small text/JSON summaries committed under profiles/.

    python tools/ncu_summarise.py gpurun_out/prof_als_r1h.ncu-rep k_als 296 profiles/r01_k_als_ncu.txt
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, kernel, clips, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    get = lambda k: (v[h.index(k)], u[h.index(k)]) if k in h else (None, None)
    lines = [f"ncu --set full --clock-control none --import-source on, kernel {v[h.index('Kernel Name')][:80]}", f"capture: {rep}; {note}", ""]
    for k in KEYS + [n for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]:
        val, unit = get(k)
        if val is not None and val not in ("0", "0.000000"):
            lines.append(f"{k:88s} {unit:16s} {val}")
    open(out, "w").write("\n".join(lines) + "\n")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(float(get(k)[0]) * scale[get(k)[1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    sj = os.path.join(os.path.dirname(out), "ncu_summary.json")
    data = json.load(open(sj)) if os.path.exists(sj) else {}
    dur, dunit = get("gpu__time_duration.sum")
    data[kernel] = {"dram_bytes": dram, "clips": clips, "duration": f"{dur} {dunit}", "source": os.path.basename(out)}
    json.dump(data, open(sj, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
