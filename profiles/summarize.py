#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (run here, no GPU needed):
    python profiles/summarize.py gpurun_out/prof.ncu-rep [section substrings...]"""
import csv
import subprocess
import sys

SECTIONS = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis",
            "Warp State Statistics", "Scheduler Statistics", "Occupancy", "Launch Statistics")
RAW = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput",
       "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak",
       "launch__registers_per_thread", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "sm__throughput.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu", "sm__pipe_fma_cycles_active",
       "sm__pipe_alu_cycles_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    i_kern = hdr.index("Kernel Name")
    i_name, i_val, i_unit, i_sec = (hdr.index(k) for k in ("Metric Name", "Metric Value", "Metric Unit", "Section Name"))
    print("kernel:", rows[1][i_kern][:100])
    for r in rows[1:]:
        if r[i_sec] in SECTIONS:
            print(f"{r[i_sec][:30]:30s} {r[i_name][:46]:46s} {r[i_val]} {r[i_unit]}")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    print("--- raw")
    for n, u, v in zip(names, units, vals):
        if any(n.startswith(k) for k in RAW):
            print(f"{n:70s} {v} {u}")


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[1] == "--compact"):
    main()


def compact(rep):
    """One line per captured launch of a multi-kernel report (python profiles/summarize.py --compact rep)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names = rows[0]
    ix = {n: i for i, n in enumerate(names)}
    want = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "MB_rd", None), ("dram__bytes_write.sum", "MB_wr", None),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("sm__inst_executed.avg.per_cycle_active", "ipc", 1),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1), ("launch__registers_per_thread", "regs", 1),
            ("smsp__inst_executed.sum", "Minst", 1e-6), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1)]
    units = rows[1]
    for r in rows[2:]:
        out = [r[ix["Kernel Name"]][:60].ljust(60)]
        for key, label, scale in want:
            if key not in ix:
                continue
            try:
                v = float(r[ix[key]].replace(",", ""))
            except ValueError:
                continue
            u = units[ix[key]]
            if label.startswith("MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            elif label == "us":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
            elif scale != 1 and scale is not None:
                v *= scale
            out.append(f"{label}={v:.2f}")
        print(" ".join(out))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "--compact":
    compact(sys.argv[2])
