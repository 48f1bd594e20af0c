#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch
with duration, DRAM bytes, tensor-pipe activity, issue utilisation, registers.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_name.csv ["note"]"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("smsp__inst_executed.sum", "warp_instructions"),
        ("sm__cycles_elapsed.max", "sm_cycles"), ("launch__grid_size", "grid"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem")]


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(k), name) for k, name in WANT if k in hdr]
    with open(dst, "w") as f:
        f.write("# %s\n# source: %s (ncu --set full --clock-control none --import-source on)\n" % (note, rep))
        f.write(",".join("%s[%s]" % (name, units[i]) if units[i] else name for i, name in cols) + "\n")
        for r in data:
            vals = []
            for i, name in cols:
                v = r[i]
                if name == "kernel":
                    v = v.split("(")[0].replace("void ", "").replace(",", ";")
                vals.append(v.replace(",", ""))
            f.write(",".join(vals) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
