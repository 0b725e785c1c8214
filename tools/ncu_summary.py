#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into a small text table for profiles/ (the reports themselves are scratch)."""
import csv, io, subprocess, sys
WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("smsp__inst_executed.sum", "warp_instr"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("lts__t_bytes.sum", "l2_bytes_MB"), ("sm__cycles_elapsed.max", "cycles")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("## " + rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name[name.find("::") + 2:][:70] if "::" in name else name[:70]
        parts = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    u = units[i]
                    if label.endswith("_MB") and u in ("byte", "Mbyte", "Gbyte", "Kbyte"):
                        f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                    v = "%.4g" % f
                except ValueError:
                    pass
                parts.append("%s=%s" % (label, v))
        print("  %s\n    %s" % (name, "  ".join(parts)))
