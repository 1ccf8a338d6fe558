#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of the rates pair kernel at the bench size (read by bench.py for roofline.traffic).

    python tools/make_traffic.py gpurun_out/<tag>/prof_rates_nx512.ncu-rep slab512 512 profiles/r02/ncu_details_rates_nx512.txt

The source hash ties the file to the build it was captured from: bench.py refuses a traffic.json whose hash differs from the sources it runs."""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_count import source_sha

rep, config, nx = sys.argv[1], sys.argv[2], int(sys.argv[3])
report_name = sys.argv[4] if len(sys.argv) > 4 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(key, scale_units=True):
    v, u = d[key]
    x = float(v.replace(",", ""))
    if scale_units:
        x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return x


rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
out = {
    "config": config, "nx": nx, "kernel": d["Kernel Name"][0].split("(")[0].replace("void ", ""), "source_sha": source_sha(),
    "rates_pair_kernel_dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
    "l1tex_throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_active", False),
    "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct", False), "l2_hit_pct": num("lts__t_sector_hit_rate.pct", False),
    "registers_per_thread": num("launch__registers_per_thread", False),
    "gpu_time_ms_under_ncu": num("gpu__time_duration.sum"),
    "ncu_report": f"{report_name} (ncu --set full --clock-control none --import-source on, bench.py --nx {nx} --steps 1 --warmup 3 --no-cpu, one launch)",
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
