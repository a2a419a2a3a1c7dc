"""Select the judged columns from an `ncu --set full` report: ncu_select.py report.ncu-rep case-label... > selected.csv
(one label per captured launch, in capture order; launches beyond the labels keep an empty label)."""
import csv, io, subprocess, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_dim_x",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
idx = {c: head.index(c) for c in COLS if c in head}
k = head.index("Kernel Name")
labels = sys.argv[2:]
w = csv.writer(sys.stdout)
w.writerow(["case", "Kernel Name"] + list(idx))
w.writerow(["", ""] + [units[i] for i in idx.values()])
for n, r in enumerate(body):
    w.writerow([labels[n] if n < len(labels) else "", r[k]] + [r[i] for i in idx.values()])
