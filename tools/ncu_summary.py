"""Turn an .ncu-rep capture (ncu --set full, one kernel) into the three-column summary kept under profiles/:
    python tools/ncu_summary.py gpurun_out/r02_pt_add.ncu-rep profiles/r02_pt_add_ncu.csv
Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and transposes the raw page: one `metric,unit,value` row per counter."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw)) if r]
# skip any ==PROF== banner lines: the header is the first row that starts with "ID"
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[h], rows[h + 1], rows[h + 2]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    kn = names.index("Kernel Name") if "Kernel Name" in names else None
    if kn is not None:
        w.writerow(["kernel", "", vals[kn]])
    for n, u, v in zip(names, units, vals):
        if n in ("ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size", "Grid Size", "Device", "CC",
                 "Section Name", "Metric Name", "Metric Unit", "Metric Value"):
            continue
        w.writerow([n, u, v])
print(f"{out}: {len(names)} columns")
