#!/usr/bin/env python
"""Updates profiles/traffic.json from an `ncu --set full` capture of lrf::render_kernel:

    python profiles/ncu_traffic.py <workload> <report.ncu-rep> [source note]

traffic = dram__bytes_read.sum + dram__bytes_write.sum of the captured launch (bench.py reads this file for
`roofline.traffic`; run where ncu is installed, no GPU needed)."""
import csv, io, json, os, subprocess, sys

wl, rep = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
row = [r for r in data if "render_kernel" in " ".join(r)][0]


def val(name):
    i = hdr.index(name)
    v = float(row[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
t = json.load(open(path)) if os.path.exists(path) else {}
t[wl] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
         "source": f"ncu --set full, {note} (profiles/)"}
json.dump(t, open(path, "w"), indent=1)
print(wl, t[wl])
