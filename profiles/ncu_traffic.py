"""Writes profiles/r02_traffic.json from an `ncu --page raw --csv` dump: dram read / write bytes and duration per kernel
(first profiled launch of each kernel name).  This file is what bench.py's `roofline.traffic` reads -- generated, not typed.
Usage: ncu -i x.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_traffic.py raw.csv profiles/r02_traffic.json"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
col = {n: hdr.index(n) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
out = {}
for r in rows[2:]:
    name = re.sub(r"^(void )?(pillars::)?(<unnamed>::|unnamed>::)*", "", r[ik]).split("(")[0]
    if name in out:
        continue
    rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
    dur = float(r[col["gpu__time_duration.sum"]])
    out[name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration": dur,
                 "duration_unit": units[col["gpu__time_duration.sum"]]}
json.dump({"source": sys.argv[1], "note": "ncu --set full --clock-control none, one launch per kernel, cfg2 (B = 16)",
           "kernels": out}, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
