#!/usr/bin/env python3
"""DRAM traffic of one image's U-Net launches from an `ncu --set full` report -> profiles/unet_dram_traffic.json
(read by bench.py for roofline.traffic).  usage: ncu_traffic.py <report.ncu-rep> <summary name it belongs to>"""
import csv, io, json, subprocess, sys

rep, src = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}


def b(d, name):
    i = col[name]
    return float(d[i].replace(",", "")) * scale.get(units[i].lower(), 1)


per = []
for d in data:
    name = d[col["Kernel Name"]]
    per.append({"kernel": name.split("(")[0].replace("void ecseg::<unnamed>::", ""), "dram_read": b(d, "dram__bytes_read.sum"),
                "dram_write": b(d, "dram__bytes_write.sum")})
total = sum(p["dram_read"] + p["dram_write"] for p in per)
out = {"bytes_per_image": total, "launches": len(per), "source": src,
       "what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the U-Net launches of one 2048x2048 image (100 tiles), ncu --set full",
       "per_launch": per}
json.dump(out, open("profiles/unet_dram_traffic.json", "w"), indent=1)
print(f"{len(per)} launches, {total / 1e9:.3f} GB per image")
