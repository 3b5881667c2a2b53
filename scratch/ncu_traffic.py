"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in `ncu --set full` raw-page CSVs
-> profiles/r2_traffic.json (read by bench.py for roofline.traffic).
    python scratch/ncu_traffic.py out.json a_raw.csv [b_raw.csv ...]"""
import csv, json, sys
from collections import defaultdict
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = defaultdict(list)
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, ri, wi, ti = (hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    for r in data:
        try:
            b = float(r[ri].replace(",", "")) * SCALE[units[ri]] + float(r[wi].replace(",", "")) * SCALE[units[wi]]
            t = float(r[ti].replace(",", ""))
        except (ValueError, KeyError):
            continue
        if t < 10.0 and units[ti] in ("us", "usecond"):
            continue      # no-op launches (done flag set)
        agg[r[ki]].append(b)
out = {k: {"dram_bytes_per_launch": sum(v) / len(v), "captured_launches": len(v)} for k, v in agg.items()}
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
