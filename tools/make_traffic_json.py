"""profiles/r2/traffic.json from an `ncu --page raw --csv` export of one launch of the persistent kernel.
usage: python tools/make_traffic_json.py RAW.csv OUT.json "<source note>" """
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(k):
    i = h.index(k)
    return float(v[i].replace(",", "")) * scale.get(u[i], 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"kernel": v[h.index("Kernel Name")], "workload": "3D Poisson GDM p=3, 256^3 cells, stiffness apply",
           "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr), "dram_bytes_per_launch": int(rd + wr),
           "algorithmic_bytes_per_launch": 271593488, "gpu_time_duration_us_under_ncu": val("gpu__time_duration.sum"),
           "source": sys.argv[3]}, open(sys.argv[2], "w"), indent=1)
print(open(sys.argv[2]).read())
