"""Per-kernel DRAM traffic of one bench step from an `ncu --set full` capture of tools/prof_step.py:
   ncu --profile-from-start off --set full --clock-control none -k regex:^k_ -o gpurun_out/step python tools/prof_step.py
   ncu -i gpurun_out/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv
   python tools/ncu_traffic.py gpurun_out/step_raw.csv            -> profiles/r02_traffic.json (+ a table on stdout)
bench.py quotes `roofline.traffic` from that file only while the library sources are the ones it was captured on."""
import csv, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
body = [r for r in rows[2:] if len(r) == len(h)]
col = {n: i for i, n in enumerate(h)}
def f(r, n):
    try: return float(r[col[n]].replace(",", ""))
    except Exception: return 0.0
per = {}
table = []
for r in body:
    name = r[col["Kernel Name"]]
    m = re.search(r"(k_\w+)", name)
    short = m.group(1) if m else name.split("(")[0]
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    ur, uw = h[col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_read.sum"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][col["dram__bytes_read.sum"]], 1)
    scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][col["dram__bytes_write.sum"]], 1)
    b = rd * scale + wr * scale_w
    per[short] = per.get(short, 0.0) + b
    table.append((short, f(r, "gpu__time_duration.sum"), rd * scale / 1e6, wr * scale_w / 1e6))
out = {"source_digest": bench.source_digest(), "how": "ncu --set full --clock-control none of one eager 64-pair step (tools/prof_step.py), "
       "dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel's launches in the step",
       "dram_bytes_per_step": per, "launches": [dict(kernel=k, us=t, dram_read_MB=a, dram_write_MB=b) for k, t, a, b in table]}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
json.dump(out, open(path, "w"), indent=1)
for k, t, a, b in table:
    print(f"{k:28s} {t:9.1f} us  read {a:8.1f} MB  write {b:8.1f} MB")
print("->", path)
