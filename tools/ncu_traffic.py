"""profiles/r02_traffic.json from the committed `ncu --set full` capture of one C2 step (gpurun_out/r02_join.ncu-rep):
dram__bytes_read.sum + dram__bytes_write.sum of the probe-phase launches (k_fj_hist / k_fj_scatter / k_fj_probe over the
1 B probe rows), per probe row.  bench.py multiplies it by the rows of its own run for `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def gb(r, m):
    v, u = float(r[ix[m]]), units[ix[m]].lower()
    return v * {"gbyte": 1.0, "mbyte": 1e-3, "kbyte": 1e-6, "byte": 1e-9}[u]


launches = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    launches.append({"kernel": name.split("(")[0].replace("<unnamed>::", "").replace("void ", ""), "dram_gb": gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum")})
# one step = build (hist, scatter, build_part) then probe (hist, scatter, probe): the probe side is the LAST hist/scatter + the probe kernel
probe = [l for l in launches if "probe" in l["kernel"]]
hists = [l for l in launches if "k_fj_hist" in l["kernel"]]
scats = [l for l in launches if "k_fj_scatter" in l["kernel"]]
total = probe[-1]["dram_gb"] + hists[-1]["dram_gb"] + scats[-1]["dram_gb"]
res = {"source": rep.split("/")[-1], "probe_rows": 1_000_000_000, "launches": launches,
       "radix": {"dram_bytes_per_probe_row": total, "dram_gb_per_1e9_rows": total,
                 "kernels": {"k_fj_hist": hists[-1]["dram_gb"], "k_fj_scatter": scats[-1]["dram_gb"], "k_fj_probe": probe[-1]["dram_gb"]}}}
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res["radix"]))
