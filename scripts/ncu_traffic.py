"""Writes profiles/r2_fused_traffic.json from an `ncu --set full` report: DRAM bytes read / written per launch of the
dominant kernel, which bench.py reports as roofline.traffic (never a literal in bench.py).
usage: python scripts/ncu_traffic.py gpurun_out/X.ncu-rep "<command that produced it>" [kernel substring]"""
import csv, io, json, os, subprocess, sys
rep, command = sys.argv[1], sys.argv[2]
sub = sys.argv[3] if len(sys.argv) > 3 else "fused_step3_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
d = [r for r in data if sub in r[ki]][-1]


def val(name):
    i = hdr.index(name)
    v = float(d[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"kernel": d[ki], "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "gpu_time_ms": float(d[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower().replace("second", "s").replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1.0),
       "commit": commit, "command": command, "file": os.path.basename(rep)}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_fused_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
