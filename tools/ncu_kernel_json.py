"""profiles/rN_ncu_<kernel>_N<n>.json from an .ncu-rep: the numbers bench.py quotes beside its roofline (traffic, FP64 pipe)
with the commit they were captured at. usage: ncu_kernel_json.py REP KERNEL_SUBSTRING N_PARTICLES OUT.json"""
import csv, json, subprocess, sys
rep, kern, n, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def val(r, name):
    i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3}.get(u, 1)
r = [r for r in rows[2:] if kern in r[hdr.index("Kernel Name")]][-1]
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "n_particles": n, "commit": commit,
     "duration_us": val(r, "gpu__time_duration.sum"),
     "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
     "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "registers": val(r, "launch__registers_per_thread"), "inst_executed": val(r, "smsp__inst_executed.sum"),
     "how": "ncu --set full --clock-control none, second of two steps of tools/prof_one.py (cold cache, serialised)"}
json.dump(d, open(out, "w"), indent=1)
print(d)
