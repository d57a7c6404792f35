"""profiles/<kind>_traffic_<config>.json from an ncu report of ONE steady-state launch + the steady-state rays per launch that
tools/profile_run.py printed in the same run (the figures bench.py's roofline / roofline_onchip quote beside the live timing).

    python tools/make_traffic_json.py <report.ncu-rep> <run log> <trace|shade> <config> <source text>"""
import csv, io, json, os, re, subprocess, sys
rep, log, kind, config, source = sys.argv[1:6]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
def g(name):
    v = vals[hdr.index(name)].replace(",", "")
    u = rows[1][hdr.index(name)]
    f = float(v)
    if "byte" in u:
        return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    return f * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)          # durations in ms
m = re.search(r"STEADY iterations (\d+) ext_per_launch ([\d.]+) shadow_per_launch ([\d.]+) pool_slots (\d+)", open(log).read())
ext, sh, pool = float(m.group(2)), float(m.group(3)), int(m.group(4))
rays = ext + sh
out = {"kernel": vals[hdr.index("Kernel Name")], "source": source,
       "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
       "dram_bytes_per_launch": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
       "launch_ms_under_ncu": g("gpu__time_duration.sum"),
       "pool_slots": pool, "rays_in_launch": {"extend": ext, "shadow": sh},
       "warp_inst": g("smsp__inst_executed.sum"), "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
       "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")}
if kind == "trace":
    out["warp_inst_per_ray"] = out["warp_inst"] / rays
    out["smem_wavefronts_per_ray"] = out["smem_wavefronts"] / rays
    out["dram_bytes_per_ray"] = out["dram_bytes_per_launch"] / rays
path = os.path.join(ROOT, "profiles", "%s_traffic_%s.json" % (kind, config))
json.dump(out, open(path, "w"), indent=1)
print(path, json.dumps(out)[:300])
