"""Turn the ncu artefacts of one round (gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_coupler.ncu-rep)
into the tracked summaries under profiles/.   usage: python scripts/summarize_ncu.py r01c"""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")

# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(go, f"{tag}_launches.csv"))) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name, val, unit = r[4], float(r[-1].replace(",", "")), r[-2]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
with open(os.path.join(pr, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"# {tag} launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --c5-n 0\n")
    f.write("# (cold-cache, serialised per-launch times: compare SHARES, not absolutes). First 400 launches of the process.\n")
    f.write(f"{'kernel':90s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{name[:90]:90s} {n:8d} {us:12.1f} {us/n:10.1f} {100*us/tot:6.1f}%\n")
import shutil
shutil.copy(os.path.join(go, f"{tag}_launches.csv"), os.path.join(pr, f"{tag}_launches.csv"))

# ---- full capture --------------------------------------------------------------------------------
rep = os.path.join(go, f"{tag}_coupler.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
traffic = {}
with open(os.path.join(pr, f"{tag}_coupler_yee_kernels.txt"), "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none --import-source on -k regex:yee_ -s 10 -c 2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --c5-n 0 (C2 coupler cpl=20, 70.7 Mcell)\n\n")
    for r in rr[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        f.write(f"{'Kernel Name':90s} {d['Kernel Name']}\n")
        for w in want:
            if w in d:
                f.write(f"{w:90s} {d[w]} {u[w]}\n")
        st = [(float(d[h]), h) for h in hdr if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
        for v, h in sorted(st, reverse=True)[:6]:
            f.write(f"{'stall: ' + h.split('issue_stalled_')[1]:90s} {v:.3f} inst\n")
        def gb(key):
            v, un = float(d[key].replace(",", "")), u[key]
            return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[un]
        t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        f.write(f"{'=> DRAM traffic per launch (read+write)':90s} {t:.4f} GB\n\n")
        traffic["yee_E" if "yee_E" in d["Kernel Name"] else "yee_H"] = t * 1e9
tj = os.path.join(pr, "traffic.json")
cur = json.load(open(tj)) if os.path.exists(tj) else {}
cur["coupler"] = traffic.get("yee_E", cur.get("coupler"))
cur["coupler_yee_H"] = traffic.get("yee_H")
cur["source"] = f"profiles/{tag}_coupler_yee_kernels.txt"
sys.path.insert(0, root)
import bench  # noqa: E402

cur["csrc_sha16"] = bench._csrc_sha()  # bench.py reports these figures only while the kernel sources are unchanged
json.dump(cur, open(tj, "w"))
print(open(os.path.join(pr, f"{tag}_launches_summary.txt")).read())
print(open(os.path.join(pr, f"{tag}_coupler_yee_kernels.txt")).read())
