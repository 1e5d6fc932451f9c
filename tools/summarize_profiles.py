"""Turn gpurun_out/{launches.csv,prof_*.ncu-rep} (tools/profile.sh) into the tracked summaries under profiles/."""
import collections, csv, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof"
os.makedirs("profiles", exist_ok=True)
rows = [r for r in csv.reader(open(os.path.join(src, "launches.csv"))) if len(r) > 5]
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0][:60]
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`" % tag,
       "# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv (tools/profile.sh)",
       "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's `kernels.*.share_of_step`",
       "kernel,launches,total_us,avg_us,share"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append("\"%s\",%d,%.1f,%.1f,%.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
open("profiles/%s_launches_summary.csv" % tag, "w").write("\n".join(out) + "\n")
shutil.copy(os.path.join(src, "launches.csv"), "profiles/%s_launches_raw.csv" % tag)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["# %s: key metrics from `ncu --set full --clock-control none --import-source on` (tools/profile.sh), one B200" % tag, "file,kernel,metric,value,unit"]
for f in sorted(os.listdir(src)):
    if not f.endswith(".ncu-rep"):
        continue
    p = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(p.splitlines()))
    if len(rr) < 3:
        continue
    h = rr[0]
    r = rr[2]
    kn = r[h.index("Kernel Name")].split("(")[0]
    import re
    extra = [m for m in h if re.search(r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu)\.sum$|sm__pipe_(fmaheavy|fp64|xu)_cycles_active.avg.pct_of_peak_sustained_active|"
                                       r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__average_warp_latency_per_inst_issued.ratio", m)]
    for w in want + sorted(extra):
        if w in h:
            lines.append("%s,\"%s\",%s,%s,%s" % (f, kn, w, r[h.index(w)].replace(",", ""), rr[1][h.index(w)]))
    # warp stall reasons from the PC sampler: share of all samples, reasons >= 3 %
    samp = {}
    for m in h:
        mm = re.fullmatch(r"smsp__pcsamp_warps_issue_stalled_([a-z_]+?)", m)
        if mm and not m.endswith("_not_issued"):
            try:
                samp[mm.group(1)] = float(r[h.index(m)].replace(",", ""))
            except ValueError:
                pass
    tot_s = sum(samp.values())
    for reason, v in sorted(samp.items(), key=lambda kv: -kv[1]):
        if tot_s and v / tot_s >= 0.03:
            lines.append("%s,\"%s\",stall_%s,%.1f,%% of pc samples" % (f, kn, reason, 100.0 * v / tot_s))
open("profiles/%s_ncu_full_summary.csv" % tag, "w").write("\n".join(lines) + "\n")
for j in ("ntt_bench.json", "combine_bench.json", "chain_ubench.json", "sha_bench.json", "encode_bench.json", "mul_ubench.json", "per_row.json"):
    if os.path.exists(os.path.join(src, j)):
        shutil.copy(os.path.join(src, j), "profiles/%s_%s" % (tag, j))
print("profiles/ updated for", tag)
