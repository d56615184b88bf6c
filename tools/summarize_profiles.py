"""Turns the raw artefacts of tools/profile_round.sh (gpurun_out/) into the tracked summaries under
profiles/:  launches_<tag>_summary.txt, ncu_kernels_<tag>.txt, roofline_traffic.json, bench_<tag>.json.
    python tools/summarize_profiles.py r1
"""
import collections, csv, io, json, os, re, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list -> shares of one eager step
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
seq = []
for r in rows[1:]:
    t = float(r[ix["Metric Value"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[ix["Metric Unit"]]]
    seq.append((r[ix["Kernel Name"]], t))
starts = [i for i, (n, _) in enumerate(seq) if "pack_cameras" in n]   # first launch of a decoder call
step = seq[starts[0]:starts[1]]
tot = sum(t for _, t in step)
agg = collections.OrderedDict()
for n, t in step:
    k = re.sub(r"\(.*", "", n).replace("void ", "")[:64]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += t
mvg_t = sum(t for n, t in step if "mvg::" in n)
with open(os.path.join(P, f"launches_{tag}_summary.txt"), "w") as f:
    f.write(f"ncu launch list of ONE eager decoder step (bench.py --no-graph), B=1 V=5 Q=1024 L=4, B200\n"
            f"command: ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 400 --csv python bench.py "
            f"--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --batch8 0   (tools/profile_round.sh)\n"
            f"(per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n\n"
            f" count    total_us   share  kernel\n")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{c:6d} {t:11.1f} {100 * t / tot:6.1f}%  {k}\n")
    f.write(f"\n{len(step)} launches, {tot:.1f} us summed kernel time\n"
            f"libmvg_b200 kernels: {mvg_t:.1f} us = {100 * mvg_t / tot:.1f}% of the step; the rest are small torch "
            f"copy / cat kernels of the host glue (stacking the per-layer outputs)\n")
print(open(os.path.join(P, f"launches_{tag}_summary.txt")).read())

# ---- ncu --set full of one layer -> per-kernel metrics
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"step_{tag}.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
hx = {h: i for i, h in enumerate(hdr)}
cols = [("dur_us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("dram_rd_MB", "dram__bytes_read.sum"),
        ("dram_wr_MB", "dram__bytes_write.sum"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("l1_datapipe%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct")]
units = rr[1]
def val(r, name):
    i = hx.get(name)
    if i is None: return float("nan")
    v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else float("nan")
    u = units[i]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    return v
traffic = None
with open(os.path.join(P, f"ncu_kernels_{tag}.txt"), "w") as f:
    f.write("ncu --set full --clock-control none, the libmvg_b200 launches of one decoder layer (+ camera packing, pyramid\n"
            "hand-off and value GEMM of the call) in an eager step (bench.py --no-graph), B200 (tools/profile_round.sh)\n"
            "columns: " + ", ".join(c for c, _ in cols) + "\n\n")
    for r in rr[2:]:
        name = re.sub(r"\(.*", "", r[hx["Kernel Name"]]).replace("void ", "")
        f.write(name + "\n   " + "  ".join(f"{c}={val(r, m):.4g}" for c, m in cols) + "\n")
        if any(k in name for k in ("gather_kernel", "project_bin", "bin_scan", "bin_scatter", "sample_params",
                                   "gather_tiles", "gather_direct")):
            # the gather STAGE = every kernel mvg_project_sample_fused launches (first layer in the capture)
            traffic = traffic or dict(dram_read_bytes=0, dram_write_bytes=0, kernels=[])
            if name not in traffic["kernels"]:
                traffic["kernels"].append(name)
                traffic["dram_read_bytes"] += int(val(r, "dram__bytes_read.sum") * 1e6)
                traffic["dram_write_bytes"] += int(val(r, "dram__bytes_write.sum") * 1e6)
print(open(os.path.join(P, f"ncu_kernels_{tag}.txt")).read())
if traffic:
    traffic["project_sample_fused_dram_bytes_per_launch"] = traffic["dram_read_bytes"] + traffic["dram_write_bytes"]
    traffic["launches_averaged"] = 1
    traffic["source"] = (f"profiles/ncu_kernels_{tag}.txt (ncu --set full, summed over the kernels of one layer's "
                         f"mvg_project_sample_fused call; gpurun_out/step_{tag}.ncu-rep)")
    json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    print(traffic)
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"bench_{tag}.json"))
