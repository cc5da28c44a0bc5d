"""gpurun_out/launches.csv + gpurun_out/prof.ncu-rep (scripts/profile.sh) -> profiles/rNN_*.{md,csv}.
usage: python scripts/summarize_profiles.py r01"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def short(k):
    return k.replace("void ", "").replace("at::native::", "").replace("<unnamed>::", "")

# ---- launch list ------------------------------------------------------------------------------------
lines = [l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"')]
open(os.path.join(P, tag + "_launches.csv"), "w").writelines(lines)
r = csv.reader(lines); hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(short(row[ki]), float(row[vi].replace(",", "")) / 1000.0) for row in r]
idx = [i for i, (k, _) in enumerate(seq) if k.startswith("sample_coarse")]
step = seq[idx[-2]:idx[-1]]
agg = collections.OrderedDict()
for k, v in step:
    k = k.split("(")[0][:80]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in agg.values())
with open(os.path.join(P, tag + "_launches_summary.md"), "w") as f:
    f.write("# %s - ncu launch list of `python bench.py --steps 2 --warmup 3 --no-graph` (first 400 launches), ONE training step\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` on one B200 (scripts/profile.sh); times are cold-cache and "
            "serialised - compare SHARES.  The step shown is the last complete one in the capture (%d launches, %.1f us).\n\n" % (len(step), tot))
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, n, v, 100 * v / tot))

# ---- full capture --------------------------------------------------------------------------------------
def raw_table(stem):
    """raw metric table of a capture: the csv written on the GPU box, or converted here from the .ncu-rep"""
    c = os.path.join(G, stem + "_raw.csv")
    if os.path.isfile(c):
        return open(c).read()
    rep = os.path.join(G, stem + ".ncu-rep")
    if os.path.isfile(rep):
        return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return ""
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "second": 1e6}
tables = []
for stem, out in (("prof", "_ncu_full_raw.csv"), ("prof_render", "_ncu_render_raw.csv")):
    raw = raw_table(stem)
    if not raw:
        continue
    open(os.path.join(P, tag + out), "w").write(raw)
    t = list(csv.reader(raw.splitlines()))
    tables.append((t[0], t[1], t[2:]))
rows = []
for hdr, units, body in tables:
    col = {h: i for i, h in enumerate(hdr)}
    for r in body:
        rows.append((col, units, r))
def g(row, name):
    """metric value normalised to us / GB (percentages and counts unchanged)"""
    col, units, r = row
    if name not in col or r[col[name]] in ("", "n/a"):
        return float("nan")
    return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)
seen = collections.OrderedDict()
for r in rows:
    name = short(r[2][r[0]["Kernel Name"]]).split("(")[0]
    dur = g(r, "gpu__time_duration.sum")
    key = (name, round(g(r, "dram__bytes_read.sum") + g(r, "dram__bytes_write.sum"), 1))
    if key in seen:
        continue
    seen[key] = (name, dur, g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum"),
                 g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                 g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                 g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                 g(r, "launch__registers_per_thread"))
with open(os.path.join(P, tag + "_kernels_ncu.md"), "w") as f:
    f.write("# %s - `ncu --set full` of the kernels inside one bench step (B200, cfg 2: 4096 rays, P = 262,144 / 524,288 points) and of the render-only forward / compositing / sampling kernels at image-sized batches (scripts/time_kernels.py)\n\n" % tag)
    f.write("Command: `scripts/profile.sh`.  Raw metric dump: `%s_ncu_full_raw.csv`.  Durations under ncu are serialised; DRAM bytes are per launch "
            "(normalised to us and GB).\n\n" % tag)
    f.write("| kernel | duration us | dram read GB | dram write GB | dram % peak | tensor pipe active % | L2 % | smem wavefronts % | regs |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for v in seen.values():
        f.write("| `%s` | %.1f | %.3f | %.3f | %.1f | %.1f | %.1f | %.1f | %d |\n" % v)
# ---- DRAM traffic per launch of the MLP kernels inside one cfg-2 step: what bench.py reports as roofline.traffic ----
import json
traffic = {}
for r in rows:
    name = short(r[2][r[0]["Kernel Name"]]).split("(")[0].split("<")[0]
    if name not in ("mlp_forward_pair_kernel", "backward_fused_kernel", "head_grads_kernel"):
        continue
    b = (g(r, "dram__bytes_read.sum") + g(r, "dram__bytes_write.sum")) * 1e9
    if b != b or g(r, "gpu__time_duration.sum") > 3000:      # the render-sized launches of scripts/prof_hbm_stages.py are not part of a step
        continue
    traffic.setdefault(name, []).append(b)
out = {"source": "profiles/%s_ncu_full_raw.csv (ncu --set full, one cfg-2 step: coarse 262,144 + fine 524,288 points; mean of the two launches)" % tag,
       "rays_per_gpu": 4096, "kernels": {}}
for k, v in traffic.items():
    v = v[:2]
    out["kernels"][k] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)}
    label = {"mlp_forward_pair_kernel": "mvip_mlp_forward"}.get(k)
    if label:
        out["kernels"][label] = out["kernels"][k]
json.dump(out, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print("wrote profiles/%s_* and profiles/ncu_traffic.json" % tag)
