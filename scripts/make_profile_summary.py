"""Builds profiles/r1_summary.md (+ copies the raw artefacts) from the files the round-1 measurement pass left in
gpurun_out/ (scripts/r1_final_runs.sh)."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = ["# Round 1 — measured on the pool's B200 (1 GPU unless stated)\n",
       "Produced by `scripts/r1_final_runs.sh` + `scripts/make_profile_summary.py`; raw artefacts alongside.\n"]


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return None
    try:
        return json.loads(open(p).readline())
    except Exception:
        return None


for name, title in (("r1_bench_clip.json", "CLIP ViT-B/32 (headline, BASELINE configs[1])"),
                    ("r1_bench_clip_reference.json", "CLIP reference arm (`--impl reference`: oracle port of the `--cpu` flow)"),
                    ("r1_bench_i3d.json", "I3D rgb (configs[2])"), ("r1_bench_raft.json", "RAFT -> I3D flow (configs[3])")):
    d = load(name)
    if not d:
        out.append(f"## {title}\n\n_missing_\n")
        continue
    shutil.copy(os.path.join(G, name), os.path.join(P, name))
    out.append(f"## {title}\n")
    out.append(f"* `{d['metric']}`: **{d['value']:.1f} {d['unit']}** ({d['ms_per_step']:.3f} ms/step, {d['steps']} steps, warm-up {d['warmup']})")
    if "e2e" in d:
        out.append(f"* end to end (host buffers in the timed region): {d['e2e']['value']:.1f} {d['e2e'].get('unit', d['unit'])}"
                   f" (H2D {d['e2e'].get('h2d_bytes_per_step', 0)/1e6:.1f} MB, D2H {d['e2e'].get('d2h_bytes_per_step', 0)/1e6:.2f} MB per step)")
    r = d.get("roofline")
    if r:
        out.append(f"* roofline ({r['bound']}): achieved {r['achieved']:.1f} {r['unit']} of {r['peak']:.1f} ({r['peak_source']}) = **{r['frac']:.3f}**;"
                   f" executed {r.get('executed_tflops', float('nan')):.1f} TF/s; GEMM share of step {r.get('gemm_share_of_step', float('nan')):.2f}")
        if "eager_ms_per_step_by_kernel" in r:
            out.append(f"* eager per-kernel-class device ms per step: {json.dumps({k: round(v, 3) for k, v in r['eager_ms_per_step_by_kernel'].items()})}")
    c = d.get("cpu_baseline")
    if c:
        out.append(f"* CPU baseline ({c['kind']}, {c['cores']} cores): {c['value']:.2f} {c['unit']} — {c['sample']}")
        out.append(f"* GPU e2e / CPU: {d.get('e2e', {}).get('value', d['value']) / c['value']:.0f}x")
    if d.get("clocks"):
        out.append(f"* clocks under load: {json.dumps(d['clocks'])}")
    out.append(f"* launches in timed region: {d.get('gpu_launches')}\n")

# ---- 2-GPU line (gpurun --gpus 2, torchrun); kept in profiles/ between passes
p2 = os.path.join(G, "r1_bench_clip_2gpu.json")
if os.path.exists(p2):
    shutil.copy(p2, os.path.join(P, "r1_bench_clip_2gpu.json"))
try:
    d2 = json.loads([l for l in open(os.path.join(P, "r1_bench_clip_2gpu.json")) if l.startswith("{")][-1])
    d1 = load("r1_bench_clip.json")
    out.append("## CLIP, 2 x B200 (`gpurun --gpus 2`, torchrun, NCCL all_gather of the (2000,512) features inside the timed region)\n")
    out.append(f"* **{d2['value']:.1f} frames/s** whole job ({d2['ms_per_step']:.3f} ms/step max over ranks; weak scaling: 1000 frames per rank per step),"
               f" e2e {d2['e2e']['value']:.1f} frames/s" + (f"; vs 1 GPU ({d1['value']:.0f} frames/s): x{d2['value']/d1['value']:.2f}" if d1 else "") + "\n")
except Exception:
    pass
p4 = os.path.join(G, "r1_bench_clip_4gpu.json")
if os.path.exists(p4):
    shutil.copy(p4, os.path.join(P, "r1_bench_clip_4gpu.json"))
try:
    d4 = json.loads([l for l in open(os.path.join(P, "r1_bench_clip_4gpu.json")) if l.startswith("{")][-1])
    d1 = load("r1_bench_clip.json")
    out.append(f"* 4 x B200 (`gpurun --gpus 4`): **{d4['value']:.1f} frames/s** ({d4['ms_per_step']:.3f} ms/step), e2e {d4['e2e']['value']:.1f} frames/s"
               + (f"; x{d4['value']/d1['value']:.2f} of 1 GPU" if d1 else "") + "\n")
except Exception:
    pass

# ---- ncu launch list
lp = os.path.join(G, "r1_launches_clip.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, "r1_launches_clip.csv"))
    rows = list(csv.DictReader([l for l in open(lp) if not l.startswith("==")]))
    agg = collections.OrderedDict()
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("vf::<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out.append("## ncu launch list of `bench.py --steps 2 --warmup 1 --no-cpu` (cold-cache, serialised: compare SHARES)\n")
    out.append(f"{len(rows)} launches captured (`-s 1092 -c 728`, ≈ the two timed steps), total {tot/1e6:.3f} ms\n")
    out.append("| kernel | launches | total µs | avg µs | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {v[1]/v[0]/1e3:.1f} | {100*v[1]/tot:.1f} % |")
    out.append("")

# ---- ncu --set full of the GEMM
rep = os.path.join(G, "r1_prof_gemm.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, "r1_prof_gemm_raw.csv"), "w").write(raw)
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) > 2:
        hdr, units, data = rr[0], rr[1], rr[2:]
        want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
                "lts__t_sector_hit_rate.pct"]
        out.append("## ncu `--set full` of the tcgen05 GEMM (4 consecutive launches of a 250-frame chunk: out-proj, fc1, fc2, QKV)\n")
        out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data))
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {units[i]} | " + " | ".join(d[i][:28] for d in data) + " |")
        out.append("")
        try:
            i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            out.append("DRAM traffic per launch (read + write, units as in the table): "
                       + ", ".join(f"{float(d[i_r]) + float(d[i_w]) if units[i_r] == units[i_w] else d[i_r] + '+' + d[i_w]}" for d in data) + "\n")
        except Exception:
            pass

# ---- I3D: per-launch list of one forward (8 clips x 64 frames)
lp = os.path.join(G, "r1_launches_i3d.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, "r1_launches_i3d.csv"))
    per = collections.OrderedDict()
    for x in csv.DictReader([l for l in open(lp) if not l.startswith("==")]):
        d = per.setdefault(x["ID"], {"name": re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("vf::<unnamed>::", "")})
        d[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    L = list(per.values())
    first = [i for i, x in enumerate(L) if "phase_pack" in x["name"]]
    if len(first) >= 3:
        S = L[first[1]:first[2]]          # the second forward: pack .. head
        tot = sum(x["gpu__time_duration.sum"] for x in S)
        md = ["# I3D rgb forward, 8 clips x 64 frames: ncu per-launch list (cold-cache, serialised; one forward)\n",
              f"{len(S)} launches, {tot/1e6:.3f} ms\n",
              "| # | kernel | µs | tensor pipe % | DRAM read MB | DRAM write MB | L2 MB | grid |\n|---|---|---|---|---|---|---|---|"]
        for i, x in enumerate(S):
            md.append(f"| {i} | `{x['name'][:40]}` | {x['gpu__time_duration.sum']/1e3:.1f} | "
                      f"{x.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                      f"{x.get('dram__bytes_read.sum', 0)/1e6:.1f} | {x.get('dram__bytes_write.sum', 0)/1e6:.1f} | "
                      f"{x.get('lts__t_bytes.sum', 0)/1e6:.0f} | {int(x.get('launch__grid_size', 0))} |")
        open(os.path.join(P, "r1_i3d_launches.md"), "w").write("\n".join(md) + "\n")
        agg = collections.OrderedDict()
        for x in S:
            a = agg.setdefault(x["name"], [0, 0.0]); a[0] += 1; a[1] += x["gpu__time_duration.sum"]
        out.append("## I3D forward (8 clips x 64 frames): ncu kernel shares of one forward (full list: `r1_i3d_launches.md`)\n")
        out.append("| kernel | launches | total µs | share |\n|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f} % |")
        stem = next((x for x in S if "gemm" in x["name"]), None)
        if stem:
            out.append(f"\nstem GEMM (first GEMM launch): {stem['gpu__time_duration.sum']/1e3:.1f} µs = {100*stem['gpu__time_duration.sum']/tot:.1f} % of the forward, "
                       f"tensor pipe {stem.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} %, L2 traffic {stem.get('lts__t_bytes.sum', 0)/1e9:.2f} GB\n")

# ---- RAFT: kernel shares of one call (8 pairs 270x480, 20 iterations)
lp = os.path.join(G, "r1_launches_raft.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, "r1_launches_raft.csv"))
    rows = list(csv.DictReader([l for l in open(lp) if not l.startswith("==")]))
    rows = rows[len(rows) // 2:]           # the second of the two calls
    agg = collections.OrderedDict()
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("vf::<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(x["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out.append("## RAFT, one call of 8 pairs 270x480 x 20 iterations: ncu kernel shares (cold-cache, serialised)\n")
    out.append(f"{len(rows)} launches, {tot/1e6:.3f} ms\n")
    out.append("| kernel | launches | total µs | avg µs | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {v[1]/v[0]/1e3:.1f} | {100*v[1]/tot:.1f} % |")
    out.append("")

# ---- ncu --set full of other kernels captured during the round (attention, epilogue-only GEMMs)
for rep_name, title in (("prof_ln", "add_layernorm768_kernel (HBM-bound; kernel unchanged since this capture)"),
                        ("r1_prof_attn", "attention50_kernel (250-frame chunk, 3000 blocks)"),
                        ("r1_prof_epi", "epilogue-dominated GEMMs: K = 64, M = 12000, N = 3072, fp16 out; launch 0 = +bias, launch 1 = +bias+QuickGELU")):
    rep = os.path.join(G, rep_name + ".ncu-rep")
    csvp = os.path.join(P, rep_name + "_raw.csv")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        if raw.strip():
            open(csvp, "w").write(raw)
    if os.path.exists(csvp):
        rr = list(csv.reader(open(csvp).read().splitlines()))
        if len(rr) > 2:
            hdr, units, data = rr[0], rr[1], rr[2:]
            want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                    "launch__registers_per_thread", "launch__grid_size"]
            out.append(f"## ncu `--set full`: {title}\n")
            out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data))
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"| {w} | {units[i]} | " + " | ".join(d[i][:14] for d in data) + " |")
            out.append("")

# ---- library-call bar: the oracle's torch modules on the same GPU (bench.py --torch-gpu)
rows = []
for w, title in (("clip", "CLIP ViT-B/32 tower"), ("i3d", "I3D rgb"), ("raft", "RAFT -> I3D flow")):
    src = os.path.join(G, f"r1_torchgpu_{w}.json")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"r1_torchgpu_{w}.json"))
    try:
        d = json.loads(open(os.path.join(P, f"r1_torchgpu_{w}.json")).readline())
        b = d["torch_gpu_baseline"]
        rows.append(f"| {title} | {d['value']:.0f} {d['unit']} | {b['fp32']:.1f} | x{d['value']/b['fp32']:.1f} | {b['tf32']:.1f} | x{d['value']/b['tf32']:.1f} | {b['what']} |")
    except Exception:
        pass
if rows:
    out.append("## Library-call bar: the fp32 torch modules of the oracle, eager, on the same B200 (`bench.py --torch-gpu`)\n")
    out.append("| workload | this engine | torch fp32 (TF32 off) | ratio | torch TF32 | ratio | what |\n|---|---|---|---|---|---|---|")
    out += rows
    out.append("\n(The reference runs I3D and RAFT in fp32 and CLIP in fp16 on CUDA; a torch fp16 tower was not measured.)\n")

t = os.path.join(G, "r1_pytest_gpu.txt")
if os.path.exists(t):
    out.append("## `pytest tests -m gpu` on the same box\n\n```\n" + open(t).read().strip() + "\n```\n")
open(os.path.join(P, "r1_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
