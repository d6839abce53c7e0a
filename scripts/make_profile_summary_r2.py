"""Builds profiles/r2_summary.md (+ copies the raw artefacts) from what scripts/r2_final_runs.sh left in gpurun_out/."""
import collections, csv, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
out = ["# Round 2 — measured on the pool's B200s\n",
       "Produced by `scripts/r2_final_runs.sh` + `scripts/make_profile_summary_r2.py`; raw artefacts alongside "
       "(`r2_*.json`, `r2_launches_clip.csv`, `r2_prof_*_raw.csv`, `r2/`).  Kernel-level analysis: `r2_gemm_notes.md`.\n"]


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        p = os.path.join(P, name)
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


def keep(name):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, name))


def describe(d, title):
    out.append(f"## {title}\n")
    out.append(f"* `{d['metric']}`: **{d['value']:.1f} {d['unit']}** ({d['ms_per_step']:.3f} ms/step, {d['steps']} steps, warm-up {d['warmup']}, {d['n_gpus']} GPU)")
    if "e2e" in d:
        out.append(f"* end to end (host buffers in the timed region): {d['e2e']['value']:.1f} {d['e2e'].get('unit', d['unit'])}"
                   f" (H2D {d['e2e'].get('h2d_bytes_per_step', 0)/1e6:.1f} MB, D2H {d['e2e'].get('d2h_bytes_per_step', 0)/1e6:.2f} MB per step)")
    r = d.get("roofline")
    if r:
        out.append(f"* roofline ({r['bound']}, {r.get('kernel', '')}): achieved {r['achieved']:.1f} {r['unit']} of {r['peak']:.1f} "
                   f"({r['peak_source']}) = **{r['frac']:.3f}**; executed/algorithmic {r.get('executed_over_algorithmic', float('nan')):.2f}; "
                   f"share of step {r.get('gemm_share_of_step', float('nan')):.2f}")
        if r.get("qkv_attention"):
            q = r["qkv_attention"]
            out.append(f"* fused QKV + attention kernel: {q['ms_per_step']:.3f} ms/step, {q['achieved_tflops']:.1f} TFLOP/s = {q['frac']:.3f} of the same peak")
        if "eager_ms_per_step_by_kernel" in r:
            out.append(f"* eager per-kernel-class device ms per step: {json.dumps({k: round(v, 3) for k, v in r['eager_ms_per_step_by_kernel'].items()})}")
        for k, v in (r.get("hbm") or {}).items():
            if v.get("achieved_gbs"):
                out.append(f"* HBM roofline, {k}: {v['achieved_gbs']:.0f} GB/s of {r.get('hbm_peak_gbs', 0):.0f} = {v['frac_of_hbm_peak']:.2f}"
                           + (f" ({v['frames_per_sec']:.0f} frames/s)" if "frames_per_sec" in v else ""))
        if "whole_step_frac" in r:
            out.append(f"* whole step: {r['whole_step_tflops']:.0f} TFLOP/s = {r['whole_step_frac']:.3f} of the tensor peak")
    c = d.get("cpu_baseline")
    if c:
        out.append(f"* CPU baseline ({c['kind']}, {c['cores']} cores): {c['value']:.2f} {c['unit']} — {c['sample']}")
        out.append(f"* GPU e2e / CPU: {d.get('e2e', {}).get('value', d['value']) / c['value']:.0f}x")
    t = d.get("torch_gpu_baseline")
    if t:
        out.append(f"* library bar (torch eager on the same GPU, {t.get('what', '')}): " +
                   ", ".join(f"{k} {t[k]:.0f} {t['unit']}" for k in ("fp32", "tf32", "fp16") if k in t))
    if d.get("clocks"):
        out.append(f"* clocks under load: {json.dumps(d['clocks'])}")
    out.append(f"* launches in timed region: {d.get('gpu_launches')}\n")


for name, title in (("r2_bench_clip.json", "CLIP ViT-B/32 (headline, BASELINE configs[1])"),
                    ("r2_bench_clip_reference.json", "CLIP reference arm (`--impl reference`: oracle port of the `--cpu` flow)"),
                    ("r2_bench_clip_torchgpu.json", "CLIP with the torch-on-GPU library bar (`--torch-gpu`)")):
    d = load(name)
    if not d:
        out.append(f"## {title}\n\n_missing_\n")
        continue
    keep(name)
    describe(d, title)
    for k, v in (d.get("secondary") or {}).items():
        if "error" in v and "value" not in v:
            out.append(f"### secondary `{k}`: {v['error']}\n")
        else:
            describe(v, f"secondary `{k}` (same bench.py run)")

for name, title in (("r2l_bench_2gpu.json", "2 x B200 (`gpurun --gpus 2`): headline + 10k-video list + RAFT -> I3D flow (BASELINE configs[3] as named: 2 GPUs)"),):
    p = os.path.join(G, name.replace(".json", ".err"))
    d = load(name)
    if d is None and os.path.exists(p):          # the self-launched run of that call printed its line on stderr
        for l in open(p):
            if l.startswith("{"):
                d = json.loads(l)
    if d:
        json.dump(d, open(os.path.join(P, "r2_bench_clip_2gpu.json"), "w"))
        describe(d, title)
        for k, v in (d.get("secondary") or {}).items():
            if "value" in v:
                describe(v, f"secondary `{k}` at 2 GPUs")

# ---- later multi-GPU runs on the final code (the 2-GPU record above predates the asynchronous entry point)
d = load("r2_c5_2gpu.json") or load("r2t_c5_2gpu.json")
if d:
    describe(d, "10k-video list at 2 GPUs, final code (`torchrun --nproc-per-node 2 bench.py --gpus 2 --workload c5`)")
d = load("r2/bench_8gpu_run2.json")
if d:
    d = dict(d)
    d.pop("secondary", None)
    describe(d, "8 x B200, final code (`torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 30 --warmup 5`; secondary lines: see r2/README.md)")

# ---- ncu launch list
src = os.path.join(G, "r2_launches_clip.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, "r2_launches_clip.csv"))
try:
    rows = [r for r in csv.reader(open(os.path.join(P, "r2_launches_clip.csv"))) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = rows[1:]
    half = data[len(data) // 2:]                 # second call of the script = warm
    agg = collections.OrderedDict()
    for r in half:
        k = r[ik].split("(")[0].replace("void ", "").replace("vf::<unnamed>::", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    out.append("## ncu launch list of one 1000-frame call (`scripts/ncu_clip_once.py 1000 250`; cold-cache, serialised: compare SHARES)\n")
    out.append(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms\n")
    out.append("| kernel | launches | total µs | avg µs | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |")
    out.append("")
except Exception as e:
    out.append(f"_launch list unavailable: {e}_\n")

# ---- ncu --set full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "lts__t_sector_hit_rate.pct"]
for tag, title in (("gemm", "the tcgen05 GEMM (4 consecutive launches of a 250-frame chunk: fc1+GELU, fc2, out-proj, fc1+GELU)"),
                   ("attn", "the fused QKV + attention kernel (250 frames)"), ("ln", "LayerNorm (12500 rows)"),
                   ("resample", "Pillow-exact resample, 256 frames 240x320 -> 224x298 bicubic (horizontal, vertical pass)")):
    p = os.path.join(P, f"r2_prof_{tag}_raw.csv")
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(open(p)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out.append(f"## ncu `--set full` of {title}\n")
    out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    out.append("|---|---|" + "---|" * len(data))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {units[i]} | " + " | ".join(d[i] for d in data) + " |")
    out.append("")

open(os.path.join(P, "r2_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
