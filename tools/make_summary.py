"""Regenerates profiles/r01_summary.md from the committed bench / breakdown / component JSON files of a version tag:
    python tools/make_summary.py v4"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "v4"
d = json.load(open(f"profiles/r01_step_breakdown_{tag}.json"))
import os
ctag = next(t for t in (tag, "v8", "v7") if os.path.exists(f"profiles/r01_components_{t}.json"))   # component bench: last run
c = json.load(open(f"profiles/r01_components_{ctag}.json"))
bench = json.loads([l for l in open(f"profiles/r01_bench_{tag}.log") if l.startswith("{")][-1])
rows = d["ops"]
tot = sum(v["ms"] for k, v in rows)
grp = lambda pred: sum(v["ms"] for k, v in rows if pred(k))
g = {"tcgen05 GEMMs (CTViT + VQ + latent projections)": grp(lambda k: k.startswith("gemm") and "bert" not in k),
     "tcgen05 GEMMs (BERT, incl. batched per-head products)": grp(lambda k: k.startswith("gemm:bert")),
     "CTViT attention fwd+bwd (spatial tcgen05 + temporal mma.sync)": grp(lambda k: k.startswith("attn")),
     "BERT fused attention fwd+bwd": grp(lambda k: k.startswith("bert_attn")),
     "LayerNorm fwd+bwd (CTViT + BERT)": grp(lambda k: k.startswith("layernorm")),
     "GEGLU fwd+bwd": grp(lambda k: k.startswith("geglu")),
     "PEG fwd + data grad + weight grad": grp(lambda k: k.startswith("peg")),
     "BERT softmax / GELU / dropout / embeddings / bf16 colsum": grp(lambda k: (k.startswith("bert_") and not k.startswith("bert_attn")) or k in ("gelu_fwd", "gelu_bwd", "dropout_add", "colsum_bf16")),
     "clip-norm + Adam": grp(lambda k: k in ("adam_step", "sumsq"))}
other = tot - sum(g.values())
L = []
L.append("# Round 1 — measurement summary (1×B200 unless noted; B = 8 volumes 480×480×240 + 8 reports × 512 ids per rank)\n")
L.append("All numbers are CUDA-event timings from `bench.py` / `tools/bench_components.py` runs on the GPU box (no profiler attached);")
L.append("ncu evidence: `r01_ncu_launches_v8.txt` + `r01_step_traffic.json` (launch list and DRAM bytes of one training step), `r01_ncu_kernels_v3.txt` / `r01_ncu_kernels_v7.txt` (`--set full` extracts).\n")
L.append(f"## Headline (`r01_bench_{tag}.log`)\n")
L.append(f"* device-resident step: **{bench['ms_per_step']:.1f} ms = {bench['value']:.1f} volumes/s** (session start: 79.2 ms / 101 volumes/s with the text tower still on torch; previous snapshots v4: 58.8 ms, v7: 52.1 ms, v8: 51.5 ms)")
L.append(f"* end to end (pinned host volumes → H2D on a copy stream → step → loss copied back to pinned host memory every step, consumed one step later): **{bench['e2e']['value']:.1f} volumes/s** ({bench['e2e']['ms_per_step']:.1f} ms/step, {bench['e2e']['h2d_bytes_per_step'] / 1e9:.2f} GB H2D per step)")
L.append(f"* all `gemm_bf16_kernel` launches of the step: {bench['roofline']['achieved']:.0f} TFLOP/s = {100 * bench['roofline']['frac']:.0f} % of the measured sustained cuBLAS bf16 rate, {100 * bench['roofline']['share_of_step']:.0f} % of the step")
try:
    b2 = json.loads([l for l in open(next(f for f in (f"profiles/r01_bench_2gpu_{tag}.log", "profiles/r01_bench_2gpu_v8.log") if os.path.exists(f))) if l.startswith("{")][-1])
    L.append(f"* 2×B200 (`r01_bench_2gpu_{tag}.log` if present, else `r01_bench_2gpu_v8.log`; weak scaling, global batch 16, latents exchanged by the fused peer-memory loss kernel (NCCL arm: `r01_bench_2gpu_v8_nccl.log`), gradients all-reduced over NVLink, the text-tower / latent-projection part overlapped with the image tower's backward): {b2['ms_per_step']:.1f} ms/step = {b2['value']:.0f} volumes/s ({100 * b2['value'] / (2 * bench['value']):.0f} % of 2× the 1-GPU step)")
except (FileNotFoundError, StopIteration):
    pass
for n in (4, 8):
    try:
        bn = json.loads([l for l in open(next(f for f in (f"profiles/r01_bench_{n}gpu_{tag}.log", f"profiles/r01_bench_{n}gpu_v8.log") if os.path.exists(f))) if l.startswith("{")][-1])
        L.append(f"* {n}×B200 (`r01_bench_{n}gpu_{tag}.log` if present, else `..._v8.log`; weak scaling, global batch {8 * n}): {bn['ms_per_step']:.1f} ms/step = {bn['value']:.0f} volumes/s ({100 * bn['value'] / (n * bench['value']):.0f} % of {n}× the 1-GPU step), end to end {bn['e2e']['value']:.0f} volumes/s")
    except (FileNotFoundError, StopIteration):
        pass
for rtag in (tag, "v7"):   # the reference arm does not depend on our kernels: the last measured run is reused
    try:
        br = json.loads([l for l in open(f"profiles/r01_bench_reference_arm_{rtag}.log") if l.startswith("{")][-1])
        L.append(f"* `bench.py --impl reference` (the reference's CPU path, oracle port on all host threads; `r01_bench_reference_arm_{rtag}.log`): {br['value']:.2f} volumes/s")
        break
    except FileNotFoundError:
        pass
L.append(f"* CPU oracle (fp32 port of the reference, {bench['cpu_baseline']['cores']} host threads): {bench['cpu_baseline']['value']:.2f} volumes/s\n")
L.append("## Where the step goes (`r01_step_breakdown_%s.json`, one instrumented step, Σ = %.1f ms of the %.1f ms step)\n" % (tag, tot, d["ms_per_step"]))
L.append("| group | ms | share |\n|---|---|---|")
for k, v in sorted(g.items(), key=lambda kv: -kv[1]):
    L.append(f"| {k} | {v:.2f} | {100 * v / tot:.1f} % |")
L.append(f"| everything else (patch LN, VQ finalize / gather / EMA, casts, fp32 colsum, loss, pooling) | {other:.2f} | {100 * other / tot:.1f} % |\n")
L.append("Top single entries:\n")
L.append("| op (shape MxNxK for GEMMs) | ms | launches | TFLOP/s |\n|---|---|---|---|")
for k, v in rows[:24]:
    tf = f"{v['flops'] / v['ms'] / 1e9:.0f}" if v["flops"] else ""
    L.append(f"| {k} | {v['ms']:.2f} | {v['n']} | {tf} |")
L.append(f"\n## Kernel rooflines on production shapes (`r01_components_{ctag}.json`)\n")
L.append("| kernel / shape | time | achieved | of measured peak |\n|---|---|---|---|")
for k, v in c.items():
    if "TFLOPs" in v:
        L.append(f"| {k} | {v['ms'] * 1e3:.0f} µs | {v['TFLOPs']:.0f} TFLOP/s | {100 * v['frac_of_measured_bf16_burst']:.0f} % of 1639 TFLOP/s (cuBLAS burst) |")
    elif "algorithmic_GBps" in v:
        fr = v.get("frac_of_measured_hbm", v["algorithmic_GBps"] / 6556.2)
        L.append(f"| {k} | {v['ms'] * 1e3:.0f} µs | {v['algorithmic_GBps']:.0f} GB/s algorithmic | {100 * fr:.0f} % of 6556 GB/s (copy bandwidth) |")
    elif "volumes_per_s" in v and "ms" in v:
        L.append(f"| {k} | {v['ms']:.1f} ms | {v['volumes_per_s']:.0f} volumes/s | |")
    elif "volumes_per_s" in v:
        L.append(f"| {k} | {v.get('s_per_volume', 0):.2f} s/volume | {v['volumes_per_s']:.2f} volumes/s | {v.get('kind', '')} |")
L.append("")
L.append(open("profiles/r01_bounds.md").read())
open("profiles/r01_summary.md", "w").write("\n".join(L) + "\n")
print("\n".join(L)[:1800])
