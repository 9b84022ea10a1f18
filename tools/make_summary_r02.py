"""Refreshes the measured lines of profiles/r02_summary.md from the committed logs (bench lines of the final code at N = 1 / 8,
the per-op breakdown); the prose and the kernel-timing table of that file are maintained by hand.
    python tools/make_summary_r02.py"""
import json
import re

P = "profiles/"
line = lambda f: json.loads([l for l in open(P + f) if l.startswith("{")][0])
b, b8 = line("r02_bench_final.log"), line("r02_bench_8gpu_c.log")
rows = dict(json.load(open(P + "r02_breakdown_final.json"))["ops"])
tot = sum(v["ms"] for v in rows.values())
g = lambda pred: sum(v["ms"] for k, v in rows.items() if pred(k))
groups = [
    ("tcgen05 GEMMs (CTViT + VQ + latent projections)", g(lambda k: k.startswith("gemm") and "bert" not in k)),
    ("tcgen05 GEMMs (BERT, incl. batched per-head products)", g(lambda k: k.startswith("gemm:bert"))),
    ("CTViT attention fwd+bwd (spatial tcgen05 + temporal mma.sync)", g(lambda k: k.startswith("attn"))),
    ("BERT fused attention fwd+bwd", g(lambda k: k.startswith("bert_attn"))),
    ("LayerNorm fwd+bwd (CTViT + BERT)", g(lambda k: k.startswith("layernorm"))),
    ("GEGLU backward (the forward lives in the FF1 epilogue)", g(lambda k: k.startswith("geglu"))),
    ("PEG fwd + data grad + weight grad", g(lambda k: k.startswith("peg"))),
    ("BERT GELU / dropout / embeddings / column sums",
     g(lambda k: (k.startswith("bert_") and not k.startswith("bert_attn")) or k in ("gelu_fwd", "gelu_bwd", "dropout_add", "colsum_bf16", "colsum"))),
    ("clip-norm + Adam", g(lambda k: k in ("adam_step", "sumsq"))),
]
groups.append(("other (VQ, CPB, patch LN, casts, loss, ...)", tot - sum(v for _, v in groups)))

s = open(P + "r02_summary.md").read()
e, r = b["e2e"], b["roofline"]
out, in_table = [], False
for l in s.split("\n"):
    if l.startswith("* device-resident step:"):
        l = re.sub(r"\*\*[\d.]+ ms = [\d.]+ volumes/s\*\*", f"**{b['ms_per_step']:.2f} ms = {b['value']:.1f} volumes/s**", l)
    elif l.startswith("* whole step:"):
        l = f"* whole step: {b['achieved_tflops_step']:.0f} TFLOP/s algorithmic"
    elif l.startswith("| 1 |"):
        l = f"| 1 | {b['value']:.1f} | {b['ms_per_step']:.2f} | 1 | {e['value']:.1f} | `r02_bench_final.log` |"
    elif l.startswith("## Where the step goes"):
        l = f"## Where the step goes (`r02_breakdown_final.json`, per-op CUDA events of one pre-fed instrumented step, sum {tot:.1f} ms)"
        in_table = True
    elif in_table and l.startswith("| ") and not l.startswith("| group") and not l.startswith("|---"):
        continue                       # old rows of the group table: re-emitted below
    elif in_table and l.startswith("|---"):
        out.append(l)
        out.extend(f"| {n} | {v:.2f} | {100 * v / tot:.1f} % |" for n, v in groups)
        continue
    elif in_table and l.startswith("## "):
        in_table = False
    out.append(l)
open(P + "r02_summary.md", "w").write("\n".join(out))
print(f"N=1 {b['value']:.1f} volumes/s ({b['ms_per_step']:.2f} ms), e2e {e['mode']} {e['value']:.1f}; N=8 {b8['value']:.1f}, e2e {b8['e2e']['value']:.1f}; "
      f"GEMM roofline {r['frac']:.3f}; breakdown sum {tot:.1f} ms")
