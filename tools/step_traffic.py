"""One training step under `ncu --profile-from-start off` (metrics pass, not a timing run):
    ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --csv --log-file gpurun_out/step_traffic.csv python tools/step_traffic.py
then   python tools/step_traffic.py --summarise gpurun_out/step_traffic.csv profiles/r01_step_traffic.json
The summary (per-kernel launches, cold-cache serialised time share, DRAM bytes) feeds bench.py's roofline.traffic."""
import sys, json, csv, collections
sys.path.insert(0, ".")


def run():
    import torch
    import bench
    from transformers import BatchEncoding
    from ctpa_clip_b200.trainer import CTClipTrainStep
    from ctpa_clip_b200 import configs as O
    cfg = O.CONFIGS["production"]
    dev = torch.device("cuda", 0)
    model = O.build_model(cfg, dev, seed=0)
    trainer = CTClipTrainStep(model)
    video_h, ids, mask = O.synth_batch(cfg, 8, seed=100)
    video = video_h.to(dev)
    text = BatchEncoding({"input_ids": ids.to(dev), "attention_mask": mask.to(dev)})
    for _ in range(2):
        trainer.step(text, video)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    trainer.step(text, video)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


def summarise(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mn, mv, idc, mu = (h.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3,
             "Mbyte": 1e6, "Gbyte": 1e9}
    per = collections.defaultdict(dict)
    name = {}
    for r in rows[hdr + 1:]:
        if len(r) <= mv or not r[idc].isdigit():
            continue
        per[int(r[idc])][r[mn]] = float(r[mv].replace(",", "")) * scale[r[mu]]
        name[int(r[idc])] = r[kn]
    import re
    agg = collections.OrderedDict()
    for i in sorted(per):
        short = re.sub(r"\(.*", "", name[i]).replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(short, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        m = per[i]
        a["launches"] += 1
        a["time_us"] += m.get("gpu__time_duration.sum", 0.0)
        a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a["time_us"] for a in agg.values())
    for a in agg.values():
        a["share_of_step"] = a["time_us"] / total
    gem = [a for k, a in agg.items() if k.startswith("gemm_bf16_kernel")]
    out = {"source": "ncu metrics pass over ONE training step (tools/step_traffic.py), cold-cache serialised launches",
           "total_time_us": total, "launches": sum(a["launches"] for a in agg.values()),
           "gemm": {"launches": sum(a["launches"] for a in gem), "time_us": sum(a["time_us"] for a in gem),
                    "share_of_step": sum(a["time_us"] for a in gem) / total,
                    "dram_bytes": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in gem)},
           "kernels": dict(sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]))}
    out["gemm"]["dram_bytes_per_launch"] = out["gemm"]["dram_bytes"] / max(1, out["gemm"]["launches"])
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["gemm"]), "total_us", total)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2], sys.argv[3])
    else:
        run()
