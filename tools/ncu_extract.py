"""One line per captured launch from `ncu --set full` reports (read with `ncu -i REP --page raw --csv`):

    python tools/ncu_extract.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/rNN_ncu_kernels.txt

Numbers under a profiler are never bench values: the columns say what bounds a kernel (pipe utilisation, DRAM bytes,
stall reasons per issue-active cycle), the timings come from CUDA events in bench.py / tools/time_*.py."""
from __future__ import annotations

import csv
import io
import subprocess
import sys

COLS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("rd_MB", "dram__bytes_read.sum"),
    ("wr_MB", "dram__bytes_write.sum"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_%", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("regs", "launch__registers_per_thread"),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("st_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("st_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("grid", "launch__grid_size"),
]
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def rows(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    head, units, body = r[0], r[1], r[2:]
    idx = {}
    for label, metric in COLS:
        hit = [i for i, n in enumerate(head) if n == metric or n.endswith("." + metric)]
        idx[label] = hit[0] if hit else None
    kname = head.index("Kernel Name")
    for b in body:
        vals = []
        for label, _ in COLS:
            i = idx[label]
            try:
                v = float(b[i].replace(",", "")) if i is not None else float("nan")
            except ValueError:
                v = float("nan")
            u = units[i]
            if label in ("rd_MB", "wr_MB", "time_us") and u in SCALE:
                v *= SCALE[u]
            if label == "inst_M":
                v *= 1e-6
            vals.append(v)
        yield b[kname], vals


def main() -> None:
    print("# ncu --set full --clock-control none --import-source on extracts; one row per captured launch "
          "(tools/ncu_extract.py). st_* = average warps stalled per issue-active cycle, by reason.")
    print("# columns: " + ", ".join(c for c, _ in COLS))
    for rep in sys.argv[1:]:
        print(f"\n## {rep.split('/')[-1]}")
        for name, vals in rows(rep):
            short = name.replace("<unnamed>::", "").replace("void ", "")[:60]
            print(f"{short:60s} " + " ".join(f"{v:9.1f}" for v in vals))


if __name__ == "__main__":
    main()
