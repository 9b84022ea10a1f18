"""Builds the C restatement (oracle/resample_oracle.c) into oracle/_build/libctclip_oracle.so with gcc.
Test infrastructure only; building the checker is not using it."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "libctclip_oracle.so"


def build(force: bool = False) -> Path:
    src = HERE / "resample_oracle.c"
    if force or not OUT.exists() or OUT.stat().st_mtime < src.stat().st_mtime:
        OUT.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(OUT), str(src), "-lm"],
                       check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
