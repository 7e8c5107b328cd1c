#!/usr/bin/env python
"""profiles/traffic.json from ncu captures of the feature kernels: DRAM bytes read + written per launch, tied to the
build by the hash of the kernel sources (bench.py reports `roofline.traffic` from it and says whether the hash still
matches the build being timed).

    python tools/ncu_traffic.py workload=path/to/capture.ncu-rep [...]"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import source_sha  # noqa: E402


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        v = float(first[i].replace(",", ""))
        tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return int(tot), first[hdr.index("Kernel Name")]


def main():
    out_path = ROOT / "profiles" / "traffic.json"
    out = {}
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    for arg in sys.argv[1:]:
        wname, rep = arg.split("=", 1)
        b, kernel = dram_bytes(rep)
        out[wname] = {"bytes_per_launch": b, "kernel": kernel, "source_sha": source_sha(), "git": git,
                      "capture": f"ncu --set full --clock-control none -k regex:features -c 1 python tools/kbench.py 1 ... ({Path(rep).name})"}
    out["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum of one launch of the feature kernel on the workload's batch"
    out_path.write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
