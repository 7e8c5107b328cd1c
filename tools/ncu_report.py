"""Turn one `ncu --set full --import-source on` capture into the text summary kept under profiles/.

    python tools/ncu_report.py gpurun_out/x/feat.ncu-rep "header line" > profiles/rNN_ncu_<kernel>.txt

Key raw metrics of the first captured launch, then the per-opcode / per-region source summary
(tools/ncu_source_summary.py)."""
import csv
import subprocess
import sys
import tempfile
from pathlib import Path

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else rep}")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    kname = first[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# kernel: {kname}\n\n## key raw metrics (first captured launch)")
    for i, h in enumerate(hdr):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print(f"{h:90s} {first[i]} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write(src)
    print("\n## source page summary (tools/ncu_source_summary.py)")
    out = subprocess.run([sys.executable, str(Path(__file__).with_name("ncu_source_summary.py")), f.name, "16"],
                         capture_output=True, text=True)
    print(out.stdout + out.stderr)


if __name__ == "__main__":
    main()
