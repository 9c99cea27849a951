"""Summarise the raw-metric CSVs written by tools/ncu_capture.sh (one `ncu --set full` launch per kernel) into
profiles/<tag>_ncu_kernels.md and profiles/<tag>_traffic.json (dram bytes per launch, read by bench.py).
    python tools/ncu_raw_summary.py r2"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "HMMA pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio")]
KEY = {"inproj": "inproj", "keyproj": "keyproj", "attn_bwd7": "attn_bwd"}


def main():
    tag = sys.argv[1]
    out = [f"# {tag}: `ncu --set full --clock-control none` of one launch of each dominant kernel\n\n"
           "Captured by `tools/ncu_capture.sh` on the B200 box from the first (eager) train step of BASELINE config 2 "
           "(B = 512, S0): the audio unit's launch of each kernel. Times under the profiler are not bench values.\n\n"]
    traffic = {}
    for f in sorted((ROOT / "gpurun_out").glob(f"ncu_{tag}_*.csv")):
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units, d = rows[0], rows[1], rows[2]
        name = f.stem[len(f"ncu_{tag}_"):]
        out.append(f"## {name}: `{d[hdr.index('Kernel Name')].replace('sdumc::', '')[:90]}`\n\n| metric | value |\n|---|---|\n")
        for m, label in KEEP:
            if m in hdr:
                out.append(f"| {label} | {d[hdr.index(m)]} {units[hdr.index(m)]} |\n")
        out.append("\n")
        try:
            def b(m):
                v, u = float(d[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            out.append(f"dram traffic per launch: {tot / 1e6:.1f} MB\n\n")
            if name in KEY:
                traffic[KEY[name]] = int(tot)
        except Exception as e:  # noqa: BLE001
            out.append(f"(traffic: {e})\n\n")
    (ROOT / "profiles" / f"{tag}_ncu_kernels.md").write_text("".join(out))
    (ROOT / "profiles" / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1))
    print("".join(out)[:6000])
    print(traffic)


if __name__ == "__main__":
    main()
