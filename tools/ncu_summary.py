"""Summarise an .ncu-rep (ncu --set full --import-source on) into markdown: headline metrics per captured
launch + the hottest SASS lines of one launch.   python tools/ncu_summary.py REP OUT.md [launch_index]"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum"]


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    md = [f"# ncu summary of `{rep}`\n\n`ncu --set full --clock-control none --import-source on` under gpurun; "
          "times under the profiler are not bench values.\n\n"]
    cols = [c for c in KEEP if c in hdr]
    md.append("| # | kernel | " + " | ".join(c.split(".")[0].replace("__", " ") for c in cols) + " |\n")
    md.append("|---|---|" + "---|" * len(cols) + "\n")
    for i, d in enumerate(data):
        name = d[hdr.index("Kernel Name")].replace("sdumc::", "")[:60]
        md.append(f"| {i} | `{name}` | " + " | ".join(f"{d[hdr.index(c)]} {units[hdr.index(c)]}" for c in cols) + " |\n")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--launch-skip", str(which),
                                           "--launch-count", "1"]))))
    if len(src) > 3:
        h = src[1]
        body = [r for r in src[2:] if len(r) == len(h) and r[0] != "Address"]
        iw, isrc = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]

        def I(x):
            try:
                return int(x)
            except ValueError:
                return 0
        tot = sum(I(r[iw]) for r in body) or 1
        agg = sorted(((s, sum(I(r[h.index(s)]) for r in body)) for s in stalls), key=lambda x: -x[1])[:6]
        md.append(f"\n## launch {which}: warp-stall samples by reason\n\n" +
                  ", ".join(f"{s[6:]} {100 * v / max(1, sum(v for _, v in agg)):.0f}%" for s, v in agg) + "\n")
        md.append("\n## hottest SASS lines (share of stall samples)\n\n| share | SASS | top reasons |\n|---|---|---|\n")
        seen = set()
        for r in sorted(body, key=lambda r: -I(r[iw])):
            key = (r[isrc], r[iw])
            if key in seen:
                continue
            seen.add(key)
            top = sorted(((s[6:], I(r[h.index(s)])) for s in stalls), key=lambda x: -x[1])[:2]
            md.append(f"| {100 * I(r[iw]) / tot:.1f}% | `{r[isrc][:80]}` | {top[0][0]} {top[0][1]}, {top[1][0]} {top[1][1]} |\n")
            if len(seen) >= 14:
                break
    open(out, "w").write("".join(md))
    print("".join(md)[:3000])


if __name__ == "__main__":
    main()
