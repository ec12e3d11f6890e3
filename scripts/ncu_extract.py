"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the few metrics the roofline needs.
  python scripts/ncu_extract.py gpurun_out/roof_TAG.raw.csv [out.txt] [traffic.json key]
"""
import csv
import json
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    traffic = []
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"{w:85s} {r[i]} {units[i]}")
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            t = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            traffic.append(t)
            out.append(f"{'dram traffic (read+write), bytes':85s} {t:.0f}")
        except ValueError:
            pass
        out.append("-" * 100)
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    if len(sys.argv) > 4 and traffic:
        path, key = sys.argv[3], sys.argv[4]
        try:
            d = json.load(open(path))
        except Exception:  # noqa: BLE001
            d = {}
        d[key] = sum(traffic) / len(traffic)
        json.dump(d, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
