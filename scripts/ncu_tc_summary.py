"""Per-kernel table out of an `ncu --set full ... --page raw --csv` dump of one training step: duration, DRAM / L2 traffic,
tensor-pipe activity (the north star's "tensor-pipe utilisation" of the basis contraction), shared-memory bank conflicts,
issue rate and the top stall reasons.
  python scripts/ncu_tc_summary.py gpurun_out/step_full_TAG.raw.csv > profiles/rNN_step_tc_kernels_full.txt"""
import csv
import sys


def f(d, k, default=0.0):
    try:
        return float(d.get(k, "").replace(",", ""))
    except ValueError:
        return default


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))

    def to_bytes(d, k):
        v = f(d, k)
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit.get(k, "byte").lower(), 1)

    def to_us(d, k):
        v = f(d, k)
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit.get(k, "us").lower(), 1)

    print("# one graph-replayed training step at 64 meshes under `ncu --set full --clock-control none` (serialised, cold caches)")
    print("# tensor% = sm__pipe_tensor_cycles_active (realtime) % of elapsed; umma = tcgen05.mma instructions; conflicts = shared-memory")
    print("# bank-conflict wavefronts / all shared wavefronts; ipc = warp instructions per active SM cycle; stalls = warps per issue")
    print(f"{'kernel':44s} {'grid':>6s} {'us':>7s} {'dramMB':>7s} {'L2rdMB':>7s} {'L1hit%':>6s} {'tensor%':>7s} {'umma':>7s} {'lsu%':>5s} {'confl%':>6s} {'ipc':>5s}  top stalls")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")[:44]
        stalls = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                stalls.append((f(d, k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        sw = f(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        bc = f(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        inst = f(d, "smsp__inst_executed.sum")
        cyc = f(d, "sm__cycles_active.sum")
        print(f"{name:44s} {int(f(d, 'launch__grid_size')):6d} {to_us(d, 'gpu__time_duration.sum'):7.1f} "
              f"{(to_bytes(d, 'dram__bytes_read.sum') + to_bytes(d, 'dram__bytes_write.sum')) / 1e6:7.1f} "
              f"{f(d, 'lts__t_sectors_srcunit_tex_op_read.sum') * 32 / 1e6:7.1f} {f(d, 'l1tex__t_sector_hit_rate.pct'):6.1f} "
              f"{f(d, 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):7.2f} "
              f"{int(f(d, 'smsp__sass_inst_executed_op_utcmma.sum')):7d} "
              f"{f(d, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):5.1f} {100 * bc / sw if sw else 0:6.1f} "
              f"{inst / cyc if cyc else 0:5.2f}  " + ", ".join(f"{n} {v:.1f}" for v, n in stalls[:3]))


if __name__ == "__main__":
    main()
