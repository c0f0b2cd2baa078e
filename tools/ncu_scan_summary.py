"""Summarises an `ncu --set full` capture of the persistent scan kernels (or any kernel) into a markdown table:
duration, IPC, issue-slot use, stall reasons per issued instruction, L2 / DRAM traffic, shared-memory wavefronts.

    python tools/ncu_scan_summary.py gpurun_out/r2_scan.ncu-rep > profiles/r2_scan_ncu.md
"""
import csv
import io
import subprocess
import sys


def rows(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = rows(rep, "raw")
    hdr, units = raw[0], dict(zip(raw[0], raw[1]))
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]

        def g(k, default="n/a"):
            return d.get(k, default)

        print(f"## `{name[:100]}`\n")
        print("| metric | value |\n|---|---|")
        for label, key in (("duration", "gpu__time_duration.sum"), ("SM clock", "sm__cycles_elapsed.avg.per_second"),
                           ("grid x block", "launch__grid_size"), ("threads per block", "launch__block_size"),
                           ("registers per thread", "launch__registers_per_thread"),
                           ("dynamic shared memory per block", "launch__shared_mem_per_block_dynamic"),
                           ("warp instructions executed", "smsp__inst_executed.sum"),
                           ("IPC (active cycles)", "sm__inst_executed.avg.per_cycle_active"),
                           ("issue slots busy (%)", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                           ("warp cycles per issued instruction", "smsp__average_warp_latency_per_inst_issued.ratio"),
                           ("L2 sectors read by SMs", "lts__t_sectors_srcunit_tex_op_read.sum"),
                           ("L2 hit rate (%)", "lts__t_sector_hit_rate.pct"),
                           ("DRAM bytes read", "dram__bytes_read.sum"), ("DRAM bytes written", "dram__bytes_write.sum"),
                           ("shared-memory wavefronts (LSU)", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                           ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                           ("tensor pipe active (% of active cycles)", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                           ("fma pipe active (%)", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")):
            if key in d:
                print(f"| {label} | {d[key]} {units.get(key, '')} |")
        stalls = []
        for h, v in d.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(v.replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("\nStall reasons (warps stalled per issued instruction; share):\n")
        print("| reason | warps / issue | share |\n|---|---|---|")
        for v, h in sorted(stalls, reverse=True)[:10]:
            print(f"| {h} | {v:.2f} | {100 * v / tot:.0f} % |")
        print()


if __name__ == "__main__":
    main()
