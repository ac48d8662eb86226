"""Key metrics of an `ncu --set full` report, one block per captured launch (run where ncu is installed)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                print(f"   {w:80s} {r[idx[w]]:>16s} {units[idx[w]]}")
        # warp-state sampling: share of each stall reason over all samples of the launch
        st = []
        for h, i in idx.items():
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    st.append((float(r[i].replace(",", "")), h.split("stalled_")[-1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("   warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
