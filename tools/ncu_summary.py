"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_op_shared_st.sum"]


def main(path, kernel_filter=None, as_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    records = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if kernel_filter and kernel_filter not in name:
            continue
        print("==", name[:90], "id", r[hdr.index("ID")])
        rec = {"kernel": name, "id": r[hdr.index("ID")]}
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    print(f"   {k} [{units[i]}] = {r[i]}")
                    rec[f"{k} [{units[i]}]"] = r[i]
        sec, req = rec.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum [sector]"), rec.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum []")
        try:
            rec["global_load_sectors_per_request"] = float(sec) / float(req)
            print("   global load sectors / request =", rec["global_load_sectors_per_request"])
        except (TypeError, ValueError, ZeroDivisionError):
            pass
        records.append(rec)
    if as_json:
        import json
        json.dump(records, open(as_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], (sys.argv[2] or None) if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
