"""Summarise an .ncu-rep (ncu --set full) into the small JSON kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep out.json "note" """
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__cycles_active.avg",
        "smsp__cycles_active.min", "smsp__cycles_active.max", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        e = {"kernel": d["Kernel Name"][0]}
        for k in KEYS:
            if k in d:
                e[k] = f"{d[k][0]} {d[k][1]}".strip()
        st = sorted(((float(v[0].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for h, v in d.items()
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)
        e["top_stalls_warps_per_issue"] = {n: round(x, 3) for x, n in st[:6]}
        launches.append(e)
    json.dump({"note": note, "launches": launches}, open(out, "w"), indent=1)
    print(json.dumps(launches[0], indent=1))


if __name__ == "__main__":
    main()
