"""Summarise an .ncu-rep (read on the CPU box): one block per profiled launch with the metrics
DESIGN.md / bench.py quote.   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    # L1 data-pipe wavefronts (what bounds the gather: DESIGN.md 4.2 quotes wavefronts per load from these)
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_per_warp.ratio", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    extra = [h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")]
    for r in data:
        print("=" * 100)
        print(r[idx["ID"]], r[idx["Kernel Name"]][:90])
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
        # derived: wavefronts per global-load request (1 = every lane of a request inside one 128-byte line)
        try:
            wf = float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_lg.sum"]].replace(",", ""))
            rq = float(r[idx["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]].replace(",", "")) + \
                float(r[idx["l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]].replace(",", ""))
            if rq > 0:
                print(f"  {'derived: L1 wavefronts per global ld/st request':72s} {wf / rq:18.2f}")
        except (KeyError, ValueError):
            pass
        stalls = sorted(((float(r[idx[h]].replace(',', '') or 0), h) for h in extra), reverse=True)[:8]
        for v, h in stalls:
            print(f"  stall {h[len('smsp__average_'):-len('_per_issue_active.ratio')]:64s} {v:10.2f}")


if __name__ == "__main__":
    main()
