"""Selected metrics of an `ncu --page raw --csv` export, one line per metric with one column per captured launch: how the
`profiles/r01f_*_ncu_raw.txt` files were made from gpurun_out/*_raw.csv (scripts/gpu_profile_round.sh).
Usage: python scripts/ncu_summary.py raw.csv out.txt ["header line"]"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(dram__bytes|dram__cycles_active|gpu__dram_throughput|gpu__time_duration|l1tex__data_pipe_lsu_wavefronts_mem_shared"
    r"|launch__(block_size|grid_size|registers|shared_mem|occupancy_limit)|sm__cycles_active|sm__cycles_elapsed|sm__throughput"
    r"|sm__inst_executed_pipe|sm__pipe_tensor|sm__pipe_.*cycles_active|smsp__cycles_active|smsp__inst_executed\.|smsp__issue"
    r"|smsp__warp_issue_stalled|smsp__average_warp|smsp__warps_active|lts__t_sector|lts__throughput|lts__t_bytes"
    r"|l1tex__throughput|smsp__pcsamp|sm__warps_active|smsp__inst_issued|sm__inst_executed\.|smsp__thread_inst"
    r"|sm__sass_inst_executed_op_shared|smsp__sass_inst_executed_op|sm__mio|tensor|sm__inst_executed_pipe_uniform"
    r"|lts__average|local_)")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    names, units, data = rows[0], rows[1], rows[2:]
    lines = [sys.argv[3]] if len(sys.argv) > 3 else []
    kn = names.index("Kernel Name")
    lines.append("Kernel Name []: " + " | ".join(r[kn] for r in data))
    for i, name in enumerate(names):
        if KEEP.match(name):
            lines.append(f"{name} [{units[i]}]: " + " | ".join(r[i] for r in data))
    open(dst, "w").write("\n".join(lines) + "\n")
    print(dst, len(lines), "lines,", len(data), "launch(es)")


if __name__ == "__main__":
    main()
