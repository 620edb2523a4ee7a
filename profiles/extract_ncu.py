"""Extracts the judged counters from `ncu -i <rep> --page raw --csv` output.
usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/prof.csv; python profiles/extract_ncu.py /tmp/prof.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "smsp__inst_executed.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel:", name[:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:72s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
