"""Text summary of an ncu report (run here, no GPU needed):  python tools/summarize_ncu.py REPORT.ncu-rep > profiles/NAME.txt

Prints the roofline-relevant raw metrics of every captured launch and the 25 hottest SASS lines with their
dominant stall reasons (needs -lineinfo / --import-source on at capture time)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== launch:", d.get("Kernel Name", "?")[:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
    if len(src) > 3:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        data = [r for r in src[2:] if len(r) == len(h) and r[ix["# Samples"]].strip().isdigit()]   # several kernels: repeated header rows
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
        print(f"== hottest SASS lines (warp-state samples, total {tot}; idle warps parked on barriers included)")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
            st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
            print(f"  {int(r[ix['# Samples']] or 0):6d}  {r[ix['Source']][:72]:72s} {st}")


if __name__ == "__main__":
    main()
