"""Extract the roofline-relevant metrics from an `ncu --set full` report:  python profiles/summarize_full.py rep.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"--- {name}  (id {r[idx['ID']]})")
        for w in WANT:
            if w in idx:
                print(f"    {w:75s} {r[idx[w]]:>16s} {units[idx[w]]}")
        rd, wr = r[idx["dram__bytes_read.sum"]], r[idx["dram__bytes_write.sum"]]
        print(f"    {'traffic = dram read + write':75s} {rd} {units[idx['dram__bytes_read.sum']]} + {wr} {units[idx['dram__bytes_write.sum']]}")


if __name__ == "__main__":
    main(sys.argv[1])
