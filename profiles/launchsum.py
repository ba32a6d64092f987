"""Summary of an ncu launch-list CSV (--metrics gpu__time_duration.sum,... --csv): python profiles/launchsum.py file.csv [first_id]"""
import csv
import sys
from collections import OrderedDict


def main(path, first=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki].split("::")[-1][:34], r[gi]), {})[r[mi]] = r[vi]

    def f(v, x):
        try:
            return float(v.get(x, "0").replace(",", ""))
        except ValueError:
            return -1.0
    for k, v in d.items():
        if k[0] < first:
            continue
        print(f"{k[0]:4d} {k[1]:34s} {k[2]:14s} t={f(v, 'gpu__time_duration.sum') / 1e3:8.1f}us rd={f(v, 'dram__bytes_read.sum') / 1e6:6.0f}MB "
              f"wr={f(v, 'dram__bytes_write.sum') / 1e6:6.0f}MB tensor={f(v, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):5.1f}% "
              f"issue={f(v, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}% "
              f"smem_conf={f(v, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:6.1f}M/{f(v, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:6.1f}M")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
