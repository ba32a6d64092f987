"""Text summary of an `ncu --set full` report: python profiles/summarize.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt
(reads the report with `ncu -i ... --page raw/source --csv`; no GPU needed)."""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    d = dict(zip(rows[0], rows[-1]))
    units = dict(zip(rows[0], rows[1]))
    print(f"report: {rep}\nkernel: {d.get('Kernel Name')}")
    for k in RAW:
        if k in d:
            print(f"  {k} = {d[k]} {units.get(k, '')}")
    st = {k[len('smsp__pcsamp_warps_issue_stalled_'):]: int(v) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued")}
    tot = sum(st.values()) or 1
    print("warp-state samples (pc sampling):")
    for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
        print(f"  {k:22s} {100.0 * v / tot:5.1f} %")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not hi:
        return
    hdr, data = rows[hi[0]], rows[hi[0] + 1:]
    iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    lines = [r for r in data if r and r[0].isdigit() and len(r) > iex]
    tex, ts = sum(num(r[iex]) for r in lines) or 1, sum(num(r[isamp]) for r in lines) or 1
    print("hottest source lines (share of executed instructions / of samples):")
    for r in sorted(lines, key=lambda r: -num(r[iex]))[:14]:
        print(f"  line {r[0]:>4s}  {100.0 * num(r[iex]) / tex:5.1f} %  {100.0 * num(r[isamp]) / ts:5.1f} %  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
