"""Text summary of an `ncu --set full` report, one section per profiled launch (no GPU needed):
    python profiles/summarize.py gpurun_out/<name>.ncu-rep ["title of launch 0" "title of launch 1" ...] > profiles/<name>.txt
Reads the report with `ncu -i ... --page raw --csv`; the per-source-line view is profiles/srcsum.py."""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def main(rep, titles):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print(f"report: {rep}")
    for li, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0]
        print(f"\n== launch {li}: {titles[li] if li < len(titles) else ''} ({name})")
        for k in RAW:
            if k in d:
                print(f"  {k} = {d[k]} {u.get(k, '')}")
        st = {}
        for k, v in d.items():
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("not_issued"):
                try:
                    st[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = int(v.replace(",", ""))
                except ValueError:
                    pass
        tot = sum(st.values()) or 1
        print("  warp-state samples: " + ", ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
