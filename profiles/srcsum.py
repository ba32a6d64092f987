"""Per-source-line summary of an `ncu --set full --import-source on` report (no GPU needed):
    python profiles/srcsum.py report.ncu-rep [top_n]
For every profiled launch: the hottest CUDA source lines by warp-state samples, with executed instructions, shared-memory
wavefronts (ideal / actual) and the two dominant stall reasons."""
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(rep, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"] + [len(rows)]
    launch = -1
    for a, b in zip(heads[:-1], heads[1:]):
        hdr = rows[a]
        fname = rows[a - 1][1] if a > 0 and rows[a - 1] and rows[a - 1][0] == "File Name" else "?"
        data = [r for r in rows[a + 1:b] if r and r[0].isdigit() and len(r) >= len(hdr) - 2]
        if not data:
            continue
        isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        iws, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        agg = {}
        for r in data:
            k = int(r[0])
            e = agg.setdefault(k, {"src": r[1].strip(), "samp": 0.0, "inst": 0.0, "ws": 0.0, "wi": 0.0, "st": {}})
            e["samp"] += num(r[isamp]); e["inst"] += num(r[iex]); e["ws"] += num(r[iws]); e["wi"] += num(r[iwi])
            for i, h in stalls:
                e["st"][h] = e["st"].get(h, 0.0) + num(r[i])
        ts = sum(e["samp"] for e in agg.values()) or 1.0
        ti = sum(e["inst"] for e in agg.values()) or 1.0
        if ts < 100:
            continue
        if fname.endswith("tq_stream.cu") or "tq_kernels.cu" in fname:
            launch += 1
        print(f"=== section file={fname} (launch ~{launch}) samples={ts:.0f} inst={ti:.0f} "
              f"smem wavefronts={sum(e['ws'] for e in agg.values()):.0f} ideal={sum(e['wi'] for e in agg.values()):.0f}")
        for k, e in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
            st = sorted(e["st"].items(), key=lambda kv: -kv[1])[:2]
            sts = " ".join(f"{h[6:]}={100 * v / max(e['samp'], 1):.0f}%" for h, v in st)
            print(f"  L{k:4d} samp {100 * e['samp'] / ts:5.1f}% inst {100 * e['inst'] / ti:5.1f}% ws {e['ws'] / 1e6:6.2f}M/{e['wi'] / 1e6:6.2f}M "
                  f"[{sts}] {e['src'][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
