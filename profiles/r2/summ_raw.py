"""Summarise `ncu --page raw --csv` (+ optional `--page source --csv`, gzipped) pages brought back from the GPU box.
  python profiles/r2/summ_raw.py gpurun_out/r2_c2_full [out.md]"""
import csv, gzip, io, sys, collections

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "smsp__pcsamp_sample_buffer_full.sum"]

def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None

def main(prefix, out=None):
    rows = list(csv.reader(open(prefix + "_raw.csv")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    o = io.StringIO()
    kern = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        kern.append(name)
        o.write(f"\n## {name}  (launch id {r[idx['ID']]}, grid {r[idx['launch__grid_size']]}, block {r[idx['launch__block_size']]})\n\n")
        o.write("| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in idx: o.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
        t = num(r[idx["gpu__time_duration.sum"]]); tu = units[idx["gpu__time_duration.sum"]]
        rd, wr = num(r[idx["dram__bytes_read.sum"]]), num(r[idx["dram__bytes_write.sum"]])
        def tob(v, u): return v * {"byte":1, "Kbyte":1e3, "Mbyte":1e6, "Gbyte":1e9}[u]
        def tos(v, u): return v * {"ns":1e-9, "us":1e-6, "usecond":1e-6, "ms":1e-3, "msecond":1e-3, "nsecond":1e-9, "second":1, "s":1}[u]
        b = tob(rd, units[idx["dram__bytes_read.sum"]]) + tob(wr, units[idx["dram__bytes_write.sum"]])
        s = tos(t, tu)
        o.write(f"| DRAM traffic | {b/1e9:.3f} | GB |\n| DRAM GB/s | {b/s/1e9:.0f} | GB/s |\n")
        # stall reasons from raw page
        st = [(h, num(r[i])) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for h, v in st if v]
        tot = sum(v for _, v in st) or 1
        o.write("\nwarp stalls (warps per issue-active cycle): " + ", ".join(f"{h} {v:.2f} ({v/tot:.0%})" for h, v in sorted(st, key=lambda kv: -kv[1])[:7]) + "\n")
    try:
        src = list(csv.reader(io.TextIOWrapper(gzip.open(prefix + "_source.csv.gz"))))
    except Exception:
        src = []
    # source page: blocks per kernel, header row starts with "Address" or contains "# Samples"
    k = -1; h2 = None; data = collections.defaultdict(list)
    for r in src:
        if "Source" in r and any(c.startswith("# Samples") or c == "Sampling Data (All)" for c in r):
            h2 = r; k += 1; continue
        if h2 and len(r) == len(h2): data[k].append(r)
    for k, rs in data.items():
        i2 = {h: i for i, h in enumerate(h2)}
        sc = "Sampling Data (All)" if "Sampling Data (All)" in i2 else "# Samples"
        ic = "Instructions Executed" if "Instructions Executed" in i2 else None
        rs2 = [r for r in rs if r[i2[sc]].replace(",", "").isdigit()]
        tot = sum(int(r[i2[sc]].replace(",", "")) for r in rs2) or 1
        o.write(f"\n### hottest SASS lines, kernel #{k} ({kern[k] if k < len(kern) else '?'}): {tot} samples\n\n")
        for r in sorted(rs2, key=lambda r: -int(r[i2[sc]].replace(",", "")))[:14]:
            o.write(f"    {int(r[i2[sc]].replace(',', ''))/tot:6.1%}  {r[i2['Source']][:90]}" + (f"   [inst {r[i2[ic]]}]" if ic else "") + "\n")
    txt = o.getvalue()
    if out: open(out, "w").write(f"# ncu --set full, {prefix}\n" + txt)
    else: print(txt)

if __name__ == "__main__":
    main(*sys.argv[1:])
