"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv profiles/r1_launches_summary.md
  python profiles/summarize.py kernel gpurun_out/x.ncu-rep profiles/r1_x.md [kernel-index]
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * 1e3 if u in ("s", "second") else v
        a = agg.setdefault(name, [0, 0.0, r[idx["Grid Size"]], r[idx["Block Size"]]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command; per-launch times "
                "are cold-cache and serialised — compare SHARES.\n\n")
        f.write("| kernel | launches | total ms | share | ms / launch | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, (n, ms, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {ms / tot:.1%} | {ms / n:.3f} | {g} | {b} |\n")
        f.write(f"\ntotal kernel time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def kernel(src, dst, which=0):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    r = rows[2 + which]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: {r[idx['Kernel Name']][:100]}\n\nsource: `{src}` (launch {which})\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in idx:
                f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
        srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--kernel-id", f":::{which + 1}"],
                              capture_output=True, text=True).stdout
        srows = list(csv.reader(srcp.splitlines()))
        if len(srows) > 2:
            h2 = srows[1]
            i2 = {h: i for i, h in enumerate(h2)}
            data = [x for x in srows[2:] if len(x) == len(h2) and x[i2["# Samples"]].isdigit()]
            data = data[:len(data) // 2] if len(data) > 1 and data[0] == data[1] else data
            seen, uniq = set(), []
            for x in data:
                key = (x[i2["Address"]], x[i2["Source"]])
                if key not in seen:
                    seen.add(key)
                    uniq.append(x)
            stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(int(x[i2["# Samples"]]) for x in uniq) or 1
            agg = {s: sum(int(x[i2[s]] or 0) for x in uniq) for s in stalls}
            f.write("\n## warp stall samples (all instructions)\n\n| stall | samples | share |\n|---|---:|---:|\n")
            for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
                f.write(f"| {s} | {v} | {v / tot:.1%} |\n")
            f.write("\n## hottest SASS instructions\n\n| samples | instruction | top stall |\n|---:|---|---|\n")
            for x in sorted(uniq, key=lambda x: -int(x[i2["# Samples"]]))[:14]:
                st = max(stalls, key=lambda s: int(x[i2[s]] or 0))
                f.write(f"| {x[i2['# Samples']]} | `{x[i2['Source']].strip()[:70]}` | {st} |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
