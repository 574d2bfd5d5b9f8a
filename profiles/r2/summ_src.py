"""Per-kernel digest of an `ncu --page source --csv` page (gzipped): stall totals, samples per barrier-delimited
phase, hottest SASS lines.   python profiles/r2/summ_src.py gpurun_out/x_source.csv.gz [kernel-substring] [top]"""
import csv, gzip, io, sys

def blocks(path):
    rows = csv.reader(io.TextIOWrapper(gzip.open(path)))
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if cur: yield cur
            cur = {"name": r[1], "hdr": None, "rows": []}
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    if cur: yield cur

def main(path, want="", top=18):
    top = int(top)
    seen = set()
    for b in blocks(path):
        if want not in b["name"]: continue
        i = {h: k for k, h in enumerate(b["hdr"])}
        key = (b["name"], len(b["rows"]), sum(int(r[i["# Samples"]] or 0) for r in b["rows"][:200]))
        if key in seen: continue
        seen.add(key)
        rs = b["rows"]
        tot = sum(int(r[i["# Samples"]] or 0) for r in rs) or 1
        inst = sum(int(r[i["Instructions Executed"]] or 0) for r in rs)
        stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        agg = {s: sum(int(r[i[s]] or 0) for r in rs) for s in stalls}
        print(f"\n## {b['name'][:80]}: {len(rs)} SASS lines, {tot} samples, {inst} warp instructions executed")
        print("stalls: " + ", ".join(f"{s[6:]} {v/tot:.1%}" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        # phases between barriers
        ph, acc, insts, out = 0, 0, 0, []
        for r in rs:
            acc += int(r[i["# Samples"]] or 0); insts += int(r[i["Instructions Executed"]] or 0)
            if "BAR.SYNC" in r[i["Source"]]:
                out.append((ph, acc, insts)); ph += 1; acc = 0; insts = 0
        out.append((ph, acc, insts))
        print("samples per barrier-delimited segment (static order): " + "  ".join(f"[{p}] {a/tot:.1%}/{n/max(inst,1):.1%}i" for p, a, n in out))
        w = i.get("L1 Wavefronts Shared"); wi = i.get("L1 Wavefronts Shared Ideal")
        if w is not None:
            tw = sum(int(r[w] or 0) for r in rs); ti = sum(int(r[wi] or 0) for r in rs)
            print(f"shared wavefronts {tw} (ideal {ti})")
        for r in sorted(rs, key=lambda r: -int(r[i["# Samples"]] or 0))[:top]:
            n = int(r[i["# Samples"]] or 0)
            dom = max(stalls, key=lambda s: int(r[i[s]] or 0))
            extra = f" wf {r[w]}/{r[wi]}" if w is not None and int(r[w] or 0) else ""
            print(f"  {n/tot:6.1%} {dom[6:]:<12} x{r[i['Instructions Executed']]:>10}  {r[i['Source']].strip()[:80]}{extra}")

if __name__ == "__main__":
    main(*sys.argv[1:])
