"""Bucket an `ncu --page source --csv` export of one kernel: share of issued warp-instructions and of stall samples inside the
FFMA2 stencil bodies vs everywhere else, the execution-count histogram (which loops run how often), and the top stall sites.

    python scripts/ncu_source_buckets.py gpurun_out/r02c_dwpw_source.csv > profiles/r02_ncu_dwpw_source.txt
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    print("#", rows[0][1] if len(rows[0]) > 1 else "")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r, "# Samples") for r in data)
    totex = sum(f(r, "Instructions Executed") for r in data)
    print(f"{len(data)} SASS instructions, {totex / 1e6:.1f} M warp-instructions executed, {int(tot)} stall samples")
    agg = {s: sum(f(r, s) for r in data) for s in stalls}
    print("stall mix:", ", ".join(f"{k[6:]} {v / tot:.2f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = [i for i, r in enumerate(data) if "FFMA2" in r[ix["Source"]]]
    inside = set()
    if idx:
        inside = set(range(max(0, idx[0] - 60), min(len(data) - 1, idx[-1] + 40) + 1))
    for name, sel in (("stencil bodies (FFMA2 region)", [r for i, r in enumerate(data) if i in inside]),
                      ("everything else", [r for i, r in enumerate(data) if i not in inside])):
        s = sum(f(r, "# Samples") for r in sel)
        ex = sum(f(r, "Instructions Executed") for r in sel)
        a = {st: sum(f(r, st) for r in sel) for st in stalls}
        top = ", ".join(f"{k[6:]} {v / max(s, 1):.2f}" for k, v in sorted(a.items(), key=lambda kv: -kv[1])[:5])
        print(f"{name}: {len(sel)} instr, {ex / totex * 100:.1f} % of warp-instructions, {s / tot * 100:.1f} % of samples ({top})")
    print("\nexecution-count histogram (count x instructions = share of warp-instructions):")
    c = collections.Counter(int(f(r, "Instructions Executed")) for r in data)
    for k, v in sorted(c.items(), key=lambda kv: -kv[0] * kv[1])[:10]:
        print(f"  executed {k:9d} x {v:4d} instructions = {k * v / totex * 100:5.1f} %")
    print("\ntop stall sites:")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:12]:
        st = sorted(((s, f(r, s)) for s in stalls if f(r, s) > 0), key=lambda kv: -kv[1])[:2]
        print(f"  {f(r, '# Samples') / tot * 100:5.2f} %  x{int(f(r, 'Instructions Executed')):8d}  {r[ix['Source']].strip()[:60]:60s} "
              + ", ".join(f"{k[6:]} {int(v)}" for k, v in st))


if __name__ == "__main__":
    main()
