"""Executed-instruction mix and hottest stall sites from an `ncu --page source --csv` export (optionally .gz).

    python scripts/ncu_source_mix.py gpurun_out/x_source.csv.gz [kernel-substring]
"""
import collections
import csv
import gzip
import io
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
    kernel, hdr, ops, tot, samples = None, None, None, 0, None
    def flush():
        if kernel is None or ops is None or (want and want not in kernel):
            return
        print("==", kernel[:120])
        print("   warp-instructions executed: %.2f M" % (tot / 1e6))
        print("   mix: " + "  ".join(f"{k}={v / tot * 100:.1f}%" for k, v in ops.most_common(16)))
        top = sorted(samples, reverse=True)[:12]
        stot = sum(s for s, _ in samples) or 1
        for s, txt in top:
            print(f"   {s / stot * 100:5.1f}%  {txt}")
    for r in csv.reader(f):
        if not r:
            continue
        if r[0] == "Kernel Name":
            flush()
            kernel, hdr, ops, tot, samples = r[1], None, collections.Counter(), 0, []
            continue
        if r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None:
            continue
        try:
            n = int(r[hdr["Instructions Executed"]])
            s = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        txt = r[hdr["Source"]].strip()
        parts = txt.split()
        op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
        ops[op.split(".")[0]] += n
        tot += n
        samples.append((s, txt[:90]))
    flush()


if __name__ == "__main__":
    main()
