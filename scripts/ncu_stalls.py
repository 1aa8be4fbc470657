#!/usr/bin/env python
"""Per-kernel warp-stall and opcode summary from `ncu -i X.ncu-rep --page source --csv | gzip` (scripts/gpu_ncu.sh).

usage: python scripts/ncu_stalls.py SOURCE.csv.gz OUT.md "<title>"
"""
import collections
import csv
import gzip
import io
import sys


def main():
    src, out, title = sys.argv[1:4]
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(src))))
    md = [f"# {title}\n",
          "Warp-state samples of `ncu --set full --import-source on` (SASS view), per kernel: where the warps' time goes, "
          "and which opcodes executed most.  `selected` = issuing; everything else = waiting for that reason.\n"]
    i = 0
    seen = set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            ix = {h: k for k, h in enumerate(hdr)}
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if rows[j]:
                    body.append(rows[j])
                j += 1
            i = j
            if "Source" not in ix or "# Samples" not in ix:
                continue
            first = body[0][ix["Source"]] if body else ""
            if not first or first.lstrip().startswith(("#", "//", "template", "namespace")):
                continue   # the CUDA-C view of the same kernel: the SASS view carries the same samples
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot = collections.Counter()
            ops = collections.Counter()
            opn = collections.Counter()
            S = 0
            for r in body:
                try:
                    n = int(r[ix["# Samples"]])
                except ValueError:
                    continue
                S += n
                for s in stalls:
                    v = r[ix[s]]
                    if v not in ("", "0"):
                        tot[s] += int(v)
                t = r[ix["Source"]].split()
                if not t:
                    continue
                o = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
                ops[o] += n
                try:
                    opn[o] += int(r[ix["Instructions Executed"]] or 0)
                except ValueError:
                    pass
            if S == 0 or (name, S) in seen:
                continue   # the source page lists a kernel once per view
            seen.add((name, S))
            md.append(f"## {name}\n")
            md.append(f"{S} samples.\n")
            md.append("| stall reason | share of samples |\n|---|---|")
            for k, v in tot.most_common(10):
                md.append(f"| {k[6:]} | {100 * v / S:.1f} % |")
            md.append("\n| opcode | share of samples |\n|---|---|")
            for o, n in ops.most_common(14):
                md.append(f"| {o} | {100 * n / S:.1f} % |")
            md.append("")
        else:
            i += 1
    open(out, "w").write("\n".join(md))
    print("wrote", out)


if __name__ == "__main__":
    main()
