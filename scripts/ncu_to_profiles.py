#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` of ONE bench step into profiles/r2_ncu_<tag>.md and an entry of
profiles/r2_traffic.json (per-pass DRAM bytes, tied to the sha256 of kangaroo_b200/csrc so that bench.py stops quoting it
once the kernels change).

usage: python scripts/ncu_to_profiles.py RAW.csv TAG WORKLOAD PAIRS_PER_LAUNCH "<command that was profiled>"
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_hash  # noqa: E402

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    raw, tag, wl, pairs, command = sys.argv[1:6]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    md = [f"# Round 2 -- ncu summary: {wl} ({tag})\n",
          f"Command: `{command}` (one bench step, `--set full --clock-control none`; times under ncu are cold-cache and "
          f"serialised).  Pairs per launch: {pairs}.  Kernel sources: sha256 {csrc_hash()[:16]}.\n"]
    passes = []
    for r in data:
        name = r[ix["Kernel Name"]]
        md.append(f"## {name}\n")
        md.append(f"- Grid Size: {r[ix['Grid Size']]} \n- Block Size: {r[ix['Block Size']]} ")
        val = {}
        for k in KEYS:
            if k in ix:
                md.append(f"- {k}: {r[ix[k]]} {units[ix[k]]}")
                try:
                    val[k] = float(r[ix[k]].replace(",", "")) * UNIT_SCALE.get(units[ix[k]], 1.0)
                except ValueError:
                    pass
        dram = val.get("dram__bytes_read.sum", 0.0) + val.get("dram__bytes_write.sum", 0.0)
        ms = val.get("gpu__time_duration.sum", 0.0)
        if ms:
            md.append(f"- dram traffic (read+write): {dram / 1e9:.3f} GB -> {dram / 1e9 / (ms * 1e-3):.0f} GB/s over the launch")
        top = sorted(((float(r[ix[s]]), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                      for s in stalls if r[ix[s]] not in ("", "n/a")), reverse=True)[:5]
        md.append("- top stalls (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in top) + "\n")
        short = re.sub(r"^void\s+", "", name).split("(")[0]
        if re.search(r"sgm_(vgroup|sweep|hsweep)_kernel", short):
            passes.append({"kernel": short, "dram_bytes": dram, "duration_ms": ms,
                           "inst_executed": val.get("smsp__inst_executed.sum"),
                           "issue_active_pct": val.get("smsp__issue_active.avg.pct_of_peak_sustained_active")})
    open(os.path.join(ROOT, "profiles", f"r2_ncu_{tag}.md"), "w").write("\n".join(md))
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    tr = json.load(open(tp)) if os.path.exists(tp) else {}
    if tr.get("csrc_sha256") != csrc_hash():
        tr = {"csrc_sha256": csrc_hash(), "workloads": {}}
    tr["workloads"][wl] = {"pairs_per_launch": int(pairs), "source": f"profiles/r2_ncu_{tag}.md ({command})", "passes": passes}
    json.dump(tr, open(tp, "w"), indent=1)
    print(f"{tag}: {len(data)} kernels, {len(passes)} aggregation passes")


if __name__ == "__main__":
    main()
