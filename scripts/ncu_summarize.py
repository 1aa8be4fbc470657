#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the markdown summary and the traffic JSON kept under profiles/.

usage: ncu -i gpurun_out/step.ncu-rep --page raw --csv > /tmp/raw.csv
       python scripts/ncu_summarize.py /tmp/raw.csv profiles/r1_ncu_summary.md profiles/r1_traffic.json 16 "<command>"
"""
import csv
import json
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    raw, out_md, out_json, pairs, command = sys.argv[1:6]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    md = [f"# Round 1 -- ncu summaries of one full step (B200)\n",
          f"Command: `{command}`.\nThe kernels below are the launches of one step, in order. Times under ncu are cold-cache and "
          "serialised; `profiles/r1_launches.csv` is the `--metrics gpu__time_duration.sum` launch list of the same "
          "command, `profiles/r1_bench_c2_b16.json` the un-profiled bench line of the same build.\n",
          f"Pairs per launch: {pairs}. Algorithmic bytes per launch: DESIGN.md section 4.\n"]
    kernels = {}
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
        base = short.split("<")[0]
        kernels.setdefault(base, []).append({"instance": short, "dram_bytes": dram, "duration_ms": ms})
    open(out_md, "w").write("\n".join(md))
    json.dump({"source": f"{out_md} (ncu --set full, {command})", "pairs_per_launch": int(pairs), "kernels": kernels},
              open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
