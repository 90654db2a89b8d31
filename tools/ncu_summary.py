"""Summarise one `ncu --set full` report into the two JSON files bench.py and DESIGN.md cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_summary_r01.json profiles/traffic_r01.json

Per kernel (fwd / energy / bwd, matched on the kernel name): duration, DRAM bytes, pipe and
shared-memory utilisation, stall-sample shares.  traffic = dram__bytes_read.sum + dram__bytes_write.sum
of that launch, in bytes.
"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "sm__icc_request_hit_rate.pct",
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, out_summary, out_traffic):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {c: i for i, c in enumerate(hdr)}
    summary, traffic = {}, {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "fwd" if "l96_fwd" in name else "energy" if "l96_energy" in name else "bwd" if "l96_bwd" in name else None
        if key is None or key in summary:
            continue
        d = {}
        for k in KEEP:
            if k in col:
                d[k] = [float(r[col[k]]), units[col[k]]]
        stalls = {c[len("smsp__pcsamp_warps_issue_stalled_"):]: int(r[i]) for c, i in col.items()
                  if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued")}
        tot = max(sum(stalls.values()), 1)
        d["stall_sample_share_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        summary[key] = d
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[k]]) * UNIT_BYTES[units[col[k]]]
        traffic[key] = b
    json.dump(summary, open(out_summary, "w"), indent=1)
    json.dump(traffic, open(out_traffic, "w"), indent=1)
    print(json.dumps({k: {"ms": v["gpu__time_duration.sum"][0], "traffic_GB": traffic[k] / 1e9} for k, v in summary.items()}))


if __name__ == "__main__":
    main(*sys.argv[1:4])
