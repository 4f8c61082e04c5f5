"""Turn the ncu captures brought back in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarise_profiles.py

Writes, per capture, a CSV of the metrics the roofline is argued from, and profiles/ncu_traffic.json
(per-config DRAM bytes of the dominant kernel, read by bench.py for `roofline.traffic`). Needs the
`ncu` CLI (no GPU): it only imports the .ncu-rep files.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__cluster_dim_x", "launch__cluster_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

# capture file -> (summary name, config key for ncu_traffic.json or None, note)
CAPTURES = {
    # round 2 (the code at HEAD): scripts/profile_pass.sh
    "r02f_prof_c3.ncu-rep": ("r02_ncu_full_c3_pair_main", "c3", "python bench.py --steps 2 --warmup 2 ...; CTA-pair (cta_group::2) streaming filter, main pass (after the sample prepass)"),
    "r02_prof_c3_eighth.ncu-rep": ("r02_ncu_full_c3_eighth_pair_main", None, "python bench.py --rows 1250000 ...: one 8-GPU shard of C3 on one GPU; CTA-pair streaming filter, main pass"),
    "r02f_prof_c4s.ncu-rep": ("r02_ncu_full_c4s_rq_main", "c4s", "python bench.py --config c4s ...; resident-query filter, main pass"),
    "r02f_prof_c5_8.ncu-rep": ("r02_ncu_full_c5_8", "c5_8", "python bench.py --config c5_8 ...; one-CTA streaming filter, batch 8, main pass"),
    "r2i_c2_rq.ncu-rep": ("r02_ncu_full_c2_rq_main", "c2", "python bench.py --config c2 ...; resident-query filter, main pass"),
    "r02f_prof_c1_direct.ncu-rep": ("r02_ncu_full_c1_direct", None, "python scripts/ubench/c1_one.py (100k x 128, one query, L2 k=10); single-launch direct scan, warm L2 (--cache-control none)"),
    # round 2, experiments that decided the design (before the sample prepass became the default for wide rows)
    "r2b_eighth_pair.ncu-rep": ("r02_ncu_full_c3_eighth_pair_noprepass", None, "one 8-GPU shard of C3, CTA-pair kernel, adaptive thresholds only (FENIX_TC_PRE_WIDE=0)"),
    "r2b_eighth_one.ncu-rep": ("r02_ncu_full_c3_eighth_onecta_noprepass", None, "one 8-GPU shard of C3, one-CTA kernel (FENIX_TC_PAIR=0), adaptive thresholds only"),
    "r2b_c3_pair.ncu-rep": ("r02_ncu_full_c3_pair_noprepass", None, "C3, CTA-pair kernel, adaptive thresholds only (FENIX_TC_PRE_WIDE=0)"),
}


def raw_page(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2]


def main():
    traffic = {}
    for fname, (name, cfg, note) in CAPTURES.items():
        path = os.path.join(ROOT, "gpurun_out", fname)
        if not os.path.exists(path):
            print("missing", fname, file=sys.stderr)
            continue
        hdr, units, vals = raw_page(path)
        col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        with open(os.path.join(OUT, name + "_summary.csv"), "w") as f:
            f.write(f"# {note}\n# kernel: {col['Kernel Name'][1]}\n")
            for m in METRICS:
                if m in col:
                    f.write(f"{m},{col[m][0]},{col[m][1]}\n")
        if cfg:
            rd = float(col["dram__bytes_read.sum"][1]) * UNIT[col["dram__bytes_read.sum"][0]]
            wr = float(col["dram__bytes_write.sum"][1]) * UNIT[col["dram__bytes_write.sum"][0]]
            traffic[cfg] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                            "kernel": col["Kernel Name"][1], "capture": name + "_summary.csv"}
    for _b in (1, 2, 4, 16, 32, 64):        # the same kernel and the same corpus pass at every small batch
        if "c5_8" in traffic:
            traffic.setdefault(f"c5_{_b}", dict(traffic["c5_8"], note="capture taken at batch 8"))
    with open(os.path.join(OUT, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
