"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/ncu_summarise.py <tag> <launches.csv> <prof.ncu-rep>
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "launch__grid_size", "launch__block_size"]


def launches(path, out_md):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out_md, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write("command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (10 M points, 3 radii, 1 B200)\n\n")
        f.write("total %.2f ms over %d launches\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n"
                % (tot, sum(v[0] for v in agg.values())))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k[:90], v[0], v[1], 100 * v[1] / tot))


def full(path, out_csv, out_json):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    idx = [(m, H.index(m)) for m in ["Kernel Name"] + METRICS if m in H]
    traffic, calls = {}, {}
    with open(out_csv, "w") as f:
        w = csv.writer(f)
        w.writerow([m for m, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split("(")[0] if m == "Kernel Name" else r[i] for m, i in idx])
            name = r[H.index("Kernel Name")]
            # one dk/dn CALL is three kernels (screening walk + the two refinement passes): their DRAM bytes are added
            # up and divided by the number of calls (= screening launches); the one-pass kernel is a call of its own
            key = "dkdn" if ("dkdn" in name or "dk_refine" in name) else ("moments" if "moments" in name else None)
            if key:
                def tobytes(col):
                    v, u = float(r[H.index(col)].replace(",", "")), units[H.index(col)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic.setdefault(key, []).append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
                if key == "moments" or "dkdn_screen" in name or "dkdn_kernel" in name:
                    calls[key] = calls.get(key, 0) + 1
    json.dump({k: sum(v) / max(calls.get(k, 1), 1) for k, v in traffic.items()}, open(out_json, "w"))


if __name__ == "__main__":
    tag, lpath, ppath = sys.argv[1:4]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(lpath, os.path.join(ROOT, "profiles", tag + "_launches.md"))
    full(ppath, os.path.join(ROOT, "profiles", tag + "_kernels.csv"), os.path.join(ROOT, "profiles", "traffic.json"))
