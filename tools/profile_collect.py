#!/usr/bin/env python
"""After tools/profile_round.sh: turn gpurun_out/<tag>_*.ncu-rep / *_launches.csv into the committed evidence under
profiles/ -- launch-list table, `--set full` summaries, and profiles/traffic.json (DRAM bytes per launch of the
dominant kernels, stamped with the sha of the kernel sources they were captured on; bench.py refuses a stale stamp).
    python tools/profile_collect.py r2"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    cmdline = ("ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:roi_|mask_|score_|cim_|nchw|pairs' "
               "-s 150 -c 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also ''   (cfg2: 8 images x 2000 "
               "proposals; the window covers the end of the device-timed loops and the e2e loop)")
    lc = os.path.join(go, f"{tag}_launches.csv")
    if os.path.exists(lc):
        open(os.path.join(pr, f"{tag}_launches.txt"), "w").write(run("python", "tools/launch_summary.py", lc, cmdline))
        subprocess.run(["cp", lc, os.path.join(pr, f"{tag}_launches.csv")])
    li = os.path.join(go, f"{tag}_launches_infer.csv")
    if os.path.exists(li):
        open(os.path.join(pr, f"{tag}_launches_infer.txt"), "w").write(run(
            "python", "tools/launch_summary.py", li,
            "ncu --metrics gpu__time_duration.sum --clock-control none python tools/infer_bench.py --images 8 --iters 2   "
            "(cfg5: HRNet-W48, 8 images x 4000 proposals, 80 classes, forward only)"))
    ln = os.path.join(go, f"{tag}_launches_nvtx.csv")
    if os.path.exists(ln):
        subprocess.run(["cp", ln, os.path.join(pr, f"{tag}_launches_nvtx.csv")])
    reps = sorted(f for f in os.listdir(go) if f.startswith(f"{tag}_full_") and f.endswith(".ncu-rep"))
    traffic = {}
    for rep in reps:
        path = os.path.join(go, rep)
        name = rep[len(tag) + 6:-8]
        open(os.path.join(pr, f"{tag}_ncu_{name}.txt"), "w").write(run("python", "tools/ncu_summary.py", path))
        import csv
        import io
        rows = list(csv.reader(io.StringIO(run("ncu", "-i", path, "--page", "raw", "--csv"))))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        vals = dict(zip(hdr, rows[2]))
        units = dict(zip(hdr, rows[1]))

        def byts(key):
            v, u = float(vals[key].replace(",", "")), units[key]
            return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1))
        traffic[name] = {"kernel": vals.get("Kernel Name", "")[:80], "dram_read": byts("dram__bytes_read.sum"),
                         "dram_write": byts("dram__bytes_write.sum"),
                         "duration_us": float(vals["gpu__time_duration.sum"].replace(",", "")) *
                         {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(
                             units["gpu__time_duration.sum"], 1)}
    if traffic:
        import bench
        out = {"_comment": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from `ncu --set full "
                           "--clock-control none` captures (tools/profile_round.sh); written by tools/profile_collect.py, "
                           "never by hand.  bench.py copies the dominant stage's entry into roofline.traffic only while "
                           "kernel_source_sha16 matches the sources it runs.",
               "workload": "cfg2_r50_voc_8x2000", "kernel_source_sha16": bench.source_sha16(), "captures": traffic}
        stage_of = {"roi_align_fwd_tile": "roi_align_fwd", "roi_align_bwd_tile": "roi_align_bwd",
                    "mask_overlap_tc": "mask_overlap"}
        for k, stage in stage_of.items():
            if k in traffic:
                out[stage] = {"kernel": traffic[k]["kernel"], "dram_read": traffic[k]["dram_read"],
                              "dram_write": traffic[k]["dram_write"]}
        json.dump(out, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
    print("collected:", [f for f in os.listdir(pr) if f.startswith(tag)])
