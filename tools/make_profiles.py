#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ pass of tools/gpu_check.sh into the tracked evidence under profiles/:
summaries of every ncu --set full capture, the launch list, the bench line, and profiles/traffic.json
(measured DRAM bytes per launch of each kernel class, read by bench.py for roofline.traffic).

    python tools/make_profiles.py gpurun_out/c2 r01_v4 [frames per launch of the captures, default 16] [output dir, default profiles/]
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
        "msecond": 1e3}
# launches per step of each captured CubePad site in the benchmark chain (pipeline.resnet50_cubepad_sites(256) + 2048x8)
SITE_COUNT = {"3_256_3": 1, "64_128_1": 1, "64_64_1": 3, "128_64_1": 1, "128_32_1": 3, "256_32_1": 1, "256_16_1": 5,
              "512_16_1": 1, "512_8_1": 2, "2048_8_1": 1,
              "3_224_3": 1, "64_112_1": 1, "64_56_1": 3, "128_56_1": 1, "128_28_1": 3, "256_28_1": 1, "256_14_1": 5,
              "512_14_1": 1, "512_7_1": 2, "2048_7_1": 1, "2000_7_1": 0, "4000_7_1": 0}


def raw(path):
    """Records of a report: from the .ncu-rep, or from `ncu -i x.ncu-rep --page raw --csv > x.raw.csv` made on the box
    (reports with embedded cubins are too large to bring back in numbers)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [({h: r[i] for i, h in enumerate(hdr)}, {h: units[i] for i, h in enumerate(hdr)}) for r in rows[2:]]


def num(rec, units, key):
    return float(rec[key].replace(",", "")) * UNIT.get(units[key], 1.0)


def main(src, tag, batch=16, prof=None):
    prof = prof or os.path.join(ROOT, "profiles")
    lines, classes = [], {}
    records = []                                   # (file, site, record, units)
    have = set(os.listdir(src))
    for fn in sorted(have):
        if fn.endswith(".raw.csv"):
            if fn.replace(".raw.csv", ".ncu-rep") in have:
                continue
        elif not fn.endswith(".ncu-rep"):
            continue
        recs = raw(os.path.join(src, fn))
        fn = fn.replace(".raw.csv", ".ncu-rep")
        order = os.path.join(src, fn.replace(".ncu-rep", ".order"))
        if os.path.exists(order):                  # tools/prof_all.py: one record per spec, in this order
            sites = open(order).read().split()
            assert len(sites) == len(recs), "%s: %d records for %d specs" % (fn, len(recs), len(sites))
            records += [(fn, st.replace("cubepad_", ""), r, u) for st, (r, u) in zip(sites, recs)]
        else:
            records += [(fn, fn.replace("full_cubepad_", "").replace(".ncu-rep", ""), r, u) for r, u in recs]
    seen = set()
    for fn, site, rec, units in records:
        if True:
            name = rec["Kernel Name"]
            if (site, name) in seen:               # the same site captured twice (e.g. once more with source import)
                continue
            seen.add((site, name))
            lines.append("### %s [%s]  ::  %s" % (fn, site, name[:90]))
            for k in KEYS:
                if k in rec:
                    lines.append("  %-62s %16s %s" % (k, rec[k], units[k]))
            st = sorted(((float(v), h) for h, v in rec.items() if h.startswith("smsp__average_warps_issue_stalled_")
                         and h.endswith("_per_issue_active.ratio") and v), reverse=True)[:5]
            lines.append("  top stalls: " + ", ".join("%s %.2f" % (h.split("stalled_")[1].split("_per_")[0], v) for v, h in st))
            dram = num(rec, units, "dram__bytes_read.sum") + num(rec, units, "dram__bytes_write.sum")
            us = num(rec, units, "gpu__time_duration.sum")
            lines.append("  => DRAM read+write %.1f MB in %.1f us = %.0f GB/s" % (dram / 1e6, us, dram / us / 1e3))
            cls = ("cubepad_row_kernel" if "row_kernel" in name else "cubepad_cube2_kernel" if "cube2" in name else
                   "e2c_kernel" if "e2c" in name else "c2e_max_kernel" if "c2e" in name else name)
            n = SITE_COUNT.get(site, 1)
            if n == 0:                              # captured for the record, not a launch of the benchmark chain
                continue
            c = classes.setdefault(cls, [0.0, 0])
            c[0] += dram * n
            c[1] += n
    with open(os.path.join(prof, "%s_ncu_full.txt" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    traffic = {k: int(v[0] / v[1]) for k, v in classes.items()}
    traffic["_frames_per_launch"] = batch
    traffic["_cube"] = int(os.environ.get("CP360_PROF_CUBE", "256"))
    traffic["_source"] = "ncu --set full, B=%d, profiles/%s_ncu_full.txt" % (batch, tag)
    traffic["_note"] = ("DRAM bytes (read+write) per launch from ncu --set full, averaged over the launches of the class in "
                        "one chain step at B=%d; isolated captures leave part of the output dirty in the 126 MB L2 at kernel "
                        "end, so writes are under-counted for sites whose output is < L2 (source: profiles/%s_ncu_full.txt)" % (batch, tag))
    if os.environ.get("CP360_PROF_TRAFFIC", "1") != "0":
        with open(os.path.join(prof, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    for a, b in (("bench.json", "%s_bench.json"), ("bench.err", "%s_bench_sites.txt"), ("kbench.txt", "%s_kbench.txt"),
                 ("launches.csv", "%s_launches.csv"), ("bench_reference.json", "%s_bench_reference.json"),
                 ("pytest_gpu.log", "%s_pytest_gpu.log"), ("smoke.log", "%s_smoke.log")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(prof, b % tag))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 16, sys.argv[4] if len(sys.argv) > 4 else None)
