#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ pass of tools/gpu_r2_profile.sh into the tracked round-2 evidence under profiles/.

    python tools/make_profiles_r02.py gpurun_out/r2prof [r02]
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_profiles  # noqa: E402


def main(src, tag="r02"):
    prof = os.path.join(ROOT, "profiles")
    # ncu --set full summaries: one pass per spec set (256 family = the headline chain, 224 family, section-8(f) rows)
    for cube, name, traffic in (("224", "224", "0"), ("0", "f_rows", "0"), ("256", "256", "1")):
        raw = os.path.join(src, "full_%s.raw.csv" % cube)
        order = os.path.join(src, "full_%s.order" % cube)
        if not (os.path.exists(raw) and os.path.exists(order)):
            print("missing", raw)
            continue
        tmp = os.path.join(src, "_mp_%s" % cube)
        os.makedirs(tmp, exist_ok=True)
        shutil.copy(raw, os.path.join(tmp, "full_all.raw.csv"))
        shutil.copy(order, os.path.join(tmp, "full_all.order"))
        os.environ["CP360_PROF_CUBE"] = cube if cube != "0" else "256"
        os.environ["CP360_PROF_TRAFFIC"] = traffic
        make_profiles.main(tmp, "%s_%s" % (tag, name), 32, prof)
        shutil.rmtree(tmp)
    copies = [("bench_256.json", "%s_bench_256.json"), ("bench_256.err", "%s_bench_256_sites.txt"),
              ("bench_224.json", "%s_bench_224.json"), ("bench_224.err", "%s_bench_224_sites.txt"),
              ("bench_reference.json", "%s_bench_reference.json"), ("batch_curve.txt", "%s_batch_curve.txt"),
              ("bench_clstm.json", "%s_bench_clstm.json"), ("bench_clstm_reference_widths.json", "%s_bench_clstm_reference_widths.json"),
              ("bench_clstm_reference_widths_b1.json", "%s_bench_clstm_reference_widths_b1.json"),
              ("bench_clstm_reference_widths_b4.json", "%s_bench_clstm_reference_widths_b4.json"),
              ("bench_corpus_n1.json", "%s_bench_corpus_n1.json"),
              ("launches_256.csv", "%s_launches_256.csv"), ("launches_224.csv", "%s_launches_224.csv"),
              ("kbench.txt", "%s_kbench.txt"), ("site_sweep.txt", "%s_site_sweep.txt"),
              ("pytest_gpu.log", "%s_pytest_gpu.log"), ("smoke.log", "%s_smoke.log"), ("smi.txt", "%s_smi.txt"),
              ("summary/row_kernel_64_128_lines.txt", "%s_row_kernel_lines.txt")]
    for a, b in copies:
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(prof, b % tag))
    # sanitizer logs: the summaries (the full racecheck logs of the builds that do report hazards run to megabytes)
    lines = []
    for fn in sorted(os.listdir(src)):
        if fn.endswith(".log") and fn.split("_")[0] in ("racecheck", "memcheck", "synccheck"):
            body = open(os.path.join(src, fn), errors="replace").read().splitlines()
            races = sum("Race reported" in ln for ln in body)
            lines.append("%s: %s%s" % (fn, body[-1] if body else "(empty)", "  [%d race records]" % races if races else ""))
            out = fn.replace(".log", ".out")
            if os.path.exists(os.path.join(src, out)):
                tail = [ln for ln in open(os.path.join(src, out), errors="replace").read().splitlines() if ln.strip()][-3:]
                lines += ["    " + t[:200] for t in tail]
    if lines:
        with open(os.path.join(prof, "%s_sanitizer_summary.txt" % tag), "w") as f:
            f.write("\n".join(lines) + "\n")
    for k in ("bench_256.json", "bench_224.json"):
        p = os.path.join(src, k)
        if os.path.exists(p):
            d = json.load(open(p))
            print(k, d["value"], d["roofline"]["frac"], d["roofline"]["chain_frac"], (d.get("e2e") or {}).get("value"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r02")
