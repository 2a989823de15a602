"""bench.py's output contract on a CPU-only box: the reference arm prints exactly ONE JSON line with the
keys the driver reads; the B200 arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(argv), capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-frames", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    # "reference" = the unmodified reference (live checkout or the sha256-pinned oracle/_ref staging copy), else the port
    from oracle import ref_loader
    want_kind = "reference" if ref_loader.available() and (ref_loader.kind() == "live" or ref_loader.verify()) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_runs_from_the_staged_copy(tmp_path):
    """What the GPU box does: /root/reference is absent there, the arm must run the staged, verified copy."""
    from oracle import ref_loader
    if not ref_loader.verify():
        pytest.skip("oracle/_ref not staged (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, CP360_REFERENCE=str(tmp_path / "absent"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-frames", "1", "--workload", "clstm", "--clstm-variant", "reference"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert line["cpu_baseline"]["kind"] == "reference" and "oracle/_ref" in line["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_b200_arm_has_no_cpu_fallback():
    r = run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(ln.lstrip().startswith("{") for ln in r.stdout.splitlines())


def test_both_arms_name_the_same_workload():
    """The driver compares the `config.workload` strings of the two arms; the first-site note must follow the flag in both."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for wl in ("chain", "corpus", "clstm"):
        for two in (False, True):
            a = argparse.Namespace(workload=wl, cube=256, clstm_variant="baseline", no_fuse_first_site=two)
            s = bench.workload_string(a)
            assert ("one kernel" in s) == (wl != "clstm" and not two), (wl, two, s)
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-frames", "1")
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert line["config"]["workload"] == bench.workload_string(
        argparse.Namespace(workload="chain", cube=256, clstm_variant="baseline", no_fuse_first_site=False))
