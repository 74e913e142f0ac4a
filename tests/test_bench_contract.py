"""bench.py's reference arm runs on the CPU, so its JSON contract can be checked without a GPU (the driver runs
`bench.py --impl reference` first and divides our arm's numbers by its value)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "tiny", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "body-substeps/sec" and d["unit"] == "body-substeps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] >= 3 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms describe the same workload: the driver compares `config` of the two JSON lines
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.CONFIG and "workload" in d["config"]


def test_solver_bytes_follow_the_survey_formula():
    """192 N_b + 64 N_c + 96 N_k + 64 N_r + 152 per touching entry (device-counted) + 52 per rejected unit, per substep."""
    sys.path.insert(0, ROOT)
    import bench
    st = {"n_pairs": 1000, "n_touching_units": 4 * 300, "n_touching_entries": 4 * 500, "n_colliders": 120, "n_constraint_entries": 10, "n_rope_particles": 7, "n_contact_entries": 1600}
    b, counts = bench.solver_bytes_per_tick(st, 100, 4)
    assert b == 4 * (192 * 100 + 64 * 120 + 96 * 10 + 64 * 7 + 152 * 500 + 52 * 700)
    assert counts["touching_entries_per_substep"] == 500 and counts["pair_units"] == 1000


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "tiny", "--steps", "1"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
