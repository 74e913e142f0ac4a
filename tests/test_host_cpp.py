"""The C++ host mirror of the reference API (moleengine_b200/csrc/host/physics_world.hpp) compiled against the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "sandbox_scene"
    lib_dir = os.path.join(ROOT, "moleengine_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", str(out), os.path.join(ROOT, "tests", "host_cpp", "sandbox_scene.cpp"),
                           "-L" + lib_dir, "-lstarframe_gpu", "-Wl,-rpath," + lib_dir])
    return str(out)


def test_builds_and_fails_loudly_without_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_sandbox_scene_runs(exe):
    r = subprocess.run([exe, "40"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    head = dict(zip(lines[0].split()[::2], map(int, lines[0].split()[1::2])))
    # 24 boxes + ball + 2 rope blocks + 20 particles, minus one particle and one block removed mid-run;
    # removing a particle split the rope in two; the constraint to the removed block died (physics.rs:423-428)
    assert head == {"bodies": 24 + 1 + 2 + 20 - 2, "colliders": 1 + 24 + 1 + 2 + 20 - 2, "ropes": 2, "constraints": 1}
    bodies = {int(l.split()[1]): list(map(float, l.split()[2:])) for l in lines if l.startswith("body")}
    assert all(v[1] > -4.6 for v in bodies.values())          # nothing fell through the ground
    ray = [l for l in lines if l.startswith("ray")][0].split()
    assert int(ray[-1]) == 1 and 0.0 < float(ray[2]) < 12.0   # the ray from above hits the ball resting on the stack
    assert int([l for l in lines if l.startswith("ball_contacts")][0].split()[1]) >= 1
    assert int([l for l in lines if l.startswith("query_point_hits")][0].split()[1]) >= 1
    assert int([l for l in lines if l.startswith("query_shape_hits")][0].split()[1]) >= 1


@pytest.fixture(scope="module")
def hecs_exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "hecs_sync_test"
    lib_dir = os.path.join(ROOT, "moleengine_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", str(out), os.path.join(ROOT, "tests", "host_cpp", "hecs_sync_test.cpp"),
                           "-L" + lib_dir, "-lstarframe_gpu", "-Wl,-rpath," + lib_dir])
    return str(out)


def test_hecs_sync_rules_on_the_host(hecs_exe):
    """HecsSyncManager mirror vs hecs_sync.rs:121-227: presets, both directions, auto-register, auto-delete, collider entities."""
    r = subprocess.run([hecs_exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
def test_physics_tick_with_hecs_sync(hecs_exe):
    """Game::physics_tick (game.rs:297-303) through the mirror on the GPU: sync in, tick, sync out; teleport; despawn."""
    r = subprocess.run([hecs_exe, "tick"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
