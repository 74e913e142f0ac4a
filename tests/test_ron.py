"""Scene ingestion (moleengine_b200/ron.py): the RON reader and the recipes of examples/sandbox/recipes.rs, on small
scenes written here; the reference's own scene files are checked against the committed fixtures when the reference tree is
present (build container), and the oracle must still reproduce the 60-tick states stored with them."""
import glob
import math
import os

import numpy as np
import pytest

from moleengine_b200 import abi, ron

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_SCENES = "/root/reference/examples/sandbox/scenes"


def test_ron_reader():
    doc = ron.parse("""( // comment
        gravity: (0.0, -1.5), recipes: [ Block (( width: 2, pose: ( position: (1, -2.5), rotation: Deg(90) ), is_static: true, )),
        Ball (( radius: 0.5 )), Thing, Other(1, 2e1), ], name: "x" )""")
    assert doc["gravity"] == [0.0, -1.5] and doc["name"] == "x"
    blk = doc["recipes"][0]
    assert blk.name == "Block" and blk.args[0]["width"] == 2.0 and blk.args[0]["is_static"] is True
    assert blk.args[0]["pose"]["rotation"].name == "Deg" and blk.args[0]["pose"]["rotation"].args == [90.0]
    assert doc["recipes"][2].name == "Thing" and doc["recipes"][2].args is None and doc["recipes"][3].args == [1.0, 20.0]


def test_block_ball_capsule_recipes():
    sc = ron.load_scene_text("""( recipes: [
        Block (( width: 4, height: 0.2, pose: ( position: (0, -1) ), is_static: true )),
        Block (( width: 1, height: 2, radius: 0.1, pose: ( position: (0.5, 1), rotation: Deg(30) ) )),
        Ball (( radius: 0.5, position: (3, 2), restitution: 0.7, start_velocity: (1, -2) )),
        Capsule (( pose: ( position: (-2, 0.5) ), length: 0.8, radius: 0.15, is_static: true )),
    ] )""")
    b, c = sc["bodies"], sc["colliders"]
    assert len(b) == 2 and len(c) == 4 and sc["gravity"] == (0.0, -9.81)
    # static block: collider carries the pose, no body; rounded block: hw = w/2 - r
    assert c["body"].tolist() == [-1, 0, 1, -1]
    assert c["kind"][0] == abi.POLY_RECT and np.allclose([c["p0"][0], c["p1"][0], c["x"][0], c["y"][0]], [2.0, 0.1, 0.0, -1.0])
    assert np.allclose([c["p0"][1], c["p1"][1], c["circle_r"][1]], [0.4, 0.9, 0.1])
    assert np.allclose([b["rot_s"][0], b["rot_b"][0]], [math.cos(math.radians(15)), -math.sin(math.radians(15))], atol=1e-6)
    area = 4 * 0.4 * 0.9 + math.pi * 0.01 + 4 * (0.4 + 0.9) * 0.1                        # collider.rs:222-240
    assert np.isclose(1.0 / b["inv_mass"][0], 0.5 * area, rtol=1e-6)
    assert np.allclose([b["vx"][1], b["vy"][1]], [1, -2]) and np.isclose(c["restitution"][2], 0.7)
    assert c["kind"][3] == abi.POLY_SEGMENT and np.isclose(c["p0"][3], 0.4) and np.isclose(c["circle_r"][3], 0.15)


def test_blockchain_and_oscillator_constraints():
    sc = ron.load_scene_text("""( gravity: (0, 0), recipes: [
        Blockchain ( width: 0.2, spacing: 0.1, links: [ (0, 0), (1, 0), (1, 1) ], anchored_start: true, anchored_end: true ),
        Oscillator ( position: (0, 4), target_length: 2.0, begin_length: 3.0, compliance: 0.5 ),
    ] )""")
    b, c, k = sc["bodies"], sc["colliders"], sc["constraints"]
    assert len(b) == 4 and len(k) == 4 and sc["gravity"] == (0.0, 0.0)
    # two capsules of length dist - spacing - width, centred between the links, second one rotated by +90 deg
    assert np.allclose(c["p0"][:2], (1.0 - 0.1 - 0.2) / 2.0) and np.allclose(c["circle_r"][:2], 0.1)
    assert np.allclose([b["x"][1], b["y"][1]], [1.0, 0.5]) and np.isclose(b["rot_s"][1], math.cos(math.pi / 4), atol=1e-6)
    # anchor at link 0 (world target), link-to-link with compliance 0.015, end anchor at the last link
    assert k["target"].tolist() == [-1, 0, -1, 3]
    assert np.allclose([k["off0x"][0], k["off1x"][0], k["off1y"][0]], [-0.5, 0.0, 0.0]) and np.isclose(k["compliance"][1], 0.015)
    # (sic) the end anchor sits half a spacing beyond the link end: prev_block_offset already holds one half (recipes.rs:353-356)
    assert np.allclose([k["off0x"][2], k["off1x"][2], k["off1y"][2]], [0.55, 1.0, 1.0])
    assert np.isclose(k["distance"][3], 2.0) and np.isclose(k["compliance"][3], 0.5) and k["linear_damping"][3] == 0.0
    assert np.allclose(sorted(b["x"][2:].tolist()), [-1.5, 1.5])


def test_rope_connected_blocks():
    sc = ron.load_scene_text("""( recipes: [ RopeConnectedBlocks(
        block1: ( pose: ( position: (-2, 2) ), is_static: true ), offset1: (0.55, 0.55),
        block2: ( pose: ( position: (2, 2) ), is_static: false ), offset2: (-0.55, 0.55) ) ] )""")
    r, k = sc["ropes"], sc["constraints"]
    n = int(r["particle_count"][0])
    dist = (2 - 0.55) - (-2 + 0.55)
    assert len(r) == 1 and n == round(dist / 0.1) + 1 and np.isclose(r["spacing"][0], dist / (n - 1))
    parts = sc["rope_particles"]
    assert len(parts) == n and np.allclose(sc["bodies"]["x"][parts[[0, -1]]], [-1.45, 1.45], atol=1e-6)
    # first particle pinned to the world point (static block), last one attached to the dynamic block's corner
    assert k["target"].tolist() == [-1, 0] and np.allclose([k["off1x"][0], k["off1y"][0]], [-1.45, 2.55], atol=1e-6)
    assert np.allclose([k["off1x"][1], k["off1y"][1]], [-0.55, 0.55])
    rope_colls = sc["colliders"][sc["colliders"]["layer"] == abi.ROPE_LAYER]
    assert len(rope_colls) == n and not (rope_colls["flags"] & abi.COLL_HAS_STATIC_FRICTION).any()


def _fixtures():
    return sorted(glob.glob(os.path.join(GOLD, "ron_*.npz")))


def test_fixtures_exist():
    assert {os.path.basename(p) for p in _fixtures()} >= {"ron_test.npz", "ron_ropes.npz", "ron_compound_colliders.npz", "ron_oscillators.npz"}


@pytest.mark.skipif(not os.path.isdir(REF_SCENES), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("path", _fixtures())
def test_reference_scene_files_match_fixtures(path):
    f = np.load(path)
    sc = ron.load_scene(os.path.join(REF_SCENES, os.path.basename(path)[4:-4] + ".ron"))
    for key in ("bodies", "colliders", "constraints", "ropes", "rope_particles"):
        assert sc[key].tobytes() == f[key].tobytes(), key
    assert tuple(f["gravity"].tolist()) == tuple(sc["gravity"])


@pytest.mark.parametrize("path", _fixtures())
def test_oracle_reproduces_fixture_states(oracle, path):
    f = np.load(path)
    sc = {k: f[k] for k in ("bodies", "colliders", "constraints", "ropes", "rope_particles")}
    w = oracle.OracleWorld(tuning=abi.default_tuning(substeps=10))
    w.load_scene(sc)
    ff = abi.gravity_field(*f["gravity"].tolist())
    for _ in range(60):
        w.tick(mode=oracle.MODE_REFERENCE, forcefield=ff)
    assert np.allclose(w.get_bodies(), f["oracle_bodies_60"], rtol=0, atol=1e-9)
