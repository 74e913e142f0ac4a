"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the header declares, the numpy
record dtypes are byte-identical to the C structs, and nothing computes without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from moleengine_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "starframe_gpu.h")


def test_exports_every_declared_symbol():
    lib = abi.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(sfg_[a-z_0-9]+)\s*\(", text))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sfg_abi_version() == 1


def test_record_sizes_match_sizeof(tmp_path):
    src = tmp_path / "sz.c"
    names = ["SfgTuning", "SfgBody", "SfgCollider", "SfgConstraint", "SfgRope", "SfgForceTerm", "SfgForceField", "SfgContactInfo", "SfgRay",
             "SfgCastHit", "SfgStats"]
    src.write_text('#include <stdio.h>\n#include "starframe_gpu.h"\nint main(){' + "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    dts = [abi.tuning_dtype, abi.body_dtype, abi.collider_dtype, abi.constraint_dtype, abi.rope_dtype, abi.force_term_dtype, abi.forcefield_dtype,
           abi.contact_dtype, abi.ray_dtype, abi.hit_dtype, abi.stats_dtype]
    assert sizes == [d.itemsize for d in dts]


def test_defaults_match_reference():
    lib = abi.load()
    t = np.zeros(1, abi.tuning_dtype)
    lib.sfg_default_tuning(abi.ptr(t))
    assert (t["substeps"][0], t["fall_asleep_frames"][0]) == (10, 10)                   # physics.rs:338-349
    assert abs(t["sleep_vel_threshold"][0] - 0.001) < 1e-9 and t["max_expected_acceleration"][0] == 10.0
    m = np.zeros(64, np.uint64)
    lib.sfg_default_mask_matrix(abi.ptr(m))
    assert np.array_equal(m, abi.default_mask_matrix())                                 # collision.rs:132-140
    assert (int(m[63]) >> 63) & 1 == 0 and (int(m[63]) >> 62) & 1 == 1 and int(m[0]) == 0xFFFFFFFFFFFFFFFF


def test_no_cpu_fallback():
    """Without a CUDA device world creation reports SFG_E_NO_DEVICE; nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = abi.load()
    h = C.c_void_p()
    rc = lib.sfg_world_create(None, None, 0, C.byref(h))
    assert rc in (abi.SFG_E_NO_DEVICE, abi.SFG_E_CUDA) and not h.value
    out = np.zeros(14, np.float32)
    z = np.zeros(8, np.float32)
    assert lib.sfg_debug_intersection_check(0, 1, abi.ptr(z), abi.ptr(z), abi.ptr(out)) != abi.SFG_OK


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "moleengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "sf_world.hpp" not in text, f
