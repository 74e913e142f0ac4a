"""The Rust -sys crate (starframe-gpu-sys/src/lib.rs) cannot be compiled here (no cargo / rustc in this image), so it is kept
in step with the C ABI mechanically: every `#[repr(C)]` record must have the fields of the same-named struct in
include/starframe_gpu.h, in the same order with the same types, and the same size and field offsets as the numpy dtype the
tested ctypes binding uses (moleengine_b200/abi.py, itself checked against the built library's behaviour by the GPU tests);
every function the header declares must be declared `extern "C"` with the same parameter types, and vice versa; every
`#define`d constant the crate re-states must carry the header's value."""
import os
import re

import numpy as np

from moleengine_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "starframe_gpu.h")).read()
LIB_RS = open(os.path.join(ROOT, "starframe-gpu-sys", "src", "lib.rs")).read()

C2R = {"uint32_t": "u32", "int32_t": "i32", "float": "f32", "uint64_t": "u64", "double": "f64", "int": "c_int", "size_t": "usize", "uint8_t": "u8", "char": "c_char", "void": "c_void"}
R2NP = {"u32": "<u4", "i32": "<i4", "f32": "<f4", "u64": "<u8"}
DTYPES = {"SfgTuning": abi.tuning_dtype, "SfgBody": abi.body_dtype, "SfgCollider": abi.collider_dtype, "SfgConstraint": abi.constraint_dtype,
          "SfgRope": abi.rope_dtype, "SfgForceTerm": abi.force_term_dtype, "SfgForceField": abi.forcefield_dtype, "SfgContactInfo": abi.contact_dtype,
          "SfgRay": abi.ray_dtype, "SfgShapeQuery": abi.shape_query_dtype, "SfgCastHit": abi.hit_dtype, "SfgStats": abi.stats_dtype}


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def c_structs():
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", strip_comments(HEADER), flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(\w+)\s+(.*)$", decl, flags=re.S)
            ctype, names = m.group(1), m.group(2)
            for n in names.split(","):
                n = n.strip()
                arr = re.match(r"(\w+)\[(\d+)\]$", n)
                fields.append((arr.group(1), ctype, int(arr.group(2))) if arr else (n, ctype, 0))
        out[name] = fields
    return out


def rust_structs():
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\]\s*(?:#\[derive\([^\]]*\)\]\s*)?pub struct (\w+)\s*\{(.*?)\n\}", strip_comments(LIB_RS), flags=re.S):
        fields = []
        for f in body.split(","):
            f = f.strip()
            if not f:
                continue
            m = re.match(r"(?:pub\s+)?(\w+)\s*:\s*(.+)$", f, flags=re.S)
            fname, ftype = m.group(1), m.group(2).strip()
            arr = re.match(r"\[(\w+);\s*(\d+)\]$", ftype)
            fields.append((fname, arr.group(1), int(arr.group(2))) if arr else (fname, ftype, 0))
        out[name] = fields
    return out


def test_records_match_the_header_and_the_numpy_dtypes():
    cs, rs = c_structs(), rust_structs()
    assert set(DTYPES) <= set(cs) and set(DTYPES) <= set(rs), (sorted(cs), sorted(rs))
    for name, dt in DTYPES.items():
        cf, rf = cs[name], rs[name]
        assert [f[0] for f in cf] == [f[0] for f in rf], name                      # same fields, same order
        for (cn, ct, ca), (rn, rt, ra) in zip(cf, rf):
            assert C2R.get(ct, ct) == rt and ca == ra, (name, cn, ct, rt)           # same types, same array lengths
        # the same layout as the dtype the tested binding uses: names, offsets, total size (repr(C) layout computed by hand)
        assert list(dt.names) == [f[0] for f in rf], name
        off = 0
        for fname, rt, ra in rf:
            sub = DTYPES[rt] if rt in DTYPES else np.dtype(R2NP[rt])
            align = max(np.dtype(sub[n]).alignment for n in sub.names) if sub.names else sub.alignment
            off = (off + align - 1) // align * align
            assert dt.fields[fname][1] == off, (name, fname, off, dt.fields[fname][1])
            assert dt.fields[fname][0].itemsize == sub.itemsize * max(ra, 1), (name, fname)
            off += sub.itemsize * max(ra, 1)
        assert dt.itemsize >= off and dt.itemsize - off < 8, (name, dt.itemsize, off)


def c_functions():
    out = {}
    src = strip_comments(HEADER)
    src = src[src.index("int sfg_abi_version"):]
    for ret, name, args in re.findall(r"(?:^|\n)\s*((?:const\s+)?\w+\s*\*?)\s*(sfg_\w+)\s*\(([^)]*)\)\s*;", src):
        params = []
        for a in args.split(","):
            a = a.strip()
            if a in ("", "void"):
                continue
            a = re.sub(r"\[\d*\]", "*", a)
            const = a.startswith("const ")
            a = a[6:] if const else a
            m = re.match(r"(\w+)\s*(\**)\s*\w*\s*(\**)$", a)
            base, stars = m.group(1), len(m.group(2)) + len(m.group(3))
            rt = C2R.get(base, base)
            for k in range(stars):
                rt = ("*const " if (const and k == 0) else "*mut ") + rt
            params.append(rt)
        out[name] = (" ".join(ret.split()), params)
    return out


def rust_functions():
    out = {}
    block = strip_comments(LIB_RS)
    block = block[block.index('extern "C"'):]
    for name, args, ret in re.findall(r"pub fn (sfg_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()]
        out[name] = ((ret or "").strip(), params)
    return out


def test_every_function_is_declared_on_both_sides_with_the_same_parameters():
    cf, rf = c_functions(), rust_functions()
    product = {n for n in cf if not n.startswith("sfg_debug_")}                  # the parity hooks are not part of the drop-in surface
    assert product == set(rf), (sorted(product - set(rf)), sorted(set(rf) - product))
    assert set(cf) == set(abi.EXPORTS), (sorted(set(cf) - set(abi.EXPORTS)), sorted(set(abi.EXPORTS) - set(cf)))
    for name in sorted(product):
        (cret, cparams), (rret, rparams) = cf[name], rf[name]
        norm = lambda t: t.replace("*const SfgWorld", "*W").replace("*mut SfgWorld", "*W")     # noqa: E731  (constness of the opaque handle is advisory)
        assert [norm(p) for p in cparams] == [norm(p) for p in rparams], (name, cparams, rparams)
        want_ret = {"int": "c_int", "void": "", "const char*": "*const c_char", "const char *": "*const c_char"}[cret]
        assert rret == want_ret, (name, cret, rret)


def test_constants_carry_the_header_values():
    defs = {k: int(v.rstrip("uU"), 0) for k, v in re.findall(r"#define (SFG_\w+)\s+(\d+[uU]?|0x[0-9a-fA-F]+[uU]?)\b", HEADER)}
    consts = {k: int(v, 0) for k, v in re.findall(r"pub const (SFG_\w+):\s*\w+\s*=\s*(\d+|0x[0-9a-fA-F]+);", LIB_RS)}
    assert len(consts) > 30
    for k, v in consts.items():
        assert defs.get(k) == v, (k, v, defs.get(k))
    must = {k for k in defs if k.startswith(("SFG_E_", "SFG_BODY_", "SFG_COLL_", "SFG_POLY_", "SFG_LIMIT_", "SFG_FF_", "SFG_TUNE_"))} | {"SFG_OK", "SFG_ABI_VERSION", "SFG_CONSTRAINT_CAN_SLEEP"}
    assert must <= set(consts), sorted(must - set(consts))


def test_wrapper_crate_covers_the_reference_surface():
    """starframe-gpu/src/lib.rs: the PhysicsWorld-shaped wrapper names every public item of the path's boundary
    and the HecsSyncManager surface (hecs_sync.rs:62-227)."""
    src = open(os.path.join(ROOT, "starframe-gpu", "src", "lib.rs")).read()
    for item in ("pub fn new(", "pub fn clear(", "pub fn tick(", "pub fn contacts_for_collider(", "pub fn query_point(", "pub fn query_shape(", "pub fn raycast(",
                 "pub fn spherecast(", "pub consts:", "pub mask_matrix:", "pub entity_set:", "pub rope_set:", "pub constraint_set:",
                 "pub struct HecsSyncOptions", "pub fn both_ways()", "pub fn hecs_to_physics_only()", "pub fn physics_to_hecs_only()", "pub fn do_not_sync()",
                 "pub fn new_autosync(", "pub fn register_body(", "pub fn register_collider(", "pub fn get_body_entity(", "pub fn get_collider_entity(",
                 "pub fn sync_hecs_to_physics(", "pub fn sync_physics_to_hecs(", "pub fn physics_tick(", "pub default_opts:"):
        assert item in src, item
    for sym in re.findall(r"sys::(sfg_\w+)", src):
        assert sym in rust_functions(), sym
