"""Decode an SFG_TRACE dump: per stage, time until the last CTA arrived at the barrier and barrier latency after that."""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
n_sub, ns, grid, _ = np.frombuffer(raw[:16], np.uint32)
stages = np.frombuffer(raw[16:16 + 16 * ns], np.uint32).reshape(ns, 4)
t = np.frombuffer(raw[16 + 16 * ns:], np.uint64).reshape(n_sub, ns, grid, 2).astype(np.int64)
names = ["integrate", "rope_step", "constraints", "precontact", "contacts", "derive", "contact_vel", "constraint_damp", "rope_damp", "rope_lateral", "fold"]
t0 = t[0, 0, :, 0].min()
prev_release = None
tot = {}
print(f"substeps {n_sub} stages {ns} grid {grid}; total {(t[-1, -1, :, 1].max() - t0) / 1e3:.1f} us")
print(f"{'stage':16s} {'units':>8s} {'first_arr':>9s} {'last_arr':>9s} {'release':>9s}  (us since previous release; substep 1)")
for s in range(n_sub):
    for k in range(ns):
        arr, rel = t[s, k, :, 0], t[s, k, :, 1]
        if arr.max() == 0:      # stage skipped (empty velocity worklist)
            continue
        start = prev_release if prev_release is not None else t0
        first, last, release = (arr.min() - start) / 1e3, (arr.max() - start) / 1e3, (rel.max() - start) / 1e3
        prev_release = rel.max()
        nm = names[stages[k, 0]]
        d = tot.setdefault(nm, [0.0, 0.0, 0.0, 0])
        d[0] += first; d[1] += last - first; d[2] += release - last; d[3] += 1
        if s == 1 or n_sub == 1:
            print(f"{nm:16s} {stages[k, 2] - stages[k, 1]:8d} {first:9.2f} {last:9.2f} {release:9.2f}")
print("totals (us): stage, n, until-first-arrival, first->last arrival (imbalance), last arrival->release (barrier)")
for nm, d in tot.items():
    print(f"  {nm:16s} {d[3]:4d} {d[0]:9.1f} {d[1]:9.1f} {d[2]:9.1f}")
