"""Per-CTA lateness in an SFG_TRACE dump: is it always the same CTAs (placement) or whoever drew the last heavy item (granularity)?
Usage: python tools/trace_ctas.py <trace.bin>"""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
n_sub, ns, grid, _ = np.frombuffer(raw[:16], np.uint32)
stages = np.frombuffer(raw[16:16 + 16 * ns], np.uint32).reshape(ns, 4)
t = np.frombuffer(raw[16 + 16 * ns:], np.uint64).reshape(n_sub, ns, grid, 2).astype(np.int64)
late = np.zeros(grid)
cnt = 0
for s in range(n_sub):
    for k in range(ns):
        arr = t[s, k, :, 0]
        if arr.max() == 0 or stages[k, 0] != 4 or stages[k, 2] - stages[k, 1] < 300000:
            continue
        off = (arr - arr.min()) / 1e3
        late += off
        cnt += 1
        if s == 1:
            q = np.percentile(off, [10, 50, 90, 99, 100])
            print(f"stage {k} units {stages[k, 2] - stages[k, 1]}: arrival offsets us p10 {q[0]:.1f} p50 {q[1]:.1f} p90 {q[2]:.1f} p99 {q[3]:.1f} max {q[4]:.1f}; latest CTAs {np.argsort(off)[-6:][::-1].tolist()}")
late /= max(cnt, 1)
order = np.argsort(late)
print("mean lateness per CTA over", cnt, "big contact stages (us): min", late.min().round(2), "median", np.median(late).round(2), "max", late.max().round(2))
print("always-late CTAs:", [(int(i), round(float(late[i]), 1)) for i in order[-12:][::-1]])
print("always-early CTAs:", [(int(i), round(float(late[i]), 1)) for i in order[:12]])
