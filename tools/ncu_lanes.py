"""Active lanes per source line from an ncu report captured with --import-source on: where do the idle lanes of a kernel sit?
Usage: python tools/ncu_lanes.py <report.ncu-rep> [top]"""
import collections, csv, subprocess, sys

def main(rep, top=45):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, hdr = None, None
    inst, tinst, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if r and r[0] and r[0].isdigit() and hdr:
            d = dict(zip(hdr, r))
            k = (cur, int(r[0]))
            text[k] = r[1].strip()[:90]
            try:
                inst[k] += int(d.get("Instructions Executed") or 0)
                tinst[k] += int(d.get("Thread Instructions Executed") or 0)
            except ValueError:
                pass
    ti, tt = sum(inst.values()) or 1, sum(tinst.values())
    print(f"warp instructions {ti}, thread instructions {tt}, average active lanes {tt / ti:.2f}")
    lost = {k: inst[k] * 32 - tinst[k] for k in inst}
    tl = sum(lost.values()) or 1
    print("--- lines by idle lane-slots (share of all idle slots, share of instructions, active lanes)")
    for k, v in sorted(lost.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v / tl * 100:5.1f}%  inst {inst[k] / ti * 100:5.1f}%  lanes {tinst[k] / max(inst[k], 1):5.1f}  {k[0]}:{k[1]}  {text[k]}")
    byfile = collections.defaultdict(lambda: [0, 0])
    for k in inst:
        byfile[k[0]][0] += inst[k]; byfile[k[0]][1] += tinst[k]
    print("--- by file:", {f: (round(v[0] / ti * 100, 1), round(v[1] / max(v[0], 1), 1)) for f, v in byfile.items()})
    # by line ranges of the solver / geometry (coarse regions)
    print("--- histogram of instructions by active lanes:")
    hist = collections.Counter()
    for k in inst:
        if inst[k]: hist[int(tinst[k] / inst[k]) // 4 * 4] += inst[k]
    for b in sorted(hist): print(f"   lanes {b:2d}-{b + 3:2d}: {hist[b] / ti * 100:5.1f}% of instructions")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
