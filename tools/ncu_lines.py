"""Per-source-line stall samples and executed warp instructions from an ncu report captured with --import-source on."""
import collections, csv, subprocess, sys

def main(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, hdr = None, None
    samp, inst, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if r and r[0] and r[0].isdigit() and hdr:
            d = dict(zip(hdr, r))
            k = (cur, int(r[0]))
            text[k] = r[1].strip()[:100]
            try:
                samp[k] += int(r[4] or 0)
                inst[k] += int(d.get("Instructions Executed") or 0)
            except ValueError:
                pass
    ts, ti = sum(samp.values()) or 1, sum(inst.values()) or 1
    print(f"total stall samples {ts}, total warp instructions {ti}")
    print("--- by stall samples")
    for k, v in samp.most_common(top):
        print(f"{v / ts * 100:5.1f}%  inst {inst[k] / ti * 100:5.1f}%  {k[0]}:{k[1]}  {text[k]}")
    print("--- by executed instructions")
    for k, v in inst.most_common(top):
        print(f"{v / ti * 100:5.1f}%  samp {samp[k] / ts * 100:5.1f}%  {k[0]}:{k[1]}  {text[k]}")
    byfile_s, byfile_i = collections.Counter(), collections.Counter()
    for k, v in samp.items(): byfile_s[k[0]] += v
    for k, v in inst.items(): byfile_i[k[0]] += v
    print("--- by file:", {f: (round(byfile_s[f] / ts * 100, 1), round(byfile_i[f] / ti * 100, 1)) for f in byfile_s})

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
