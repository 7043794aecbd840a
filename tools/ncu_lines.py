"""Instruction counts / stall samples per SOURCE line of an .ncu-rep captured with --import-source on (no GPU needed):
    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [top_n]
All captured launches are summed."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def I(x):
    try:
        return int(x)
    except ValueError:
        return 0
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(out)):
    if r and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {}
        for i, n in enumerate(r):
            hdr.setdefault(n, i)
    elif hdr and len(r) > 8 and r[0].isdigit():
        a = agg.setdefault((cur, int(r[0])), [r[1][:110], 0, 0])
        a[1] += I(r[hdr["Instructions Executed"]]); a[2] += I(r[hdr["# Samples"]])
tot = sum(a[1] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print(f"{tot} warp instructions, {ts} samples")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{100 * a[1] / tot:5.1f}% inst {100 * a[2] / ts:5.1f}% samp  {f}:{l}  {a[0]}")
