"""Per-source-line totals (warp instructions executed, stall samples) of an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py <file.ncu-rep> [top-N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; recs = []; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        recs.append((cur, int(r[0]), r[1].strip(), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])))
tot_i = sum(x[4] for x in recs); tot_s = sum(x[3] for x in recs)
print(f"# total warp-instructions {tot_i}, stall samples {tot_s}")
for f, ln, src, s, i in sorted(recs, key=lambda x: -x[4])[:top]:
    print(f"{100*i/tot_i:5.1f}% instr {100*s/max(tot_s,1):5.1f}% stall  {f}:{ln}  {src[:110]}")
