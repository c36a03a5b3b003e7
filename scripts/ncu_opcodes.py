"""Executed warp-instructions per opcode (and per item) of an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_opcodes.py <file.ncu-rep> [items]   (items: divide the counts, e.g. frames x items per frame)"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; items = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, iex = hdr.index('Source'), hdr.index('Instructions Executed')
cls = {}
for r in data:
    op = re.sub(r'^@!?U?P\d\s+', '', r[isrc].strip()).split()[0].split('.')[0]
    cls[op] = cls.get(op, 0) + int(r[iex])
tot = sum(cls.values())
print(f"# warp-instructions executed {tot} ({tot / items:.1f} per item)")
for k, s in sorted(cls.items(), key=lambda x: -x[1])[:30]:
    print(f"{k:10s} {100 * s / tot:5.1f}%  {s / items:9.1f}")
