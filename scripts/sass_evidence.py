"""SASS evidence for profiles/: per kernel of libphnms.so, how often the Blackwell-specific mnemonics appear (cuobjdump -sass).
    UBLKCP  = cp.async.bulk (TMA 1-D bulk copy)      LDGSTS = cp.async            SYNCS = mbarrier ops
    STAS    = st.async (remote shared-memory store)   UCGABAR = barrier.cluster    FADD2 = packed fp32x2 add (sub.f32x2)
    R2P     = register -> predicates                  REDUX / CREDUX = warp reductions
    FFMA must be 0 in every kernel that evaluates the pair predicate (the reference's fp32 sum has no FMA).
usage: python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "phnet_b200", "csrc", "libphnms.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UBLKCP", "LDGSTS", "SYNCS", "STAS", "UCGABAR", "FADD2", "FADD", "FFMA", "R2P", "REDUX", "LDS", "STG", "LDG"]
print(f"# {so}: cuobjdump -sass, instruction counts per kernel (static code, not executed counts)")
print("kernel," + ",".join(ops) + ",instructions")
for fn in re.split(r"\n\s*Function : ", sass)[1:]:
    name = fn.split("\n", 1)[0].strip()
    body = [l for l in fn.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    mn = [re.sub(r"^@!?U?P\w+\s+", "", re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l)).split()[0].split(".")[0] for l in body]
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    cnt = [sum(1 for m in mn if m == op or (op == "REDUX" and m in ("REDUX", "CREDUX")) or (op == "UCGABAR" and m.startswith("UCGABAR"))) for op in ops]
    print(demangled + "," + ",".join(str(c) for c in cnt) + f",{len(body)}")
