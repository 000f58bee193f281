"""Scratch: per-kernel SASS facts of the built library (cuobjdump -sass / -res-usage): registers, 256-bit gathers / stores,
local-memory (spill) instructions, reductions, FP64 instructions.   python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import re, subprocess, sys, os
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "voronoids_b200", "libvoronoids_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
dem = subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
want = ("k_attempt_hot_tiled", "k_attempt_hot<", "k_attempt_slow", "k_attempt_coop", "k_commit_tiled", "k_commit_coop", "k_spheres")
print("# SASS summary of libvoronoids_b200.so (cuobjdump -sass, -res-usage), sm_100a\n")
print("| kernel | regs | smem B | stack B | SASS instr | LDG.E.ENL2.256 | STG.E.ENL2.256 | LDL/STL | RED/ATOM | MATCH/VOTE | F64 (D*) | F2F |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, d in sorted(zip(funcs, dem), key=lambda x: x[1]):
    if not any(w in d for w in want):
        continue
    body = funcs[name]
    c = lambda pat: sum(1 for l in body if re.search(pat, l))
    r = regs.get(name, (0, 0, 0))
    short = re.sub(r"\(.*", "", d).replace("void vor::", "")
    print(f"| `{short}` | {r[0]} | {r[1]} | {r[2]} | {len(body)} | {c(r'LDG\.E\.ENL2\.256')} | {c(r'STG\.E\.ENL2\.256')} | {c(r'\b(LDL|STL)')} | "
          f"{c(r'\b(REDG|ATOMG|ATOMS|RED\.|ATOM\.)')} | {c(r'\b(MATCH|VOTE)')} | {c(r'\b(DADD|DMUL|DFMA|DSETP|DMNMX)')} | {c(r'F2F')} |")
