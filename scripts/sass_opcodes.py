"""Per-kernel histogram of the SASS mnemonics that show what the shipped library really does (TMA, bulk copies,
mbarrier, cp.async, integer dot products, packed min/max, shared / global atomics, warp reductions) -- from
`cuobjdump -sass graphics-project_b200/liblineengine.so`, no GPU needed.  Output: profiles/sass_opcodes.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphics-project_b200", "liblineengine.so")
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "IDP", "VIMNMX", "PRMT", "ATOMS", "RED", "ATOMG", "ATOM",
         "REDUX", "MATCH", "VOTE", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "MUFU", "HMMA", "UTC"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
kern, hist, total, arch = None, collections.OrderedDict(), {}, set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for wname in WATCH:
            if op == wname or op.startswith(wname + "."):
                hist[kern][wname] += 1
                break
lines = [f"# SASS mnemonics per kernel of liblineengine.so (cuobjdump -sass; arch: {', '.join(sorted(arch))}); made by scripts/sass_opcodes.py",
         "# UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, IDP = integer dot product,",
         "# VIMNMX = packed min/max, ATOMS = shared atomic, RED / ATOMG = global reduction / atomic, REDUX = warp reduce, no HMMA / UTC*MMA (no tensor cores: nothing is a contraction)", ""]
agg = collections.Counter()
for k in hist:
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)
    ops = ", ".join(f"{o} {c}" for o, c in sorted(hist[k].items(), key=lambda t: -t[1]))
    lines.append(f"{name}: {total[k]} instructions; {ops}")
    agg.update(hist[k])
lines += ["", "TOTAL: " + ", ".join(f"{o} {c}" for o, c in sorted(agg.items(), key=lambda t: -t[1]))]
open(os.path.join(ROOT, "profiles", "sass_opcodes.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
print(len(hist), "kernels")
