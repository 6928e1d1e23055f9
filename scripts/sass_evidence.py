"""Counts the SASS mnemonics that prove which hardware paths each kernel uses (cuobjdump -sass of the built objects):
UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, LDGSTS = cp.async,
SYNCS = mbarrier, MUFU, multimem.  Writes profiles/r1_sass_evidence.txt.  Runs on the CPU box (no GPU needed)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = ["conv_tc.o", "linear_tc.o", "adam.o", "bce_ts.o", "stitch.o", "pool.o"]
pats = (("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
        ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("MUFU", r"\bMUFU"), ("multimem", r"MULTIMEM|\.MMEM|LDGMC|STGMC|RED\.MC"))
out = ["# SASS evidence (cuobjdump -sass of driving-dirty_b200/csrc/build/*.o, sm_100a), instruction counts per kernel:",
       "# UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, LDGSTS = cp.async,",
       "# SYNCS = mbarrier ops, MUFU = special-function unit, multimem = NVLink multicast ld_reduce / st.  scripts/sass_evidence.py", ""]
for o in objs:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "driving-dirty_b200/csrc/build", o)], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)[:110]
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for key, pat in pats:
            if re.search(pat, line):
                counts[cur][key] += 1
    out.append(f"== {o}")
    for k, c in counts.items():
        if c:
            out.append(f"  {k}\n      " + "  ".join(f"{a}={b}" for a, b in c.items()))
    out.append("")
open(os.path.join(ROOT, "profiles/r1_sass_evidence.txt"), "w").write("\n".join(out))
print("\n".join(out))
