"""SASS opcode histogram per kernel of libkgeb200.so -> profiles/r02_sass_histogram.md (run here: cuobjdump needs no GPU)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "kge-keras_b200", "lib", "libkgeb200.so")
cols = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "RED", "ATOMG", "LDGSTS", "CCTL", "LDG", "STG", "LDS", "STS",
        "FFMA", "FADD", "FADD2", "FFMA2", "FMUL", "HMMA", "MUFU", "SHFL"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
hist, variants, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        k = re.search(r"(k_[a-z0-9_]+)", d)
        cur = k.group(1) if k else d[:40]
        variants[cur] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        hist[cur][op] += 1
out = ["# SASS instruction histogram of libkgeb200.so (sm_100a), round 2 (final build)", "",
       "`python profiles/sass_histogram.py` = `cuobjdump -sass kge-keras_b200/lib/libkgeb200.so`, opcodes counted per kernel (template instantiations of one kernel summed; `variants` = how many).",
       "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc, UBLKCP = cp.async.bulk (1-D bulk copy), UTMALDG = cp.async.bulk.tensor (TMA through a tensor map), "
       "SYNCS = mbarrier ops, REDG = red.global (vector reductions), CCTL = prefetch.global.L2, FADD2 / FFMA2 = packed fp32 (add.f32x2 / fma.f32x2).", "",
       "| kernel | variants | instr | " + " | ".join(cols) + " |", "|---|---|---|" + "---|" * len(cols)]
for k, n in sorted(total.items(), key=lambda x: -x[1]):
    if not k.startswith("k_"):
        continue
    out.append(f"| `{k}` | {variants[k]} | {n} | " + " | ".join(str(hist[k][c]) if hist[k][c] else "" for c in cols) + " |")
open(os.path.join(ROOT, "profiles", "r02_sass_histogram.md"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if "k_rank_t" in l))
