"""SASS evidence per kernel: opcode histogram of every kernel in libwgg_sm100.so that uses the Blackwell-native
instructions (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier).
  python scripts/sass_histogram.py > profiles/r02_sass_opcode_histogram.md     (runs cuobjdump; no GPU needed)"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "wordgesture-gan_b200", "libwgg_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "")
        kern = re.sub(r"^void ", "", re.sub(r"\(.*", "", kern))
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "HMMA", "FFMA", "LDGSTS")
print("# SASS opcode histogram of libwgg_sm100.so (sm_100a), `cuobjdump -sass`\n")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM load), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (bulk async copy),")
print("UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA with a tensor map),")
print("SYNCS = mbarrier ops, HMMA = mma.sync (the 3xTF32 / generic-shape engine), LDGSTS = cp.async.  Kernels without any")
print("tensor-core or async-copy instruction are listed at the end with their instruction count only.\n")
print("| kernel | instr | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
plain = []
for k, c in hist.items():
    if any(c[o] for o in ("UTCHMMA", "LDTM", "UBLKCP", "HMMA", "UTMALDG")):
        print(f"| `{k[:90]}` | {sum(c.values())} | " + " | ".join(str(c[o]) for o in KEY) + " |")
    else:
        plain.append((k, sum(c.values()), c["MUFU"], c["FFMA"]))
print("\nOther kernels (no tensor-core / bulk-copy instructions): " + ", ".join(f"`{k[:60]}` ({n})" for k, n, _, _ in plain))
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
print(f"\nTotals over the library: " + ", ".join(f"{o} {tot[o]}" for o in KEY))
