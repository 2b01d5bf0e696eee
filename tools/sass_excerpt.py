"""SASS evidence for the Blackwell-native claims: per kernel of the shipped library, the counts of the mnemonics that
matter (UTMALDG = TMA tensor loads, SYNCS = mbarrier operations, VABSDIFF4 = packed-byte SAD, REDUX / MATCH, DADD / DMUL of the
FP64 DCT replay) plus the first lines of the search kernel's inner loop.  usage: python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "streamoptima_b200", "libstreamoptima_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = re.search(r"arch = (\S+)", out)
print("library:", os.path.relpath(lib, ROOT), " arch =", arch.group(1) if arch else "?")
funcs = re.split(r"\n\s*Function : ", out)[1:]
WANT = ["UTMALDG", "SYNCS", "VABSDIFF4", "REDUX", "CREDUX", "MATCH", "DADD", "DMUL", "LDS", "ATOMG", "REDG", "UTCHMMA", "HMMA"]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    ops = collections.Counter()
    n = 0
    for line in f.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
            n += 1
    rows.append((dem, n, ops))
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{w:>9s}" for w in WANT))
for dem, n, ops in sorted(rows, key=lambda r: -r[2]["VABSDIFF4"]):
    if n == 0:
        continue
    print(f"{dem[:58]:58s} {n:6d} " + " ".join(f"{ops[w]:9d}" for w in WANT))
print("\n-- me_ring2_kernel<false>: the TMA issue and the start of the SAD loop")
for f in funcs:
    if "me_ring2_kernelILb0" in f.split("\n", 1)[0]:
        lines = [l for l in f.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        for i, l in enumerate(lines):
            if "UTMALDG" in l:
                print("\n".join(x.split("/*", 2)[0] + x.split("*/", 1)[1].rsplit("/*", 1)[0] for x in lines[max(0, i - 3):i + 2]))
                print("   ...")
        first = next(i for i, l in enumerate(lines) if "VABSDIFF4" in l)
        print("\n".join(x.split("*/", 1)[1].rsplit("/*", 1)[0] for x in lines[first - 6:first + 30]))
        break
