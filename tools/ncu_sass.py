"""Dynamic SASS view of a kernel from an ncu report (`--import-source on`): opcode mix by executed warp instructions and, with a
count argument, the instructions executed exactly that many times (e.g. once per bundle): `python tools/ncu_sass.py x.ncu-rep [count ...]`."""
import collections, csv, re, subprocess, sys
rep, counts = sys.argv[1], [int(v) for v in sys.argv[2:]]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
ins = []
for r in rows:
    if len(r) <= ie or r is hdr:
        continue
    try:
        ins.append((r[isrc].strip(), int(r[ie])))
    except ValueError:
        pass
def opc(s):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    return m.group(2) if m else "?"
tot = sum(n for _, n in ins)
mix = collections.Counter()
for s, n in ins:
    mix[opc(s)] += n
print("executed warp instructions:", tot)
for k, v in mix.most_common(24):
    print(f"{v:>11} {100 * v / tot:5.2f}% {k}")
levels = collections.Counter(n for _, n in ins)
print("execution-count levels (count: static instructions):", sorted(levels.items(), reverse=True)[:12])
for i, (s, n) in enumerate(ins):
    if n in counts:
        print(i, n, s[:100])
