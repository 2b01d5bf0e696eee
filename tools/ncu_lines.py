"""Per-source-line totals (warp instructions executed, stall samples) of a kernel from an ncu report: `python tools/ncu_lines.py x.ncu-rep [top]`."""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
out, fname = [], None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) <= ie or not r[0].isdigit():
        continue
    try:
        out.append((int(r[ie]), int(r[isamp] or 0), fname, r[0], r[1][:110]))
    except ValueError:
        pass
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
print("warp instructions", tot, "samples", tots)
for n, sm, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{n:>11} {100 * n / tot:5.1f}%  samp {100 * sm / max(1, tots):5.1f}%  {f}:{ln}: {src}")
