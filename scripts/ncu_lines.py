"""Per-source-line share of executed warp instructions and stall samples of an .ncu-rep (needs -lineinfo):
    python scripts/ncu_lines.py <report.ncu-rep> [top N]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if r[ie].isdigit():
            data.append((int(r[ie]), int(r[isamp]), fname, int(r[0]), r[1].strip()[:110]))
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"{tot} warp instructions, {ts} samples")
for e, s, f, ln, src in sorted(data, key=lambda t: -t[0])[:top]:
    print(f"{100 * e / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% samples  {f}:{ln}  {src}")
