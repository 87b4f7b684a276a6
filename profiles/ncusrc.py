"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs --import-source on).
usage: python profiles/ncusrc.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, res = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        try:
            res.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), cur, int(r[0]), r[1].strip()[:100]))
        except Exception:
            pass
ts, ti = sum(o[0] for o in res) or 1, sum(o[1] for o in res) or 1
print(f"total: {ti} warp instructions, {ts} stall samples")
for o in sorted(res, reverse=True)[:top]:
    print(f"smp={o[0] / ts * 100:5.1f}% ins={o[1] / ti * 100:5.1f}% {o[2]}:{o[3]}  {o[4]}")
