"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
warp-instructions executed, stall samples, avg active threads.  usage: ncu_lines.py dump.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None:
        continue
    if r[0] != "":
        d = dict(zip(hdr[4:], r[4:]))
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]), int(d["stall_no_inst"]),
                        int(d["stall_long_sb"]), int(d["stall_short_sb"]), int(d["stall_wait"]), float(d["Avg. Threads Executed"] or 0)))
        except (KeyError, ValueError):
            pass
ti = sum(a[3] for a in agg)
ts = sum(a[4] for a in agg)
print("total warp-instr %d  samples %d" % (ti, ts))
print("%-22s %5s %7s %7s %6s %6s %6s %6s %5s  %s" % ("file", "line", "inst%", "smpl%", "noins", "longsb", "shrtsb", "wait", "thr", "source"))
for a in sorted(agg, key=lambda x: -x[4])[:top]:
    print("%-22s %5d %6.2f%% %6.2f%% %6d %6d %6d %6d %5.1f  %s" % (a[0], a[1], 100 * a[3] / ti, 100 * a[4] / ts, a[5], a[6], a[7], a[8], a[9], a[2]))
