"""Rank source lines / call-site phases of an ncu report by executed instructions and stall samples.
   python tools/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; lines = []; ins = []; line = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Function Name'): continue
    if r[0] and r[0].isdigit():
        line = int(r[0])
        try: lines.append((cur, line, r[1], int(r[6]), int(r[7])))
        except ValueError: pass
        continue
    if len(r) > 7 and r[2].startswith('0x'):
        try: ins.append((int(r[2], 16), int(r[7]), int(r[6]), cur, line, r[3].strip()))
        except ValueError: pass
ti = sum(l[4] for l in lines); ts = sum(l[3] for l in lines)
print("total inst", ti, "samples", ts)
lines.sort(key=lambda x: -x[4])
for f, l, s, samp, inst in lines[:topn]:
    print("%-16s %4d inst %5.1f%% samp %5.1f%%  %s" % (f[:16], l, 100 * inst / ti, 100 * samp / ts, s[:110]))
print("---- by call-site phase (bc_assign.cu / umi_dist.cu line that precedes in address order) ----")
ins.sort(); seg = 0; acc = {}; order = []
for a, n, s, f, l, sa in ins:
    if f in ('bc_assign.cu', 'umi_dist.cu', 'bc_collide.cu'): seg = l
    if seg not in acc: acc[seg] = [0, 0]; order.append(seg)
    acc[seg][0] += n; acc[seg][1] += s
for k in order:
    n, s = acc[k]
    if n / max(ti, 1) > 0.004: print("line %4d  inst %5.1f%%  samples %5.1f%%" % (k, 100 * n / ti, 100 * s / ts))
