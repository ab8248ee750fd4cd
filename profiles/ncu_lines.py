"""Per-source-line instruction and stall-sample totals of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on).  Usage: python profiles/ncu_lines.py REPORT [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); src = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    d = dict(zip(hdr, r))
    try: ln = int(r[0])
    except ValueError: continue
    key = (fname, ln)
    inst[key] += int(d['Instructions Executed'] or 0)
    samp[key] += int(d['# Samples'] or 0)
    src[key] = r[1].strip()[:110]
ti, ts = sum(inst.values()), sum(samp.values())
print(f'total warp instructions {ti}, stall samples {ts}')
byfile = collections.Counter()
for (f, l), v in inst.items(): byfile[f] += v
print('by file:', {f: f'{100*v/ti:.1f}%' for f, v in byfile.items()})
print(f'{"file:line":28s} {"inst%":>6s} {"samp%":>6s}  source')
for key, v in inst.most_common(top):
    print(f'{key[0]+":"+str(key[1]):28s} {100*v/ti:6.2f} {100*samp[key]/max(ts,1):6.2f}  {src[key]}')
