"""Rank CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` dump by warp-stall samples.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/hot_lines.py src.csv [file-filter] [top]"""
import csv, sys, collections
path = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ''; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), ''])
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Name': cur = r[1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    # columns: first 'Source' = cuda text, second = sass; dict keeps the last -> use indices
    line = int(r[0]); text = r[1]
    samples = int(r[hdr.index('# Samples')] or 0); inst = int(r[hdr.index('Instructions Executed')] or 0)
    thr = int(r[hdr.index('Thread Instructions Executed')] or 0)
    key = (cur, line)
    a = agg[key]; a[0] += samples; a[1] += inst; a[2] += thr; a[4] = text
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h and r[i] not in ('', '0'):
            a[3][h] += int(r[i])
    tot_s += samples; tot_i += inst
print(f'total samples {tot_s} warp-instructions {tot_i}')
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if filt and filt not in (f or ''): continue
    st = ', '.join(f'{k[6:]} {v}' for k, v in a[3].most_common(2))
    print(f'{(f or "").split("/")[-1]}:{line:<5d} samples {100*a[0]/max(tot_s,1):5.1f}%  inst {100*a[1]/max(tot_i,1):5.1f}% ({a[1]/1e6:7.1f}M, {a[2]/max(a[1],1):4.1f} thr)  [{st}]  {a[4].strip()[:90]}')
