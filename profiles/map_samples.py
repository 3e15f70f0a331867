#!/usr/bin/env python
"""Map ncu --page source (SASS) stall samples to CUDA source lines using nvdisasm -g output.
usage: map_samples.py prof_src.csv kernel.sass kernel_name source.cu [topN]"""
import csv, re, sys
src_csv, sass, kname, cu = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 30
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
cur = None; off2line = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .text'):
        break
    m = re.search(r'//## File "[^"]*", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]; si = h.index('# Samples')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
base = int(data[0][0], 16)
agg = {}; st = {}
for r in data:
    ln = off2line.get(int(r[0], 16) - base)
    agg[ln] = agg.get(ln, 0) + int(r[si])
    d = st.setdefault(ln, {})
    for i in stall_cols:
        d[h[i]] = d.get(h[i], 0) + int(r[i] or 0)
src = open(cu).read().split('\n')
tot = sum(agg.values())
print('total samples', tot)
for ln, c in sorted(agg.items(), key=lambda x: -x[1])[:topn]:
    top = sorted(st[ln].items(), key=lambda x: -x[1])[:2]
    print(f"{ln} {c} {100*c/tot:.1f}% {src[ln-1].strip()[:90] if (ln and ln <= len(src)) else None}  {top}")
