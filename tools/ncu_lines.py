"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
text = {}
cur_line = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r
        ii, si, ti = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
        continue
    if hdr and len(r) == len(hdr):
        if r[0].isdigit():
            cur_line = (cur_file, int(r[0]))
            text[cur_line] = r[1]
            continue                      # the CUDA line row carries the aggregate of its SASS rows: skip, sum SASS rows instead
        if r[2]:                          # SASS row (has an address)
            try:
                agg[cur_line][0] += int(r[ii] or 0); agg[cur_line][1] += int(r[si] or 0); agg[cur_line][2] += int(r[ti] or 0)
            except ValueError:
                pass
ti_ = sum(v[0] for v in agg.values()); ts_ = sum(v[1] for v in agg.values())
print('total warp-inst', ti_, 'samples', ts_)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f'{k[0]}:{k[1]:>4} inst {100 * v[0] / ti_:5.1f}% samp {100 * v[1] / ts_:5.1f}% act {v[2] / max(v[0], 1):4.1f} | {text.get(k, "")[:100].strip()}')
