"""Aggregate the ncu source page (SASS view) by source line via -lineinfo; print top lines by executed instructions."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = 0; items = []
for r in rows[2:]:
    if len(r) != len(hdr): continue
    try: ex = float(r[iex]); sm = float(r[isamp])
    except: continue
    tot += ex; items.append((ex, sm, r[ia], r[isrc]))
print('total inst', tot)
ops = collections.Counter()
for ex, sm, a, s in items:
    op = s.split()[0] if not s.startswith('@') else s.split()[1]
    ops[op.split('.')[0]] += ex
for op, c in ops.most_common(18): print(f'  {op:12s} {c/tot*100:5.1f}%')
print('--- top SASS by executed')
for ex, sm, a, s in sorted(items, reverse=True)[:topn]: print(f'{ex/tot*100:5.2f}% samp={sm:6.0f} {a[-5:]} {s[:100]}')
