"""Aggregate an ncu report's SASS page by CUDA source line: joins the per-instruction rows of
`ncu --page source --print-source sass` with the line markers of `nvdisasm -g` on the object's
cubin (same build).  usage: ncu_lines.py report.ncu-rep object.o kernel_substring [topn]"""
import collections, csv, re, subprocess, sys, tempfile, os
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
kregex = sys.argv[5] if len(sys.argv) > 5 else None
lskip = sys.argv[6] if len(sys.argv) > 6 else '0'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur_fn, cur_line, in_fn = None, None, False
for l in dis:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        in_fn = kname in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', l)
    if m and cur_line: line_of[int(m.group(1), 16)] = cur_line
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'] + (['-k', 'regex:' + kregex, '--launch-skip', lskip, '--launch-count', '1'] if kregex else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, iex, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot_ex = tot_s = 0
for r in rows[2:]:
    if len(r) != len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    ln = line_of.get(a - base, ('?', 0))
    ex, sm = float(r[iex]), float(r[isamp])
    agg[ln][0] += ex; agg[ln][1] += sm
    for i in stall_cols:
        try: agg[ln][2][hdr[i]] += float(r[i])
        except ValueError: pass
    tot_ex += ex; tot_s += sm
src = {}
print(f'total inst {tot_ex:.0f} samples {tot_s:.0f}')
for ln, (ex, sm, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    f, n = ln
    if f not in src and f != '?':
        for root in ('classmate_rag_b200/csrc', '.'):
            pth = os.path.join(root, f)
            if os.path.exists(pth): src[f] = open(pth).read().splitlines(); break
    text = src.get(f, [''] * (n + 1))[n - 1].strip()[:70] if n else ''
    top = ','.join(f'{k[6:]}={v:.0f}' for k, v in st.most_common(2))
    print(f'{sm / tot_s * 100:5.1f}% samp {ex / tot_ex * 100:5.1f}% inst  {f}:{n:<4d} {top:34s} {text}')
