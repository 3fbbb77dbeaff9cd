"""Summarise an ncu report by CUDA source line: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; items = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    elif r and r[0] == 'Line No': hdr = r
    elif hdr and len(r) >= 10 and r[0].isdigit() and r[2] == '-':
        d = dict(zip(hdr, r))
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '-', '0')}
        items.append((int(r[6] or 0), int(r[7] or 0), cur, int(r[0]), r[1].strip()[:90], st))
tot = sum(i[0] for i in items)
print('total samples', tot)
for s, ie, f, l, src, st in sorted(items, key=lambda x: -x[0])[:top]:
    top3 = ','.join('%s:%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:6d} {100*s/max(tot,1):5.1f}% inst={ie:9d} {f}:{l:<4d} {src:90s} {top3}")
