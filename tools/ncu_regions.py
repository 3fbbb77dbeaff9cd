"""Stall samples of an ncu report grouped by source-line ranges: python tools/ncu_regions.py rep.ncu-rep file:lo-hi=name ..."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
regs = []
for a in sys.argv[2:]:
    loc, name = a.split('=')
    f, rng = loc.split(':')
    lo, hi = rng.split('-')
    regs.append((f, int(lo), int(hi), name))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None
tot = collections.Counter(); st_by = collections.defaultdict(collections.Counter); inst = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    elif r and r[0] == 'Line No': hdr = r
    elif hdr and len(r) >= 10 and r[0].isdigit() and r[2] == '-':
        d = dict(zip(hdr, r))
        line = int(r[0]); s = int(r[6] or 0)
        name = 'other'
        for f, lo, hi, n in regs:
            if cur == f and lo <= line <= hi:
                name = n; break
        tot[name] += s; inst[name] += int(r[7] or 0)
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '-', '0'):
                st_by[name][k[6:]] += int(v)
T = sum(tot.values())
for n, s in tot.most_common():
    top = ', '.join('%s %.1f%%' % (k, 100.0 * v / max(s, 1)) for k, v in st_by[n].most_common(5))
    print('%-12s %7d %5.1f%%  inst=%11d  %s' % (n, s, 100.0 * s / T, inst[n], top))
