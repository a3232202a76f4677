"""Mnemonic counts of the shipped library (cuobjdump -sass): tcgen05 / TMEM / TMA / async-copy instructions per kernel.
usage: sass_summary.py [lib.so] > profiles/<round>_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'simmodel_b200/libs2v_b200.so'
WANT = ('UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
per, cur, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        mangled = m.group(1)
        try:
            cur = subprocess.run(['cu++filt', mangled], capture_output=True, text=True).stdout.strip() or mangled
        except FileNotFoundError:
            cur = mangled
        cur = cur.replace('void ', '').replace('(anonymous namespace)', '<unnamed>')
        cur = re.sub(r'\((bool|int|unsigned int|long)\)', '', cur)
        cur = re.sub(r'^.*?(k_\w+(<[^(]*>)?).*$', r'\1', cur)
        while cur in per:
            cur += "'"                                   # further instantiations with the same printed name
        per[cur] = collections.Counter()
        continue
    m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur is not None:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + '.'):
                per[cur][w] += 1; total[w] += 1
print('# SASS summary of %s (cuobjdump -sass, sm_100a): mnemonic counts\n' % lib)
print('whole library: ' + ', '.join('%s %d' % (k, total[k]) for k in sorted(total)))
print('\nper kernel (kernels with tcgen05 / bulk-copy / async-copy instructions):')
for k, c in sorted(per.items(), key=lambda kv: -kv[1]['UTCHMMA']):
    if any(c[w] for w in WANT if w != 'SYNCS'):
        print('  %-52s %s' % (k[:52], ', '.join('%s %d' % (w, c[w]) for w in sorted(c))))
