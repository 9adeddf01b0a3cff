#!/usr/bin/env python
"""SASS evidence of the Blackwell-native instructions in the shipped library (no GPU needed): per kernel, counts of the mnemonics that
B200_PROFILING.md names -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (tensor TMA), UBLKCP (bulk copy),
HMMA (legacy mma.sync: must be absent) -- plus registers per thread.  usage: sass_summary.py [lib.so] > profiles/r2_sass_summary.txt"""
import collections, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'pasta-gan_b200', 'lib', 'libpasta_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
regs = dict(re.findall(r'Function (\S+):\s*\n\s*REG:(\d+)', res))
pat = re.compile(r'\b(UTC[A-Z]*MMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTMAPF|SYNCS|HMMA|HGMMA|QGMMA|IGMMA|LDGSTS|ELECT|UTCATOMSWS|REDG|ATOMG|FFMA2|FMUL2)\b')
per = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', ln)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    if cur:
        for mm in pat.findall(ln):
            per[cur][mm] += 1
digest = hashlib.sha256(open(lib, 'rb').read()).hexdigest()[:16]
stamp = os.path.join(os.path.dirname(lib), 'libpasta_b200.stamp')
print(f'# {os.path.relpath(lib, ROOT)}  sha256[:16] = {digest}  source digest = {open(stamp).read().strip()[:16] if os.path.exists(stamp) else "?"}')
arch = sorted(set(re.findall(r'arch = (sm_\w+)', sass)))
print(f'# cubin architectures: {arch}')
tot = collections.Counter()
for fn, c in per.items():
    tot.update(c)
demangle = lambda s: subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip() or s
print('\n# totals over the library')
print('  ' + '  '.join(f'{k}={v}' for k, v in sorted(tot.items())))
print('\n# per kernel (only kernels with tensor-core / TMA / bulk-copy instructions)')
for fn, c in per.items():
    if any(k.startswith('UTC') or k in ('UTMALDG', 'UBLKCP', 'LDTM') for k in c):
        name = demangle(fn)
        print(f'{name[:150]}\n    regs={regs.get(fn, "?")}  ' + '  '.join(f'{k}={v}' for k, v in sorted(c.items())))
print('\n# legacy tensor path (must be zero): HMMA=%d HGMMA=%d' % (tot.get('HMMA', 0), tot.get('HGMMA', 0)))
