#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS stall samples by CUDA source line (needs -lineinfo and the matching .so).
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-substring> <source.cu> [top]
The substring must select ONE kernel of the library (e.g. conv_igemm_kernelILb0 for conv_igemm_kernel<false>, not conv_igemm_kernel, which
also matches the <true> and persistent variants), and the .so must be the build the report was captured with."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kname, srcfile = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'pasta-gan_b200', 'lib', 'libpasta_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith('.cubin'):
        out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in out:
            dis = out
seq, in_k, cur = [], False, None
base = os.path.basename(srcfile)
for ln in dis.split('\n'):
    if ln.startswith('.text.'):
        in_k = kname in ln
    if not in_k:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(base) else None
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*;', ln):
        seq.append(cur)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) > ci['# Samples']]
if len(body) != len(seq):
    print(f'warning: {len(body)} ncu instrs vs {len(seq)} disassembled (stale .so?)')
src = open(srcfile).read().split('\n')
agg = {}
for line, r in zip(seq, body):
    s = int(r[ci['# Samples']]) if r[ci['# Samples']].isdigit() else 0
    e = int(r[ci['Instructions Executed']]) if r[ci['Instructions Executed']].isdigit() else 0
    a = agg.setdefault(line, [0, 0]); a[0] += s; a[1] += e
tot = sum(a[0] for a in agg.values()) or 1; tote = sum(a[1] for a in agg.values()) or 1
print(f'total samples {tot}, warp-instructions {tote}')
for line, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[line - 1].strip()[:105] if line else '(inlined / other file)'
    print(f'{s:6d} {100 * s / tot:5.1f}%  exec {100 * e / tote:5.1f}%  L{line}: {text}')
