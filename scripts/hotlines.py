#!/usr/bin/env python
"""Attribute executed SASS instructions of one kernel in an ncu report to CUDA source lines.
usage: hotlines.py <report.ncu-rep> <object.o> <kernel-substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, obj, kname = sys.argv[1:4]
mangled = sys.argv[5] if len(sys.argv) > 5 else kname
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith('.text.') or '.section' in ln and '.text.' in ln:
        infn = mangled in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S+)', ln)
    if m and infn and cur:
        addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name' and kname in r[1]]
hdr = rows[start[0] + 1]
ia, ie, isrc = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Source')
ist = hdr.index('Warp Stall Sampling (All Samples)')
base = None
agg, aggs, tot, tots = collections.Counter(), collections.Counter(), 0, 0
ops = collections.Counter()
for r in rows[start[0] + 2:]:
    if not r or r[0] == 'Kernel Name':
        break
    try:
        a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    n = float(r[ie] or 0); s = float(r[ist] or 0)
    key = addr2line.get(a - base, ('?', 0, ()))
    outer = key[2][-1] if key[2] else (key[0], key[1])
    agg[(key[0], key[1], outer)] += n; aggs[(key[0], key[1], outer)] += s
    tot += n; tots += s
    ops[r[isrc].split()[0] if r[isrc] else '?'] += n
print('total instr %.3g, stall samples %.3g' % (tot, tots))
for (f, l, outer), n in agg.most_common(top):
    print('%5.1f%% instr %5.1f%% stall  %s:%d  (in %s:%d)' % (100 * n / tot, 100 * aggs[(f, l, outer)] / max(tots, 1), f, l, outer[0], outer[1]))
print('opcodes:', ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in ops.most_common(14)))
