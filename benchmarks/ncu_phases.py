#!/usr/bin/env python
"""Attribute an `ncu --page source --csv` export to the '// ---- N:' phase markers of a .cu file.
usage: ncu_phases.py <nvdisasm -g -c output> <ncu source csv> <mangled kernel name> <source .cu>"""
import collections, csv, re, sys
sass_path, csv_path, kname, cu = sys.argv[1:5]
src = open(cu).read().splitlines()
marks = [(i + 1, l.strip()[:60]) for i, l in enumerate(src) if re.match(r'\s*// ---- \w+', l)]
lines = open(sass_path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':'))
off2line, cur = {}, None
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*)', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ia, isamp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ia].startswith('0x'):
        if data:
            break
        continue
    data.append(r)
base = int(data[0][ia], 16)
cuname = cu.split('/')[-1]
# instructions inlined from headers inherit the phase of the last instruction that mapped to the .cu file
agg = collections.OrderedDict()
phase = 'prologue'
tot_s = tot_i = 0
for r in data:
    fl = off2line.get(int(r[ia], 16) - base)
    if fl and fl[0] == cuname:
        ph = 'prologue'
        for ln, name in marks:
            if fl[1] >= ln:
                ph = name
        phase = ph
    s, n = int(r[isamp] or 0), int(r[iinst] or 0)
    a = agg.setdefault(phase, [0, 0])
    a[0] += s
    a[1] += n
    tot_s += s
    tot_i += n
print('total samples %d, warp instructions %.1fM' % (tot_s, tot_i / 1e6))
for ph, (s, n) in agg.items():
    print('%-62s samples %5.1f%%  inst %5.1f%% (%6.2fM)' % (ph, 100 * s / tot_s, 100 * n / tot_i, n / 1e6))
