import csv, re, sys, collections
sass_path, csv_path, kname = sys.argv[1], sys.argv[2], sys.argv[3]
# 1) map instruction offset -> (file,line) from nvdisasm -g output for the kernel section
lines = open(sass_path).read().splitlines()
start = next(i for i,l in enumerate(lines) if l.startswith('.text.'+kname+':'))
off2line = {}
cur = None
for l in lines[start+1:]:
    if l.startswith('//--------------------- .text') : break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*)', l)
    if m: off2line[int(m.group(1),16)] = (cur, m.group(2))
# 2) read ncu source csv (first kernel instance)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ia, isamp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ia].startswith('0x'):
        if data: break
        continue
    data.append(r)
base = int(data[0][ia],16)
agg = collections.defaultdict(lambda: [0,0,collections.Counter()])
tot_s = tot_i = 0
for r in data:
    off = int(r[ia],16)-base
    key = off2line.get(off, (('?',0),''))[0] or ('?',0)
    s, n = int(r[isamp] or 0), int(r[iinst] or 0)
    agg[key][0]+=s; agg[key][1]+=n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: agg[key][2][hdr[c]] += v
    tot_s+=s; tot_i+=n
print('total samples', tot_s, 'total inst', tot_i)
src = {}
for (f,l),(s,n,st) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    top = ', '.join('%s:%d'%(k.replace('stall_',''),v) for k,v in st.most_common(3))
    print('%-18s:%4d  samples %5.1f%%  inst %5.1f%%   %s' % (f,l,100*s/tot_s,100*n/tot_i, top))
