"""Aggregate an `ncu --page source --csv` dump per SASS opcode (stall samples, executed instructions) and list the hottest
instructions.   ncu -i rep.ncu-rep --page source --csv --kernel-id :::2 > k.csv; python tools/sass_agg.py k.csv [top_n]"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[r for r in rows if 'Source' in r and '# Samples' in r][0]
iS=hdr.index('Source'); iN=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
data=[r for r in rows if len(r)==len(hdr) and r[iN].isdigit()]
tot=sum(int(r[iN]) for r in data); totE=sum(int(r[iE]) for r in data)
print('rows',len(data),'total samples',tot,'inst',totE)
def opof(s):
    t=s.split()
    if t[0].startswith('@'): t=t[1:]
    return t[0].split('.')[0]
byop=collections.Counter(); byopE=collections.Counter()
for r in data:
    op=opof(r[iS]); byop[op]+=int(r[iN]); byopE[op]+=int(r[iE])
for op,c in byop.most_common(24): print('%-10s samples %6d %5.1f%%   executed %9d %5.1f%%'%(op, c, 100*c/tot, byopE[op], 100*byopE[op]/totE))
print()
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
top=sorted(range(len(data)), key=lambda i:-int(data[i][iN]))[:N]
for i in sorted(top):
    print(i, data[i][iN], data[i][iE], data[i][iS][:110])
