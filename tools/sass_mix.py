"""Instruction mix of one kernel from `ncu -i rep --page source --csv --print-source sass` output."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; si=hdr.index('Source'); ei=hdr.index('Instructions Executed'); sa=hdr.index('# Samples')
ops=collections.Counter(); samp=collections.Counter(); tot=0
for r in rows[hi+1:]:
    if len(r)<=ei or not r[ei].isdigit(): continue
    t=r[si].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    n=int(r[ei]); ops[op]+=n; tot+=n; samp[op]+=int(r[sa])
print('total warp-inst',tot)
for op,n in ops.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25): print('%-10s %9d %5.1f%%  samples %d'%(op,n,100*n/tot,samp[op]))
