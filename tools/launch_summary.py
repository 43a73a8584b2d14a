"""Per-kernel totals of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py launches.csv"""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; seq=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',',''))
            if d.get('Metric Unit')=='us': v*=1000
            seq.append((d['Kernel Name'],d.get('Grid Size',''),v))
# take the last step: find last occurrence of concat kernel
idx=[i for i,(k,g,v) in enumerate(seq) if 'concat_ndhwc' in k or 'pack_nhwc' in k]
print('launches',len(seq),'concat at',idx[-6:])
# the last COMPLETE step: from the second-to-last concat launch to the last one (the capture may end mid-step)
step=seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq[idx[-1]:]
agg=collections.OrderedDict(); tot=0
for k,g,v in step:
    key=k.replace('<unnamed>::','')[:80]+' '+g
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
for k,(n,t) in agg.items(): print('%2d x %8.1f us = %8.1f us  %s'%(n,t/n/1000,t/1000,k))
print('total us',tot/1000)
