"""Per-layer kernel times of the last step in an ncu launch list CSV (gpu__time_duration), next to the HBM ideal."""
import csv, sys
sys.path.insert(0, '.')
def load(f):
    rows=list(csv.reader(open(f)))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    hdr=rows[hi]; body=rows[hi+1:]
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
    names=[r[ki] for r in body]
    s=[i for i,n in enumerate(names) if 'logmel' in n][-1]
    return [(r[ki].split('(')[0].replace('ac::','').replace('void ',''), r[gi], float(r[vi].replace(',',''))/1000) for r in body[s:]]
b=load(sys.argv[1])
from audiocaption_b200 import roofline as rl
stem, blocks, last = rl.effb2_geometry(64,1001)
dws=[x for x in b if 'dwconv' in x[0]]; ses=[x for x in b if 'se_kernel' in x[0]]; gm=[x for x in b if 'gemm' in x[0]]
gi=0
tot=dict(dw=0,se=0,ex=0,pr=0)
for i,((cin,cout,e,k,s,lo,hi,nsq,skip),(H,W),(Ho,Wo)) in enumerate(blocks):
    ce=cin*e; by=64*(H*W+Ho*Wo)*ce*4
    ex=None
    if e!=1: ex=gm[gi]; gi+=1
    pr=gm[gi]; gi+=1
    exb=64*H*W*(cin+ce)*4; prb=64*Ho*Wo*(ce+cout*(2 if skip else 1))*4
    print(f"b{i:2d} k{k}s{s} C={ce:4d} {H:2d}x{W:3d}->{Ho:2d}x{Wo:3d} | expand {ex[2] if ex else 0:6.1f} ({exb/6.55e6 if ex else 0:5.1f}) | dw {dws[i][2]:6.1f} ({by/6.55e6:5.1f}) {by/dws[i][2]/1e3:5.0f} GB/s | se {ses[i][2]:5.1f} | project {pr[2]:6.1f} ({prb/6.55e6:5.1f})")
    tot['dw']+=dws[i][2]; tot['se']+=ses[i][2]; tot['pr']+=pr[2]; tot['ex']+=ex[2] if ex else 0
print(tot, 'others:', [(x[0][:20], round(x[2],1)) for x in b if not any(t in x[0] for t in ('dwconv','se_kernel','gemm_tc'))], 'tail gemm', [round(x[2],1) for x in gm[gi:]])
