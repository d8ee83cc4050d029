import re, sys, numpy as np
txt=open(sys.argv[1]).read().split("==== timed batch ====")[1]
L=[]
for l in txt.splitlines():
    m=re.search(r"jobs (\d+) blocks (\d+) kernel ([\d.]+) ms, expansions max (\d+) sum (\d+)",l)
    if m: L.append(tuple(float(x) for x in m.groups()))
hosts=[l for l in txt.splitlines() if 'mrp_host' in l]
for h in hosts[:2]: print(h[:330])
A=np.array(L); print(len(A),'launches; sum kernel ms %.0f; sum maxExp*3.2us %.0f ms'%(A[:,2].sum(), A[:,3].sum()*3.2e-3))
us=A[:,2]*1e3/np.maximum(A[:,3],1); print('us/exp longest: median %.2f p90 %.2f'%(np.median(us),np.percentile(us,90)))
