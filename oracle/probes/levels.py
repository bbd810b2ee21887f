import sys; sys.path.insert(0,'/tmp/draco_probe')
import io,contextlib,runpy,collections
sys.argv=['uvn.py',sys.argv[1]]
with contextlib.redirect_stdout(io.StringIO()): g=runpy.run_path('/tmp/draco_probe/uvn.py')
opp,c2v,pd2c,pv2d,nv=g['opp'],g['c2v'],g['pd2c'],g['pv2d'],g['r']['nv']
from conn import nxt,prv
lvl=[0]*nv
for p in range(1,nv):
    ci=pd2c[p]; o=opp[ci]; deps=None
    if o>=0:
        a,b,c=pv2d[c2v[o]],pv2d[c2v[nxt(o)]],pv2d[c2v[prv(o)]]
        if a<p and b<p and c<p: deps=(a,b,c)
    if deps is None: deps=(p-1,)
    lvl[p]=1+max(lvl[d] for d in deps)
h=collections.Counter(lvl); L=max(lvl)+1; w=sorted(h.values())
print('POSITION parallelogram DAG: entries',nv,'levels',L,'mean width %.1f'%(nv/L),'median width',w[len(w)//2],'max width',w[-1],'levels with width>=32: %d (%.0f%% of entries)'%(sum(1 for x in w if x>=32),100*sum(x for x in w if x>=32)/nv))
