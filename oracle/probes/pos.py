# Throwaway probe: validate depth-first traversal + parallelogram/wrap restatement on POSITION of frame N.
import sys,bisect,math; sys.path.insert(0,'/tmp/draco_probe')
import probe as P, conn
from conn import nxt,prv,INV
fn=sys.argv[1]
r=conn.decode(fn); opp=r['opp']; c2v=r['c2v']; F=r['nf']; V=r['V']
R,log,B,decs=P.probe(fn)
# left-most corner / boundary test on base table
lmc=[INV]*V
# recompute lmc as conn.decode doesn't return it: any corner of v then swing left to the end
for c in range(3*F):
    v=c2v[c]
    if lmc[v]==INV: lmc[v]=c
def swl(c):
    o=opp[nxt(c)]; return nxt(o) if o>=0 else INV
def on_boundary(v):
    c=lmc[v]; first=c
    while True:
        n=swl(c)
        if n==INV: return True
        if n==first: return False
        c=n
fvis=[False]*F; vvis=[False]*V; d2c=[]; v2d=[-1]*V
def visit(v,c):
    vvis[v]=True; v2d[v]=len(d2c); d2c.append(c)
def fv(c): return True if c<0 else fvis[c//3]
for f in range(F):
    c=3*f
    if fvis[f]: continue
    st=[c]
    nv_=c2v[nxt(c)]; pv_=c2v[prv(c)]
    if not vvis[nv_]: visit(nv_,nxt(c))
    if not vvis[pv_]: visit(pv_,prv(c))
    while st:
        c=st[-1]
        if c<0 or fvis[c//3]: st.pop(); continue
        while True:
            fvis[c//3]=True
            v=c2v[c]
            if not vvis[v]:
                ob=on_boundary(v); visit(v,c)
                if not ob:
                    c=opp[nxt(c)]; continue
            rc=opp[nxt(c)]; lc=opp[prv(c)]
            if fv(rc):
                if fv(lc): st.pop(); break
                c=lc
            else:
                if fv(lc): c=rc
                else: st[-1]=lc; st.append(rc); break
assert len(d2c)==r['nv'],(len(d2c),r['nv'])
# decode symbols
B.p=R['attr_data_start']; pm,tt,comp,scheme,mbl=B.i8(),B.i8(),B.u8(),B.u8(),B.u8(); probs=P.rans_sym_create(B)
syms,_,_=P.rans_sym_decode(B,probs,mbl,r['nv']*3)
mn,mx=B.i32(),B.i32(); md=1+mx-mn
corr=[(s>>1) if not s&1 else -(s>>1)-1 for s in syms]
out=[0]*(3*r['nv']); npar=0; nwrap=0
def orig(pred,k,p):
    global nwrap
    for c in range(3):
        pc=min(mx,max(mn,pred[c])); o=pc+corr[p*3+c]
        if o>mx: o-=md; nwrap+=1
        elif o<mn: o+=md; nwrap+=1
        out[p*3+c]=o
orig([0,0,0],0,0)
for p in range(1,r['nv']):
    ci=d2c[p]; oci=opp[ci]; ok=False
    if oci>=0:
        a,b_,c_=v2d[c2v[oci]],v2d[c2v[nxt(oci)]],v2d[c2v[prv(oci)]]
        if a<p and b_<p and c_<p:
            ok=True; npar+=1
            orig([out[b_*3+k]+out[c_*3+k]-out[a*3+k] for k in range(3)],0,p)
    if not ok: orig(out[(p-1)*3:(p-1)*3+3],0,p)
print('values range',min(out),max(out),'(wrap bounds',mn,mx,') parallelogram-predicted',npar,'of',r['nv'],'wrap events',nwrap)
# edge length stats in quantized units
import statistics
el=[]
for f in range(0,F,7):
    a,b_=v2d[c2v[3*f]],v2d[c2v[3*f+1]]
    el.append(math.dist(out[a*3:a*3+3],out[b_*3:b_*3+3]))
print('edge length (quantized units) median %.2f p99 %.2f max %.2f'%(statistics.median(el),sorted(el)[int(len(el)*.99)],max(el)))
print('abs correction mean %.2f'%(sum(abs(x) for x in corr)/len(corr)))
