# Throwaway probe: validate TEX_COORDS_PORTABLE and GEOMETRIC_NORMAL restatements on a fixture frame.
import sys,math; sys.path.insert(0,'/tmp/draco_probe')
import probe as P, conn
from conn import nxt,prv,INV
fn=sys.argv[1]
r=conn.decode(fn); opp=r['opp']; c2v=r['c2v']; F=r['nf']; V=r['V']; lmc=r['lmc']
R,log,B,decs=P.probe(fn)
class Table:   # corner table view (base or attribute)
    def __init__(s,c2v,lmc,eos=None,vos=None): s.c2v=c2v; s.lmc=lmc; s.eos=eos; s.vos=vos
    def opp(s,c):
        if c<0: return INV
        if s.eos is not None and s.eos[c]: return INV
        return opp[c]
    def swl(s,c):
        o=s.opp(nxt(c)); return nxt(o) if o>=0 else INV
    def swr(s,c):
        o=s.opp(prv(c)); return prv(o) if o>=0 else INV
    def on_boundary(s,v):
        c=s.lmc[v]
        if c<0: return True
        if s.vos is not None: return s.vos[c2v[c]]   # MeshAttributeCornerTable::IsOnBoundary == IsCornerOnSeam(leftmost corner)
        return s.swl(c)==INV
def traverse(T,nverts):
    fvis=[False]*F; vvis=[False]*nverts; d2c=[]; v2d=[-1]*nverts
    def visit(v,c): vvis[v]=True; v2d[v]=len(d2c); d2c.append(c)
    def fv(c): return True if c<0 else fvis[c//3]
    for f in range(F):
        c=3*f
        if fvis[f]: continue
        st=[c]
        for cc in (nxt(c),prv(c)):
            if not vvis[T.c2v[cc]]: visit(T.c2v[cc],cc)
        while st:
            c=st[-1]
            if c<0 or fvis[c//3]: st.pop(); continue
            while True:
                fvis[c//3]=True; v=T.c2v[c]
                if not vvis[v]:
                    ob=T.on_boundary(v); visit(v,c)
                    if not ob: c=T.opp(nxt(c)); continue
                rc=T.opp(nxt(c)); lc=T.opp(prv(c))
                if fv(rc):
                    if fv(lc): st.pop(); break
                    c=lc
                else:
                    if fv(lc): c=rc
                    else: st[-1]=lc; st.append(rc); break
    return d2c,v2d
def zz(s): return (s>>1) if not s&1 else -(s>>1)-1
def isqrt(n):
    if n==0: return 0
    a=n; r=1
    while a>=2: r*=2; a//=4
    while True:
        r=(r+n//r)//2
        if r*r<=n: return r
def tdiv(a,b):  # C++ truncating division
    q=abs(a)//abs(b); return q if (a>=0)==(b>=0) else -q
# ---------- positions ----------
base=Table(c2v,lmc)
# base on_boundary must use full swing-left loop end: leftmost corner for boundary verts is the true leftmost, so swl==INV test suffices
pd2c,pv2d=traverse(base,V)
B.p=R['attr_data_start']; pm,tt,comp,scheme,mbl=B.i8(),B.i8(),B.u8(),B.u8(),B.u8(); probs=P.rans_sym_create(B)
syms,_,_=P.rans_sym_decode(B,probs,mbl,r['nv']*3); mn,mx=B.i32(),B.i32(); md=1+mx-mn
corr=[zz(s) for s in syms]; pos=[0]*(3*r['nv'])
def wrap_add(pred,cor,mn,mx):
    md=1+mx-mn; out=[]
    for a,b in zip(pred,cor):
        o=min(mx,max(mn,a))+b
        if o>mx:o-=md
        elif o<mn:o+=md
        out.append(o)
    return out
pos[0:3]=wrap_add([0,0,0],corr[0:3],mn,mx)
for p in range(1,r['nv']):
    ci=pd2c[p]; oci=opp[ci]; pred=None
    if oci>=0:
        a,b_,c_=pv2d[c2v[oci]],pv2d[c2v[nxt(oci)]],pv2d[c2v[prv(oci)]]
        if a<p and b_<p and c_<p: pred=[pos[b_*3+k]+pos[c_*3+k]-pos[a*3+k] for k in range(3)]
    if pred is None: pred=pos[(p-1)*3:(p-1)*3+3]
    pos[p*3:p*3+3]=wrap_add(pred,corr[p*3:p*3+3],mn,mx)
pmin=[B.f32() for _ in range(3)]; prange=B.f32(); pbits=B.u8()
def P3(corner): e=pv2d[c2v[corner]]; return pos[e*3:e*3+3]
# ---------- UV ----------
A=r['atts'][0]; T=Table(A['c2v'],A['lmc'],A['eos'],A['vos']); nU=A['num_vertices']
ud2c,uv2d=traverse(T,nU); assert len(ud2c)==nU,(len(ud2c),nU)
pm,tt,comp,scheme,mbl=B.i8(),B.i8(),B.u8(),B.u8(),B.u8(); probs=P.rans_sym_create(B)
syms,_,_=P.rans_sym_decode(B,probs,mbl,nU*2); ucorr=[zz(s) for s in syms]
nor=B.i32(); p0=B.u8(); size=B.varint(); rs=conn.RabsStream(B.b,p0,B.p,size); B.p+=size
orient=[]; last=True
for i in range(nor):
    if not rs.bit(): last=not last
    orient.append(last)
umn,umx=B.i32(),B.i32()
uv=[0]*(2*nU); used=0; npred=0; ndeg=0; nfb=0; wraps=0
for p in range(nU):
    c=ud2c[p]; nc_,pc_=nxt(c),prv(c)
    nd,pd=uv2d[T.c2v[nc_]],uv2d[T.c2v[pc_]]
    pred=None
    if pd<p and nd<p:
        nuv=uv[nd*2:nd*2+2]; puv=uv[pd*2:pd*2+2]
        if nuv==puv: pred=list(puv); ndeg+=1
        else:
            tip,npos,ppos=P3(c),P3(nc_),P3(pc_)
            pn=[ppos[k]-npos[k] for k in range(3)]; pn2=sum(x*x for x in pn)
            if pn2!=0:
                cn=[tip[k]-npos[k] for k in range(3)]; dot=sum(pn[k]*cn[k] for k in range(3))
                pnuv=[puv[0]-nuv[0],puv[1]-nuv[1]]
                xuv=[nuv[0]*pn2+dot*pnuv[0], nuv[1]*pn2+dot*pnuv[1]]
                xpos=[npos[k]+tdiv(dot*pn[k],pn2) for k in range(3)]
                cx2=sum((tip[k]-xpos[k])**2 for k in range(3))
                cxuv=[pnuv[1],-pnuv[0]]; ns=isqrt(cx2*pn2); cxuv=[cxuv[0]*ns,cxuv[1]*ns]
                o=orient.pop(); used+=1
                if o: pu=[tdiv(xuv[0]+cxuv[0],pn2),tdiv(xuv[1]+cxuv[1],pn2)]
                else: pu=[tdiv(xuv[0]-cxuv[0],pn2),tdiv(xuv[1]-cxuv[1],pn2)]
                pred=pu; npred+=1
    if pred is None:
        nfb+=1
        if nd<p: pred=uv[nd*2:nd*2+2]
        elif p>0: pred=uv[(p-1)*2:(p-1)*2+2]   # upstream quirk: 'else' binds to the next-corner test only, so prev-only falls to last entry
        else: pred=[0,0]
    before=[min(umx,max(umn,x))+y for x,y in zip(pred,ucorr[p*2:p*2+2])]
    wraps+=sum(1 for x in before if x>umx or x<umn)
    uv[p*2:p*2+2]=wrap_add(pred,ucorr[p*2:p*2+2],umn,umx)
print('UV: entries',nU,'orientations in stream',nor,'consumed',used,'left',len(orient),'| predicted',npred,'degenerate-copy',ndeg,'fallback',nfb,'| range',min(uv),max(uv),'wrap events',wraps,'| mean|corr| %.2f'%(sum(abs(x) for x in ucorr)/len(ucorr)))

# ---------- UV dequant params ----------
umin=[B.f32(),B.f32()]; urange=B.f32(); ubits=B.u8()
# ---------- NORMALS ----------
A=r['atts'][1]; T=Table(A['c2v'],A['lmc'],A['eos'],A['vos']); nN=A['num_vertices']
nd2c,nv2d=traverse(T,nN); assert len(nd2c)==nN
pm,tt,comp,scheme,mbl=B.i8(),B.i8(),B.u8(),B.u8(),B.u8(); probs=P.rans_sym_create(B)
ncorr,_,_=P.rans_sym_decode(B,probs,mbl,nN*2)      # corrections are positive: NO zigzag
mq,cv=B.i32(),B.i32(); p0=B.u8(); size=B.varint(); flips=conn.RabsStream(B.b,p0,B.p,size); B.p+=size
qbits=B.b[B.p]  # NORMALS transform data: quantization_bits (read after all portable data of this decoder)
MAXQ=(1<<8)-1; MAXV=MAXQ-1; CEN=MAXV//2
assert mq==MAXQ and cv==CEN,(mq,cv)
def cross(a,b): return [a[1]*b[2]-a[2]*b[1],a[2]*b[0]-a[0]*b[2],a[0]*b[1]-a[1]*b[0]]
def in_diamond(s,t): return abs(s)+abs(t)<=CEN
def invert_diamond(s,t):
    if s>=0 and t>=0: ss,st=1,1
    elif s<=0 and t<=0: ss,st=-1,-1
    else: ss=1 if s>0 else -1; st=1 if t>0 else -1
    cs,ct=ss*CEN,st*CEN
    us=2*s-cs; ut=2*t-ct
    if ss*st>=0: us,ut=-ut,-us
    else: us,ut=ut,us
    us+=cs; ut+=ct
    return tdiv(us,2),tdiv(ut,2)
def rotcount(p):
    x,y=p
    if x==0: return 0 if y==0 else (3 if y>0 else 1)
    if x>0: return 2 if y>=0 else 1
    return 0 if y<=0 else 3
def rot(p,k):
    if k==1: return [p[1],-p[0]]
    if k==2: return [-p[0],-p[1]]
    if k==3: return [-p[1],p[0]]
    return list(p)
def modmax(x):
    if x>CEN: return x-MAXQ
    if x<-CEN: return x+MAXQ
    return x
noct=[0]*(2*nN); nflip=0; small=0
pred3=[None]*nN
for p in range(nN):
    c0=nd2c[p]; cent=P3(c0); n=[0,0,0]
    # VertexCornersIterator over the attribute table: swing left to the end, then right from start
    cs=[]; c=c0; left=True
    while c>=0:
        cs.append(c)
        if left:
            c2=T.swl(c)
            if c2<0: c=T.swr(c0); left=False
            elif c2==c0: c=INV
            else: c=c2
        else: c=T.swr(c)
    for c in cs:
        dn=[a-b for a,b in zip(P3(nxt(c)),cent)]; dp=[a-b for a,b in zip(P3(prv(c)),cent)]
        x=cross(dn,dp); n=[n[k]+x[k] for k in range(3)]
    asum=sum(abs(x) for x in n)
    if asum>(1<<29):
        q=asum//(1<<29); n=[tdiv(x,q) for x in n]
    pred3[p]=list(n)
    # canonicalize
    asum=sum(abs(x) for x in n)
    if asum==0: v=[CEN,0,0]
    else:
        v0=tdiv(n[0]*CEN,asum); v1=tdiv(n[1]*CEN,asum)
        v2=CEN-abs(v0)-abs(v1)
        if n[2]<0: v2=-v2
        v=[v0,v1,v2]
    if flips.bit(): v=[-x for x in v]; nflip+=1
    if v[0]>=0: s,t=v[1]+CEN,v[2]+CEN
    else:
        s=abs(v[2]) if v[1]<0 else MAXV-abs(v[2])
        t=abs(v[1]) if v[2]<0 else MAXV-abs(v[1])
    if (s==0 and t==0) or (s==0 and t==MAXV) or (s==MAXV and t==0): s,t=MAXV,MAXV
    elif s==0 and t>CEN: t=CEN-(t-CEN)
    elif s==MAXV and t<CEN: t=CEN+(CEN-t)
    elif t==MAXV and s<CEN: s=CEN+(CEN-s)
    elif t==0 and s>CEN: s=CEN-(s-CEN)
    pr=[s-CEN,t-CEN]
    ind=in_diamond(*pr)
    if not ind: pr=list(invert_diamond(*pr))
    bl=(pr[0]==0 and pr[1]==0) or (pr[0]<0 and pr[1]<=0)
    rc=rotcount(pr)
    if not bl: pr=rot(pr,rc)
    o=[modmax(pr[0]+ncorr[p*2]),modmax(pr[1]+ncorr[p*2+1])]
    if not bl: o=rot(o,(4-rc)%4)
    if not ind: o=list(invert_diamond(*o))
    noct[p*2]=o[0]+CEN; noct[p*2+1]=o[1]+CEN
def oct2vec(s,t):
    sc=2.0/MAXV; y=s*sc-1.0; z=t*sc-1.0; x=1.0-abs(y)-abs(z); xo=max(0.0,-x)
    y+= xo if y<0 else -xo; z+= xo if z<0 else -xo
    d=math.sqrt(x*x+y*y+z*z); return [x/d,y/d,z/d] if d>1e-3 else [0,0,0]
dots=[]
for p in range(nN):
    g=pred3[p]; gl=math.sqrt(sum(x*x for x in g))
    if gl==0: continue
    v=oct2vec(noct[p*2],noct[p*2+1]); dots.append(sum(v[k]*g[k]/gl for k in range(3)))
dots.sort()
cm=[modmax(x) if x<=CEN else x-MAXQ for x in ncorr]
print('NORMAL: entries',nN,'oct range',min(noct),max(noct),'(0..%d)'%MAXV,'flips',nflip,'| dot(decoded, area-weighted geometric): mean %.4f p01 %.3f p10 %.3f min %.3f'%(sum(dots)/len(dots),dots[len(dots)//100],dots[len(dots)//10],dots[0]),'| mean |corr| (mod) %.2f'%(sum(abs(x) for x in cm)/len(cm)))
print('transform-data byte for NORMAL quantization_bits =',qbits,'| pos dequant',pmin,prange,pbits,'| uv dequant',umin,urange,ubits)

# ---------- end-to-end: UV triangle coverage vs decoded texture atlas ----------
import numpy as np, cv2, os
if os.path.exists('/tmp/basis_probe/slice0.png'):
    tex=cv2.imread('/tmp/basis_probe/slice0.png'); H=tex.shape[0]
    mask_tex=(tex.max(axis=2)>12).astype(np.uint8)
    A=r['atts'][0]; ac2v=A['c2v']; delta=urange/float((1<<ubits)-1)
    U=np.array(uv,dtype=np.float64).reshape(-1,2)*delta+np.array(umin)
    tri=np.array([[uv2d[ac2v[3*f+k]] for k in range(3)] for f in range(F)])
    for name,vv in (('row=v*H',U[:,1]),('row=(1-v)*H',1.0-U[:,1])):
        pts=np.stack([U[:,0]*H,vv*H],axis=1)
        m=np.zeros((H,H),np.uint8)
        for t in tri: cv2.fillConvexPoly(m,np.round(pts[t]).astype(np.int32),1)
        inter=(m&mask_tex).sum(); print('UV coverage',name,': IoU %.3f  (uv-covered %.3f, tex-nonblack %.3f)'%(inter/((m|mask_tex).sum()),m.mean(),mask_tex.mean()))
