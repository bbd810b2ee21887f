# Throwaway probe: full Draco v2.2 valence-edgebreaker connectivity decode, to validate the survey's restatement. NOT product code.
import sys,struct; sys.path.insert(0,'/tmp/draco_probe')
import probe as P
INV=-1
def nxt(c): return INV if c<0 else (c-2 if c%3==2 else c+1)
def prv(c): return INV if c<0 else (c+2 if c%3==0 else c-1)
class RabsStream:
    def __init__(s,b,p0,off,size): s.b=b; s.off=off; s.p=256-p0; s.st,s.bo=P.rabs_init(b,off,size)
    def bit(s):
        if s.st<4096 and s.bo>0: s.bo-=1; s.st=s.st*256+s.b[s.off+s.bo]
        q,r=divmod(s.st,256); xn=q*s.p
        if r<s.p: s.st=xn+r; return 1
        s.st=s.st-xn-s.p; return 0
def decode(fn):
    b=open(fn,'rb').read(); B=P.Buf(b,11)
    trav=B.u8(); nv=B.varint(); nf=B.varint(); nad=B.u8(); nsym=B.varint(); nsplit=B.varint()
    nts=B.varint(); last=0; splits=[]
    for i in range(nts):
        d=B.varint(); src=last+d; d2=B.varint(); splits.append([src,src-d2,0]); last=src
    if nts:
        for i in range(nts): splits[i][2]=(b[B.p+(i>>3)]>>(i&7))&1   # bit decoder LSB-first, 1 bit per event: source_edge
        B.p+=(nts+7)//8
    def rabs(): 
        p0=B.u8(); size=B.varint(); r=RabsStream(b,p0,B.p,size); B.p+=size; return r
    start_faces=rabs(); seams=[rabs() for _ in range(nad)]
    log=[]; ctx=[]
    for i in range(6):
        n=B.varint(); ctx.append(P.decode_symbols(B,n,1,log,'ctx',want=True) if n>0 else [])
    cnt=[len(x) for x in ctx]
    end_conn=B.p
    maxv=nv+nsplit; F=nf
    opp=[INV]*(3*F); c2v=[INV]*(3*F); lmc=[]; hole=[True]*maxv; val=[0]*maxv
    def addv(): lmc.append(INV); return len(lmc)-1
    def setopp(a,c): opp[a]=c; opp[c]=a
    def swl(c): return nxt(opp[nxt(c)]) if c>=0 and opp[nxt(c)]>=0 else INV
    def swr(c): return prv(opp[prv(c)]) if c>=0 and opp[prv(c)]>=0 else INV
    SYM=['C','S','L','R','E']
    active=-1; last_sym=None
    stack=[]; split_active={}
    numf=0; symhist={k:0 for k in SYM}
    init_cfg=[]; init_corners=[]
    # topology split lookup: events consumed from the back (encoder order sorted ascending by source id)
    ts=list(splits)
    for sid in range(nsym):
        face=numf; numf+=1
        if active!=-1:
            cnt[active]-=1; assert cnt[active]>=0
            s=SYM[ctx[active][cnt[active]]]
        else: s='E'
        symhist[s]+=1; chk=False; c0=3*face
        if s=='C':
            a=stack[-1]; vx=c2v[nxt(a)]; bcr=nxt(lmc[vx]); assert a!=bcr and opp[a]<0 and opp[bcr]<0
            setopp(a,c0+1); setopp(bcr,c0+2)
            vap=c2v[prv(a)]; vbn=c2v[nxt(bcr)]; assert vx!=vap and vx!=vbn
            c2v[c0]=vx; c2v[c0+1]=vbn; c2v[c0+2]=vap; lmc[vap]=c0+2; hole[vx]=False; stack[-1]=c0
        elif s in 'RL':
            a=stack[-1]; assert opp[a]<0
            if s=='R': oc,cl,cr=c0+2,c0+1,c0
            else: oc,cl,cr=c0+1,c0,c0+2
            setopp(oc,a); nvx=addv(); assert len(lmc)<=maxv
            c2v[oc]=nvx; lmc[nvx]=oc; vr=c2v[prv(a)]; c2v[cr]=vr; lmc[vr]=cr; c2v[cl]=c2v[nxt(a)]; stack[-1]=c0; chk=True
        elif s=='S':
            bcr=stack.pop()
            if sid in split_active: stack.append(split_active[sid])
            a=stack[-1]; assert a!=bcr and opp[a]<0 and opp[bcr]<0
            setopp(a,c0+2); setopp(bcr,c0+1)
            vp=c2v[prv(a)]; c2v[c0]=vp; c2v[c0+1]=c2v[nxt(a)]; vbp=c2v[prv(bcr)]; c2v[c0+2]=vbp; lmc[vbp]=c0+2
            cn=nxt(bcr); vn=c2v[cn]; val[vp]+=val[vn]; lmc[vp]=lmc[vn]
            first=cn
            while cn>=0:
                c2v[cn]=vp; cn=swl(cn); assert cn!=first
            lmc[vn]=INV; stack[-1]=c0
        else:
            v0=addv(); v1=addv(); v2=addv(); assert len(lmc)<=maxv
            c2v[c0]=v0;c2v[c0+1]=v1;c2v[c0+2]=v2; lmc[v0]=c0;lmc[v1]=c0+1;lmc[v2]=c0+2; stack.append(c0); chk=True
        # NewActiveCornerReached
        c=stack[-1]; n_=nxt(c); p_=prv(c)
        if s in 'CS': val[c2v[n_]]+=1; val[c2v[p_]]+=1
        elif s=='R': val[c2v[c]]+=1; val[c2v[n_]]+=1; val[c2v[p_]]+=2
        elif s=='L': val[c2v[c]]+=1; val[c2v[n_]]+=2; val[c2v[p_]]+=1
        else: val[c2v[c]]+=2; val[c2v[n_]]+=2; val[c2v[p_]]+=2
        active=min(7,max(2,val[c2v[n_]]))-2
        if chk:
            enc_id=nsym-sid-1
            while ts and ts[-1][0]==enc_id:
                src,spl,edge=ts.pop()
                top=stack[-1]
                nac = nxt(top) if edge==1 else prv(top)   # RIGHT_FACE_EDGE=1
                split_active[nsym-spl-1]=nac
    assert len(lmc)<=maxv
    while stack:
        corner=stack.pop()
        if start_faces.bit():
            a=corner; vn=c2v[nxt(a)]; cb=nxt(lmc[vn]); vx=c2v[nxt(cb)]; cc=nxt(lmc[vx])
            assert len({a,cb,cc})==3 and opp[a]<0 and opp[cb]<0 and opp[cc]<0
            vp=c2v[nxt(cc)]; face=numf; numf+=1; nc=3*face
            setopp(nc,a); setopp(nc+1,cb); setopp(nc+2,cc)
            c2v[nc]=vx; c2v[nc+1]=vp; c2v[nc+2]=vn
            for k in range(3): hole[c2v[nc+k]]=False
            init_cfg.append(True); init_corners.append(nc)
        else: init_cfg.append(False); init_corners.append(corner)
    assert numf==F,(numf,F)
    assert all(x==0 for x in cnt),cnt
    # attribute seams
    seam_corners=[[] for _ in range(nad)]
    for f in range(F):
        for c in (3*f,3*f+1,3*f+2):
            o=opp[c]
            if o<0:
                for i in range(nad): seam_corners[i].append(c)
                continue
            if o//3<f: continue
            for i in range(nad):
                if seams[i].bit(): seam_corners[i].append(c)
    V=len(lmc)
    atts=[]
    for i in range(nad):
        eos=[False]*(3*F); vos=[False]*V
        for c in seam_corners[i]:
            eos[c]=True; vos[c2v[nxt(c)]]=True; vos[c2v[prv(c)]]=True
            o=opp[c]
            if o>=0: eos[o]=True; vos[c2v[nxt(o)]]=True; vos[c2v[prv(o)]]=True
        def aopp(c): return INV if (c<0 or eos[c]) else opp[c]
        def aswl(c):
            x=aopp(nxt(c)); return nxt(x) if x>=0 else INV
        ac2v=[INV]*(3*F); nnew=0; alm=[]
        for v in range(V):
            c=lmc[v]
            if c<0: continue
            fid=nnew; nnew+=1; fc=c
            if vos[v]:
                a=aswl(fc)
                while a>=0:
                    fc=a; a=aswl(a)
                    assert a!=c
            ac2v[fc]=fid; alm.append(fc)
            a=swr(fc)
            while a>=0 and a!=fc:
                if eos[nxt(a)]:
                    fid=nnew; nnew+=1; alm.append(a)
                ac2v[a]=fid; a=swr(a)
        atts.append(dict(num_vertices=nnew,c2v=ac2v,lmc=alm,eos=eos,vos=vos,interior_seams=sum(1 for c in seam_corners[i] if opp[c]>=0)))
    # AssignPointsToCorners
    c2p=[INV]*(3*F); p2c=[]
    for v in range(V):
        c=lmc[v]
        if c<0: continue
        dfc=c
        if not hole[v]:
            for A in atts:
                vid=A['c2v'][c]; a=swr(c); found=False
                while a!=c:
                    assert a>=0
                    if A['c2v'][a]!=vid: dfc=a; found=True; break
                    a=swr(a)
                if found: break
        c=dfc; c2p[c]=len(p2c); p2c.append(c); pc=c; c=swr(c)
        while c>=0 and c!=dfc:
            if any(A['c2v'][c]!=A['c2v'][pc] for A in atts): c2p[c]=len(p2c); p2c.append(c)
            else: c2p[c]=c2p[pc]
            pc=c; c=swr(c)
    return dict(lmc=lmc,p2c=p2c,hole=hole,B=B,b=b,nv=nv,nf=F,V=V,isolated=sum(1 for x in lmc if x<0),symhist=symhist,start_faces=init_cfg,boundary_verts=sum(hole[:V]),atts=atts,num_points=len(p2c),c2p=c2p,c2v=c2v,opp=opp,end_conn=end_conn)
if __name__=='__main__':
    r=decode(sys.argv[1])
    print({k:v for k,v in r.items() if k in('nv','nf','V','isolated','symhist','start_faces','boundary_verts','num_points','end_conn')})
    for i,A in enumerate(r['atts']): print('att_data',i,'num_attr_vertices',A['num_vertices'],'interior seam edges',A['interior_seams'])
