# Throwaway survey probe: verify BasisLZ (ETC1S) global-data layout on the example .ktx2. NOT product code.
import struct,sys
class Bits:
    def __init__(s,b): s.b=b; s.pos=0  # bit position, LSB-first
    def get(s,n):
        if n==0: return 0
        byte=s.pos>>3; v=int.from_bytes(s.b[byte:byte+5],'little')>>(s.pos&7); s.pos+=n; return v&((1<<n)-1)
    def bytes_used(s): return (s.pos+7)>>3
SORTED=[17,18,19,20,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15,16]
class Huff:
    def __init__(s,sizes):
        s.sizes=sizes; s.used=sum(1 for x in sizes if x)
        # canonical codes, then bit-reverse for LSB-first reading
        cnt=[0]*17
        for x in sizes: cnt[x]+=1
        cnt[0]=0; nxt=[0]*18; code=0
        for l in range(1,17): code=(code+cnt[l-1])<<1; nxt[l]=code
        s.map={}
        for sym,l in enumerate(sizes):
            if l:
                c=nxt[l]; nxt[l]+=1
                r=int(bin(c)[2:].zfill(l)[::-1],2)
                s.map[(l,r)]=sym
        s.maxl=max(sizes) if sizes else 0
    def dec(s,B):
        if s.used==0: return 0
        c=0
        for l in range(1,s.maxl+1):
            c|=B.get(1)<<(l-1)
            if (l,c) in s.map: return s.map[(l,c)]
        raise Exception('bad huff')
def read_huff(B):
    total=B.get(14)
    if total==0: return Huff([])
    ncl=B.get(5); assert 1<=ncl<=21
    cls=[0]*21
    for i in range(ncl): cls[SORTED[i]]=B.get(3)
    H=Huff(cls); sizes=[]
    while len(sizes)<total:
        c=H.dec(B)
        if c<=16: sizes.append(c)
        elif c==17: sizes+= [0]*(B.get(3)+3)
        elif c==18: sizes+= [0]*(B.get(7)+11)
        elif c==19: sizes+= [sizes[-1]]*(B.get(2)+3)
        elif c==20: sizes+= [sizes[-1]]*(B.get(7)+7)
    assert len(sizes)==total,(len(sizes),total)
    return Huff(sizes)
def probe(fn):
    b=open(fn,'rb').read()
    vk,ts,w,h,d,layers,faces,levels,sc=struct.unpack_from('<9I',b,12)
    sgdOff,sgdLen=struct.unpack_from('<2Q',b,64)
    ec,scnt,eb,sb,tb,xb=struct.unpack_from('<HHIIII',b,sgdOff)
    n=max(layers,1)*faces*levels
    p=sgdOff+20+20*n
    ep=b[p:p+eb]; sel=b[p+eb:p+eb+sb]; tab=b[p+eb+sb:p+eb+sb+tb]
    out={}
    # endpoints
    B=Bits(ep)
    m0,m1,m2,mi=read_huff(B),read_huff(B),read_huff(B),read_huff(B)
    gray=B.get(1)
    prev=[16,16,16]; pinten=0; eps=[]
    for i in range(ec):
        pinten=(pinten+mi.dec(B))&7
        for c in range(1 if gray else 3):
            m = m0 if prev[c]<=9 else (m1 if prev[c]<=21 else m2)
            prev[c]=(prev[c]+m.dec(B))&31
        if gray: prev[1]=prev[2]=prev[0]
        eps.append((prev[0],prev[1],prev[2],pinten))
    out['endpoints']=dict(count=ec,bytes=eb,bytes_used=B.bytes_used(),gray=gray,model_syms=[m0.used,m1.used,m2.used,mi.used],first=eps[:4])
    # selectors
    B=Bits(sel)
    glob=B.get(1); hyb=B.get(1); raw=B.get(1)
    sels=[]
    if not glob and not hyb:
        if raw:
            for i in range(scnt): sels.append([B.get(8) for _ in range(4)])
        else:
            dm=read_huff(B); prevb=[0]*4
            for i in range(scnt):
                if i==0: prevb=[B.get(8) for _ in range(4)]
                else: prevb=[prevb[j]^dm.dec(B) for j in range(4)]
                sels.append(list(prevb))
    out['selectors']=dict(count=scnt,bytes=sb,bytes_used=B.bytes_used(),global_=glob,hybrid=hyb,raw=raw,first=sels[:3])
    # tables
    B=Bits(tab)
    epm,dem,sm,rle=read_huff(B),read_huff(B),read_huff(B),read_huff(B)
    hist=B.get(13)
    out['tables']=dict(bytes=tb,bytes_used=B.bytes_used(),endpoint_pred_syms=len(epm.sizes),delta_endpoint_syms=len(dem.sizes),selector_syms=len(sm.sizes),selector_rle_syms=len(rle.sizes),selector_history_buf_size=hist)
    return out
if __name__=='__main__':
    r=probe(sys.argv[1])
    for k,v in r.items(): print(k,v)

def vlc(B,cb):
    v=0;ofs=0
    while True:
        ch=B.get(cb+1); v|=(ch&((1<<cb)-1))<<ofs; ofs+=cb
        if not ch&(1<<cb): return v
def slices(fn,maxslices=99):
    b=open(fn,'rb').read()
    vk,ts,w,h,d,layers,faces,levels,sc=struct.unpack_from('<9I',b,12)
    sgdOff,sgdLen=struct.unpack_from('<2Q',b,64)
    lvOff,lvLen,lvUnc=struct.unpack_from('<3Q',b,80)
    ec,scnt,eb,sb,tb,xb=struct.unpack_from('<HHIIII',b,sgdOff)
    n=max(layers,1)*faces*levels
    descs=[struct.unpack_from('<5I',b,sgdOff+20+20*i) for i in range(n)]
    p=sgdOff+20+20*n
    tab=b[p+eb+sb:p+eb+sb+tb]
    B=Bits(tab); epm,dem,sm,rle=read_huff(B),read_huff(B),read_huff(B),read_huff(B); hist=B.get(13)
    bx,by=(w+3)//4,(h+3)//4
    prev_frame=None; res=[]
    for si,(flags,off,ln,aoff,aln) in enumerate(descs[:maxslices]):
        B=Bits(b[lvOff+off:lvOff+off+ln])
        hb=[0]*hist; rover=hist//2
        rle_cnt=0; prev_sym=0; rep=0; prev_ep=0
        rows=[[[0,0] for _ in range(bx)] for _ in range(2)]
        cur_frame=[[None]*bx for _ in range(by)]
        stats=dict(pred=[0,0,0,0],sel_direct=0,sel_hist=0,sel_rle_blocks=0)
        for y in range(by):
            cur=y&1
            for x in range(bx):
                if x&1==0:
                    if y&1==0:
                        if rep: rep-=1; bits=prev_sym
                        else:
                            bits=epm.dec(B)
                            if bits==256: rep=vlc(B,4)+3-1; bits=prev_sym
                            else: prev_sym=bits
                        rows[cur^1][x][1]=bits>>4
                    else: bits=rows[cur][x][1]
                pred=bits&3; bits>>=2; stats['pred'][pred]+=1
                cr=False
                if pred==0: assert x>0; e=prev_ep
                elif pred==1: assert y>0; e=rows[cur^1][x][0]
                elif pred==2:
                    assert prev_frame is not None,'CR pred in I-frame'
                    e,s=prev_frame[y][x]; cr=True
                else:
                    e=dem.dec(B)+prev_ep
                    if e>=ec: e-=ec
                rows[cur][x][0]=e; prev_ep=e
                if not cr:
                    if rle_cnt>0: rle_cnt-=1; s=hb[0]; stats['sel_rle_blocks']+=1
                    else:
                        s=sm.dec(B)
                        if s==scnt+hist:
                            r=rle.dec(B)
                            rle_cnt=(vlc(B,7)+3) if r==63 else r+3
                            s=hb[0]; rle_cnt-=1; stats['sel_rle_blocks']+=1
                        elif s>=scnt:
                            i=s-scnt; s=hb[i]; stats['sel_hist']+=1
                            if i: hb[i//2],hb[i]=hb[i],hb[i//2]
                        else:
                            stats['sel_direct']+=1
                            hb[rover]=s; rover+=1
                            if rover==hist: rover=hist//2
                assert e<ec and s<scnt,(e,s)
                cur_frame[y][x]=(e,s)
        prev_frame=cur_frame
        res.append(dict(slice=si,flags=flags,bytes=ln,bytes_used=B.bytes_used(),**stats))
    return res
if __name__=='__main__' and len(sys.argv)>2:
    for r in slices(sys.argv[1],int(sys.argv[2])): print(r)
