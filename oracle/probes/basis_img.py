# Throwaway probe: ETC1S slice -> RGB image for visual verification.
import sys,struct; sys.path.insert(0,'/tmp/basis_probe')
import numpy as np, cv2
import probe as P
from probe import Bits,read_huff,vlc
fn=sys.argv[1]; nsl=int(sys.argv[2])
b=open(fn,'rb').read()
vk,ts,w,h,d,layers,faces,levels,sc=struct.unpack_from('<9I',b,12)
sgdOff,sgdLen=struct.unpack_from('<2Q',b,64); lvOff,lvLen,_=struct.unpack_from('<3Q',b,80)
ec,scnt,eb,sb,tb,xb=struct.unpack_from('<HHIIII',b,sgdOff)
n=max(layers,1)*faces*levels
descs=[struct.unpack_from('<5I',b,sgdOff+20+20*i) for i in range(n)]
p=sgdOff+20+20*n
# palettes
B=Bits(b[p:p+eb]); m0,m1,m2,mi=read_huff(B),read_huff(B),read_huff(B),read_huff(B); gray=B.get(1)
prev=[16,16,16]; pin=0; E=np.zeros((ec,4),np.int32)
for i in range(ec):
    pin=(pin+mi.dec(B))&7
    for c in range(3):
        m=m0 if prev[c]<=9 else (m1 if prev[c]<=21 else m2); prev[c]=(prev[c]+m.dec(B))&31
    E[i]=(prev[0],prev[1],prev[2],pin)
B=Bits(b[p+eb:p+eb+sb]); B.get(3); dm=read_huff(B); S=np.zeros((scnt,4,4),np.uint8); pb=[0]*4
for i in range(scnt):
    pb=[B.get(8) for _ in range(4)] if i==0 else [pb[j]^dm.dec(B) for j in range(4)]
    for y in range(4):
        for x in range(4): S[i,y,x]=(pb[y]>>(2*x))&3
INT=np.array([[-8,-2,2,8],[-17,-5,5,17],[-29,-9,9,29],[-42,-13,13,42],[-60,-18,18,60],[-80,-24,24,80],[-106,-33,33,106],[-183,-47,47,183]],np.int32)
base=(E[:,:3]<<3)|(E[:,:3]>>2)
COL=np.clip(base[:,None,:]+INT[E[:,3]][:,:,None],0,255).astype(np.uint8)   # [ec,4 selectors,3]
B=Bits(b[p+eb+sb:p+eb+sb+tb]); epm,dem,sm,rle=read_huff(B),read_huff(B),read_huff(B),read_huff(B); hist=B.get(13)
bx,by=w//4,h//4; prevf=None
for si in range(nsl):
    flags,off,ln,_,_=descs[si]; B=Bits(b[lvOff+off:lvOff+off+ln])
    hb=[0]*hist; rover=hist//2; rc=0; ps=0; rep=0; pe=0
    rows=[[[0,0] for _ in range(bx)] for _ in range(2)]
    EI=np.zeros((by,bx),np.int32); SI=np.zeros((by,bx),np.int32)
    for y in range(by):
        cur=y&1
        for x in range(bx):
            if x&1==0:
                if y&1==0:
                    if rep: rep-=1; bits=ps
                    else:
                        bits=epm.dec(B)
                        if bits==256: rep=vlc(B,4)+2; bits=ps
                        else: ps=bits
                    rows[cur^1][x][1]=bits>>4
                else: bits=rows[cur][x][1]
            pred=bits&3; bits>>=2; cr=False
            if pred==0: e=pe
            elif pred==1: e=rows[cur^1][x][0]
            elif pred==2: e=int(prevf[0][y,x]); s=int(prevf[1][y,x]); cr=True
            else:
                e=dem.dec(B)+pe
                if e>=ec: e-=ec
            rows[cur][x][0]=e; pe=e
            if not cr:
                if rc>0: rc-=1; s=hb[0]
                else:
                    s=sm.dec(B)
                    if s==scnt+hist:
                        r_=rle.dec(B); rc=(vlc(B,7)+3) if r_==63 else r_+3; s=hb[0]; rc-=1
                    elif s>=scnt:
                        i=s-scnt; s=hb[i]
                        if i: hb[i//2],hb[i]=hb[i],hb[i//2]
                    else:
                        hb[rover]=s; rover+=1
                        if rover==hist: rover=hist//2
            EI[y,x]=e; SI[y,x]=s
    prevf=(EI,SI)
    sel=S[SI]                                  # [by,bx,4,4]
    img=COL[EI[:,:,None,None],sel]             # [by,bx,4,4,3]
    img=img.transpose(0,2,1,3,4).reshape(h,w,3)
    cv2.imwrite(f'/tmp/basis_probe/slice{si}.png',cv2.cvtColor(cv2.resize(img,(512,512),interpolation=cv2.INTER_AREA),cv2.COLOR_RGB2BGR))
    print('slice',si,'flags',flags,'bytes used',B.bytes_used(),'/',ln,'mean rgb',img.reshape(-1,3).mean(0).round(1))
