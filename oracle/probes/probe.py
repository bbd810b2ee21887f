# Throwaway survey probe: walk a .drc far enough to list features used. NOT product code.
import struct,sys
class Buf:
    def __init__(s,b,p=0): s.b=b; s.p=p
    def u8(s): v=s.b[s.p]; s.p+=1; return v
    def i8(s): v=struct.unpack_from('<b',s.b,s.p)[0]; s.p+=1; return v
    def u16(s): v=struct.unpack_from('<H',s.b,s.p)[0]; s.p+=2; return v
    def i32(s): v=struct.unpack_from('<i',s.b,s.p)[0]; s.p+=4; return v
    def u32(s): v=struct.unpack_from('<I',s.b,s.p)[0]; s.p+=4; return v
    def f32(s): v=struct.unpack_from('<f',s.b,s.p)[0]; s.p+=4; return v
    def varint(s):
        r=0;sh=0
        while True:
            c=s.u8(); r|=(c&0x7f)<<sh; sh+=7
            if not c&0x80: return r
def rabs_init(b,off,size):
    x=b[off+size-1]>>6
    if x==0: bo=size-1; st=b[off+size-1]&0x3f
    elif x==1: bo=size-2; st=int.from_bytes(b[off+size-2:off+size],'little')&0x3fff
    elif x==2: bo=size-3; st=int.from_bytes(b[off+size-3:off+size],'little')&0x3fffff
    else: raise Exception('bad rabs')
    return [st+4096,bo]
def rabs_bits(b,off,size,p0,n):
    st,bo=rabs_init(b,off,size); out=[]
    p=256-p0
    for _ in range(n):
        if st<4096 and bo>0: bo-=1; st=st*256+b[off+bo]
        q,r=divmod(st,256); xn=q*p
        if r<p: st=xn+r; out.append(1)
        else: st=st-xn-p; out.append(0)
    return out
def skip_rabs(B,name,log):
    p0=B.u8(); size=B.varint(); off=B.p; B.p+=size
    log.append(f'  rANS-bit[{name}] prob_zero={p0} bytes={size}')
    return (p0,off,size)
def rans_sym_create(B):
    n=B.varint(); probs=[0]*n; i=0
    while i<n:
        d=B.u8(); t=d&3
        if t==3:
            off=d>>2
            i+=off+1
        else:
            pr=d>>2
            for k in range(t): pr|=B.u8()<<(8*(k+1)-2)
            probs[i]=pr; i+=1
    return probs
def rans_sym_decode(B,probs,nbits_unique,count):
    pb=(3*nbits_unique)//2; pb=max(12,min(20,pb)); prec=1<<pb; lbase=prec*4
    nbytes=B.varint(); off=B.p; B.p+=nbytes
    if count==0 or nbytes==0: return [],nbytes,pb
    b=B.b
    x=b[off+nbytes-1]>>6
    k=x+1
    st=int.from_bytes(b[off+nbytes-k:off+nbytes],'little')&((1<<(8*k-2))-1)
    bo=nbytes-k; st+=lbase
    cum=[0]*(len(probs)+1)
    for i,p in enumerate(probs): cum[i+1]=cum[i]+p
    assert cum[-1]==prec,(cum[-1],prec)
    lut=[0]*prec
    for i,p in enumerate(probs):
        for j in range(cum[i],cum[i+1]): lut[j]=i
    out=[]
    for _ in range(count):
        while st<lbase and bo>0: bo-=1; st=st*256+b[off+bo]
        q,r=divmod(st,prec); s=lut[r]; st=q*probs[s]+r-cum[s]; out.append(s)
    return out,nbytes,pb
def decode_symbols(B,num_values,nc,log,name,want=False):
    scheme=B.u8()
    if scheme==0:
        probs=rans_sym_create(B)
        assert num_values<10**9,'TAGGED scheme with unknown count'
        tags,nbytes,pb=rans_sym_decode(B,probs,5,num_values//nc)
        bits=sum(tags)*nc
        start=B.p
        vals=None
        if want:
            big=int.from_bytes(B.b[start:start+(bits+7)//8],'little'); vals=[];bp=0
            for t in tags:
                for c in range(nc): vals.append((big>>bp)&((1<<t)-1)); bp+=t
        B.p+=(bits+7)//8
        hist={}
        for t in tags: hist[t]=hist.get(t,0)+1
        log.append(f'  symbols[{name}] TAGGED n={num_values} nc={nc} tag_syms={len(probs)} tag_rans_bytes={nbytes} raw_bits={bits} ({(bits+7)//8} B) taghist={dict(sorted(hist.items()))}')
        return vals
    elif scheme==1:
        mbl=B.u8(); probs=rans_sym_create(B)
        if want:
            vals,nbytes,pb=rans_sym_decode(B,probs,mbl,num_values)
        else:
            nbytes=B.varint(); B.p+=nbytes; vals=None; pb=max(12,min(20,3*mbl//2))
        log.append(f'  symbols[{name}] RAW n={num_values if num_values<10**9 else "?"} max_bit_length={mbl} alphabet={len(probs)} rans_precision_bits={pb} rans_bytes={nbytes}')
        return vals
    else: raise Exception('scheme %d'%scheme)
PRED={-2:'NONE',0:'DIFFERENCE',1:'MESH_PARALLELOGRAM',2:'MESH_MULTI_PARALLELOGRAM',3:'MESH_TEX_COORDS_DEPRECATED',4:'MESH_CONSTRAINED_MULTI_PARALLELOGRAM',5:'MESH_TEX_COORDS_PORTABLE',6:'MESH_GEOMETRIC_NORMAL'}
XFORM={-1:'NONE',0:'DELTA',1:'WRAP',2:'NORMAL_OCTAHEDRON',3:'NORMAL_OCTAHEDRON_CANONICALIZED'}
ATT={0:'POSITION',1:'NORMAL',2:'COLOR',3:'TEX_COORD',4:'GENERIC'}
DT={1:'INT8',2:'UINT8',3:'INT16',4:'UINT16',5:'INT32',6:'UINT32',7:'INT64',8:'UINT64',9:'FLOAT32',10:'FLOAT64',11:'BOOL'}
SEQ={0:'GENERIC',1:'INTEGER',2:'QUANTIZATION',3:'NORMALS'}
def probe(fn,verbose=True):
    b=open(fn,'rb').read(); B=Buf(b); log=[]; R={}
    assert b[:5]==b'DRACO'; B.p=5
    maj,mino,etype,meth,flags=B.u8(),B.u8(),B.u8(),B.u8(),B.u16()
    R.update(version=(maj,mino),etype=etype,method=meth,flags=flags,size=len(b))
    log.append(f'header v{maj}.{mino} geometry_type={etype}(1=TRIANGULAR_MESH) method={meth}(1=EDGEBREAKER) flags=0x{flags:04x}')
    assert (maj,mino)==(2,2) and meth==1 and not flags&0x8000
    trav=B.u8(); nv=B.varint(); nf=B.varint(); nad=B.u8(); nsym=B.varint(); nsplit=B.varint()
    R.update(traversal=trav,num_encoded_vertices=nv,num_faces=nf,num_attribute_data=nad,num_symbols=nsym,num_split_symbols=nsplit)
    log.append(f'connectivity: traversal_type={trav}(0=STANDARD,2=VALENCE) num_encoded_vertices={nv} num_faces={nf} num_attribute_data={nad} num_encoded_symbols={nsym} num_split_symbols={nsplit}')
    nts=B.varint(); last=0; splits=[]
    for i in range(nts):
        d=B.varint(); src=last+d; d2=B.varint(); splits.append((src,src-d2)); last=src
    if nts>0: B.p+=(nts+7)//8
    R['topology_splits']=nts
    log.append(f'topology split events={nts} ends@{B.p}')
    c0=B.p
    if trav==0:
        tsz=B.varint(); log.append(f'  traversal symbol buffer bytes={tsz}'); B.p+=tsz; R['trav_bytes']=tsz
    skip_rabs(B,'start_faces',log)
    for a in range(nad): skip_rabs(B,f'attr_seams[{a}]',log)
    if trav==2:
        R['ctx_counts']=[]
        for ctx in range(6):
            n=B.varint(); R['ctx_counts'].append(n)
            if n>0: decode_symbols(B,n,1,log,f'valence_ctx[{ctx+2}]')
            else: log.append(f'  valence_ctx[{ctx+2}] empty')
    R['connectivity_bytes']=B.p-11
    log.append(f'connectivity section total bytes={B.p-11} (of {len(b)})')
    # attributes
    ndec=B.u8(); decs=[]
    for i in range(ndec):
        adid=B.i8(); dtype=B.u8(); tm=B.u8(); decs.append(dict(att_data_id=adid,decoder_type=dtype,traversal_method=tm))
    log.append(f'num_attributes_decoders={ndec} '+str(decs)+'  (decoder_type 0=MESH_VERTEX_ATTRIBUTE 1=MESH_CORNER_ATTRIBUTE; traversal 0=DEPTH_FIRST 1=PREDICTION_DEGREE)')
    for d in decs:
        na=B.varint(); d['atts']=[]
        for j in range(na):
            at,dt,ncmp,norm=B.u8(),B.u8(),B.u8(),B.u8(); uid=B.varint()
            d['atts'].append(dict(type=ATT.get(at,at),dtype=DT.get(dt,dt),nc=ncmp,normalized=norm,unique_id=uid))
        for j in range(na): d['atts'][j]['seq_decoder']=SEQ.get(B.u8())
    for d in decs: log.append(f' decoder att_data_id={d["att_data_id"]}: {d["atts"]}')
    R['decoders']=decs
    R['attr_data_start']=B.p
    return R,log,B,decs
def walk_attr_data(R,log,B,decs,npoints_per_dec):
    # npoints_per_dec: number of encoded entries per decoder; unknown for corner attrs without connectivity decode -> we attempt using value counts carried by stream? Draco does not store them; skip when unknown.
    for di,d in enumerate(decs):
        n=npoints_per_dec[di]
        if n is None: n=10**9  # unknown count: only RAW scheme is skippable
        for a in d['atts']:
            s0=B.p
            pm=B.i8(); tt=None
            if pm!=-2: tt=B.i8()
            comp=B.u8()
            nc=a['nc'] if a['seq_decoder']!='NORMALS' else 2
            log.append(f' att {a["type"]}: prediction={PRED.get(pm,pm)} transform={XFORM.get(tt,tt)} compressed={comp} portable_components={nc}')
            if comp>0: decode_symbols(B,n*nc,nc,log,a['type'])
            else:
                nb=B.u8(); B.p+=nb*n*nc; log.append(f'  raw ints bytes_per={nb}')
            # prediction data
            if pm==5:
                no=B.i32(); skip_rabs(B,'texcoord_orientations n=%d'%no,log)
            if tt==1:
                mn,mx=B.i32(),B.i32(); log.append(f'  wrap transform min={mn} max={mx}')
            if tt in (2,3):
                mq,cv=B.i32(),B.i32(); log.append(f'  octahedron transform max_quantized_value={mq} center_value={cv}')
            if pm==6: skip_rabs(B,'normal_flip_bits',log)
            a['bytes']=B.p-s0
            log.append(f'  -> attribute portable data bytes={B.p-s0}')
        for a in d['atts']:
            if a['seq_decoder']=='QUANTIZATION':
                mins=[B.f32() for _ in range(a['nc'])]; rng=B.f32(); qb=B.u8()
                a.update(min=mins,range=rng,qbits=qb); log.append(f' dequant {a["type"]}: min={mins} range={rng} quantization_bits={qb}')
            elif a['seq_decoder']=='NORMALS':
                qb=B.u8(); a['qbits']=qb; log.append(f' dequant NORMAL octahedral quantization_bits={qb}')
        log.append(f' decoder {di} ends@{B.p}')
if __name__=='__main__':
    fn=sys.argv[1]
    R,log,B,decs=probe(fn)
    try:
        walk_attr_data(R,log,B,decs,[R['num_encoded_vertices']]+[None]*(len(decs)-1))
    except Exception as e: log.append('walk failed: %r'%e)
    print('\n'.join(log)); print('pos',B.p,'of',len(B.b))
