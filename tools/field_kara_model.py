"""Limb-level model (test infrastructure, not product): one-level Karatsuba on the 8x8 product part (3 x 16 wide
multiply-adds instead of 64) followed by a separated Montgomery reduction that keeps the even/odd accumulator
structure of field.cuh -- the shift of each round, the computation of m and the odd half of m*p fuse into ONE carry
chain (add.cc, mul.lo, madc...), so the reduction costs what it costs inside the interleaved product, and the high
limbs of the product enter one per round at relative column 7.  120 instead of 137 multiply instructions per product
for roughly +100 integer adds; not implemented in CUDA yet.  Carry-loss assertions included."""
import random
P=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
M32=0xffffffff
INV=(-pow(P,-1,1<<32))%(1<<32)
pl=[(P>>(32*i))&M32 for i in range(8)]
def limbs(v,n=8): return [(v>>(32*i))&M32 for i in range(n)]
def val(l): return sum(x<<(32*i) for i,x in enumerate(l))

def add_chain(dst, src, carry_in=0):
    """dst[i] += src[i] (+carry) over len(src); returns carry out"""
    c=carry_in
    for i in range(len(src)):
        s=dst[i]+src[i]+c; dst[i]=s&M32; c=s>>32
    return c

def mul4x4(x,y):
    """8-limb product of two 4-limb numbers via two interleaved accumulators (16 wide products)"""
    ev=[0]*8; od=[0]*8   # od[k] is column k+1
    for j in range(4):
        # even limbs (x0,x2) * y_j at columns j, j+2 ; odd limbs (x1,x3) * y_j at columns j+1, j+3
        pe=[x[0]*y[j], x[2]*y[j]]; po=[x[1]*y[j], x[3]*y[j]]
        def mad(acc, k, prods):
            c=0
            for pr in prods:
                for part in (pr&M32, pr>>32):
                    s=acc[k]+part+c; acc[k]=s&M32; c=s>>32; k+=1
            while c and k<8:
                s=acc[k]+c; acc[k]=s&M32; c=s>>32; k+=1
            assert c==0
        if j%2==0: mad(ev,j,pe); mad(od,j,po)        # od index j <-> column j+1
        else:      mad(od,j-1,pe); mad(ev,j+1,po)    # column j (odd) <-> od index j-1 ; column j+1 even
    r=ev[:]
    c=add_chain(r[1:], od[:7])   # r[1:] is a copy; redo in place
    r=limbs(val(ev)+(val(od)<<32),8)
    assert val(ev)+(val(od)<<32) < 1<<256
    return r

def kara(a,b):
    al,ah,bl,bh=a[:4],a[4:],b[:4],b[4:]
    z0=mul4x4(al,bl); z2=mul4x4(ah,bh)
    da=val(al)-val(ah); db=val(bl)-val(bh)
    sa=da<0; sb=db<0
    zm=mul4x4(limbs(abs(da),4),limbs(abs(db),4))
    # middle = z0 + z2 - (da*db) ; da*db = (+/-) zm
    mid=val(z0)+val(z2) - (val(zm) if sa==sb else -val(zm))
    assert 0<=mid< 1<<258
    T=val(z0)+(mid<<128)+(val(z2)<<256)
    assert T==val(a)*val(b)
    return limbs(T,16)

def chain(x, start, prods):
    c=0;k=start
    for pr in prods:
        for part in (pr&M32, pr>>32):
            s=x[k]+part+c; x[k]=s&M32; c=s>>32; k+=1
    return c

def sos_reduce(T):
    pe=[pl[0],pl[2],pl[4],pl[6]]; po=[pl[1],pl[3],pl[5],pl[7]]
    A=T[:8]; B=[0]*8       # A column-0 aligned, B column-1 aligned (B[k] = column k+1)
    m=(A[0]*INV)&M32
    c=chain(B,0,[x*m for x in po]); assert c==0
    c=chain(A,0,[x*m for x in pe]); B[7]+=c; assert B[7]<=M32
    for i in range(1,8):
        # shift: B[0] += A[1] ; m from the new column 0 ; A'[k] = A[k+2] + odd(p)*m (carry chain) ; feed T[7+i] at relative column 7
        s=B[0]+A[1]; B[0]=s&M32; cc=s>>32
        m=(B[0]*INV)&M32
        src=A[2:]+[0,0]; out=[0]*8; prods=[x*m for x in po]
        for k in range(8):
            pr=prods[k//2]; part=(pr&M32) if k%2==0 else (pr>>32)
            s=src[k]+part+cc; out[k]=s&M32; cc=s>>32
        assert cc==0
        A=out                                   # now the column-1 aligned accumulator
        c=chain(B,0,[x*m for x in pe]); A[7]+=c; assert A[7]<=M32
        # incoming limb of T: absolute column 7+i is relative column 7 of the column-0 aligned accumulator B -- wait: base is abs column i
        s=B[7]+T[7+i]; B[7]=s&M32; A[7]+=s>>32; assert A[7]<=M32
        A,B=B,A
    # final: A column-0 aligned with A[0]==0 after last round? (value/2^32 = (A>>32) + B) + the last limb T[15]
    assert A[0]==0
    r=(val(A)>>32)+val(B)+(T[15]<<(32*7))
    return r

Rinv=pow(1<<256,-1,P)
random.seed(3)
for t in range(3000):
    a=random.randrange(P) if t>3 else [P-1,0,1,P-2][t]
    b=random.randrange(P) if t>3 else [P-1,P-1,1,2][t]
    T=kara(limbs(a),limbs(b))
    r=sos_reduce(T)
    assert r<2*P, hex(r)
    if r>=P: r-=P
    assert r==a*b*Rinv%P, (hex(a),hex(b))
print("kara + SOS reduction model ok")
