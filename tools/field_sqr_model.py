# Python model of field.cuh's even/odd CIOS product and of the proposed squaring, limb-exact with carry flags.
import random
P=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
M32=0xffffffff
INV=(-pow(P,-1,1<<32))%(1<<32)
pl=[(P>>(32*i))&M32 for i in range(8)]
def limbs(v): return [(v>>(32*i))&M32 for i in range(8)]
def val(l): return sum(x<<(32*i) for i,x in enumerate(l))

def chain(x, start, prods, top=None):
    """x[start..] += products (list of 64-bit values at consecutive 2-limb slots), one carry chain; carry out added to top (or must be 0)."""
    c=0; k=start
    for pr in prods:
        lo,hi=pr&M32,pr>>32
        s=x[k]+lo+c; x[k]=s&M32; c=s>>32; k+=1
        s=x[k]+hi+c; x[k]=s&M32; c=s>>32; k+=1
    return c
def mad_even(x, top, a_even, b, z=0):  # a_even = [a0,a2,a4,a6]; skip first z products
    c=chain(x, 2*z, [a*b for a in a_even[z:]])
    return top+c  # new top
def mad_odd(y, a_odd, b, z=0):
    c=chain(y, 2*z, [a*b for a in a_odd[z:]])
    assert c==0, "mad_odd carry out"
def shift_mad_odd(e, x0, a_odd, b, z=0):
    # x0 += e[1]; e'[k] = e[k+2] + products (carry chain); returns new x0
    s=x0+e[1]; x0=s&M32; c=s>>32
    src=e[2:]+[0,0]
    out=[0]*8
    prods=[a*b for a in a_odd]
    for k in range(8):
        pr=prods[k//2] if (k//2)>=z else 0
        part=(pr&M32) if k%2==0 else (pr>>32)
        s=src[k]+part+c; out[k]=s&M32; c=s>>32
    assert c==0, "shift_mad_odd carry out"
    e[:]=out
    return x0

def mont(a,b,rows=None):
    """generic: rows[i] = (even limbs, odd limbs, multiplier, ze, zo) ; default = plain product"""
    al=limbs(a); bl=limbs(b)
    if rows is None:
        rows=[([al[0],al[2],al[4],al[6]],[al[1],al[3],al[5],al[7]],bl[i],0,0) for i in range(8)]
    pe=[pl[0],pl[2],pl[4],pl[6]]; po=[pl[1],pl[3],pl[5],pl[7]]
    ev=[0]*8; od=[0]*8
    e_,o_,b_,ze,zo=rows[0]
    chain(ev,0,[x*b_ for x in e_]); chain(od,0,[x*b_ for x in o_])
    m=(ev[0]*INV)&M32
    mad_odd(od,po,m); od[7]=mad_even(ev,od[7],pe,m); assert od[7]<=M32
    A,B=ev,od  # A: column-0 aligned with A[0]==0 ; B: column-1 aligned
    for i in range(1,8):
        e_,o_,b_,ze,zo=rows[i]
        B[0]=shift_mad_odd(A,B[0],o_,b_,zo)          # A becomes the odd accumulator
        A[7]=mad_even(B,A[7],e_,b_,ze); assert A[7]<=M32
        m=(B[0]*INV)&M32
        mad_odd(A,po,m); A[7]=mad_even(B,A[7],pe,m); assert A[7]<=M32
        A,B=B,A
    # after round 7: A is column-0 aligned with A[0]==0, B column-1 aligned: value/2^32 = (A>>32) + B
    r=(val(A)>>32)+val(B)
    if r>=P: r-=P
    return r

def sqr_rows(a):
    al=limbs(a)
    d=[0]*8
    for j in range(1,8): d[j]=((al[j]<<1)|(al[j-1]>>31))&M32
    rows=[]
    for i in range(8):
        v=[0]*8
        v[i]=al[i]
        if i+1<8: v[i+1]=d[i+1]&~1&M32
        for j in range(i+2,8): v[j]=d[j]
        ze=(i+1)//2; zo=i//2
        assert all(v[2*k]==0 for k in range(ze)) and all(v[2*k+1]==0 for k in range(zo))
        rows.append(([v[0],v[2],v[4],v[6]],[v[1],v[3],v[5],v[7]],al[i],ze,zo))
    return rows

Rinv=pow(1<<256,-1,P)
random.seed(1)
for t in range(3000):
    a=random.randrange(P) if t>5 else [0,1,P-1,P-2,(1<<254)-1-P+P-1 if False else P-1, 2][t]
    b=random.randrange(P)
    assert mont(a,b)==a*b*Rinv%P
    assert mont(a,a,sqr_rows(a))==a*a*Rinv%P, hex(a)
print("model ok: even/odd CIOS product and the 36-product squaring agree with big-int arithmetic")

# ---- mul4_add: a*b + c*d + e*f + g*h with ONE reduction (bound check with worst-case inputs) ----------------
def mont_sum(pairs):
    pe=[pl[0],pl[2],pl[4],pl[6]]; po=[pl[1],pl[3],pl[5],pl[7]]
    L=[(limbs(a),limbs(b)) for a,b in pairs]
    ev=[0]*8; od=[0]*8
    first=True
    A,B=ev,od
    for i in range(8):
        for k,(al,bl) in enumerate(L):
            e_=[al[0],al[2],al[4],al[6]]; o_=[al[1],al[3],al[5],al[7]]
            if i==0:
                if k==0:
                    chain(A,0,[x*bl[0] for x in e_]); chain(B,0,[x*bl[0] for x in o_])
                else:
                    mad_odd(B,o_,bl[0]); B[7]=mad_even(A,B[7],e_,bl[0]); assert B[7]<=M32
            else:
                if k==0:
                    B[0]=shift_mad_odd(A,B[0],o_,bl[i])   # A becomes the odd accumulator
                    A[7]=mad_even(B,A[7],e_,bl[i]); assert A[7]<=M32
                else:
                    mad_odd(A,o_,bl[i]); A[7]=mad_even(B,A[7],e_,bl[i]); assert A[7]<=M32
        if i==0:
            m=(A[0]*INV)&M32
            mad_odd(B,po,m); B[7]=mad_even(A,B[7],pe,m); assert B[7]<=M32
        else:
            m=(B[0]*INV)&M32
            mad_odd(A,po,m); A[7]=mad_even(B,A[7],pe,m); assert A[7]<=M32
            A,B=B,A
    r=(val(A)>>32)+val(B)
    assert r<2*P, "more than one subtraction needed"
    if r>=P: r-=P
    return r
for t in range(2000):
    if t<4: vals=[P-1]*8
    elif t<8: vals=[P-1 if (t>>k)&1 else random.randrange(P) for k in range(8)]
    else: vals=[random.randrange(P) for _ in range(8)]
    pairs=[(vals[0],vals[1]),(vals[2],vals[3]),(vals[4],vals[5]),(vals[6],vals[7])]
    exp=sum(a*b for a,b in pairs)*Rinv%P
    assert mont_sum(pairs)==exp
    assert mont_sum(pairs[:2])==sum(a*b for a,b in pairs[:2])*Rinv%P
print("model ok: four-product sum with one reduction (no lost carry at the worst case (p-1)^2 x 4, result < 2p)")
