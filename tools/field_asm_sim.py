# Executes the inline-asm carry chains of field.cuh in Python (operand numbering included) and runs Fe::sqr's exact
# call sequence against big-integer arithmetic.
import re, random, sys
src=open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'ethrex_b200', 'csrc', 'field.cuh')).read()
P=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
M32=0xffffffff
INV=0xe4866389
pl=[(P>>(32*i))&M32 for i in range(8)]

# ---- parse detail:: functions -------------------------------------------------------------------------------
funcs={}
for m in re.finditer(r'B2_D void (\w+)\(([^)]*)\)\s*\{\s*asm\((.*?)\);\s*\}', src, re.S):
    name,params,body=m.group(1),m.group(2),m.group(3)
    # split asm text and constraints
    parts=re.split(r'\n\s*:\s', body)
    text="".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0]))
    text=text.replace('\\n','\n').replace('\\t',' ')
    outs=re.findall(r'"([+=]r)"\(([^)]*)\)', parts[1]) if len(parts)>1 else []
    ins=re.findall(r'"(r)"\(([^)]*)\)', parts[2]) if len(parts)>2 else []
    pnames=[p.strip().split()[-1].lstrip('*&') for p in params.split(',')]
    funcs[name]=(pnames,text,[o[1] for o in outs],[i[1] for i in ins])

def run(name, env):
    """env: dict var -> python object (list for arrays, [value] boxes for scalars by reference, ints for values)"""
    pnames,text,outs,ins=funcs[name]
    ops=outs+ins
    def get(expr):
        m=re.match(r'(\w+)\[(\d+)\]$',expr)
        if m: return env[m.group(1)][int(m.group(2))]
        v=env[expr]; return v[0] if isinstance(v,list) else v
    def put(expr,val):
        m=re.match(r'(\w+)\[(\d+)\]$',expr)
        if m: env[m.group(1)][int(m.group(2))]=val; return
        assert isinstance(env[expr],list), expr
        env[expr][0]=val
    vals=[get(o) for o in ops]
    cc=0
    for ins_ in [i.strip() for i in text.replace('\n',';').split(';') if i.strip()]:
        mm=re.match(r'([\w.]+)\s+(.*)$',ins_); opc=mm.group(1); args=[a.strip() for a in mm.group(2).split(',')]
        def rd(a): return vals[int(a[1:])] if a.startswith('%') else int(a,0)
        dst=int(args[0][1:])
        base=opc.replace('.u32','')
        if base in('mul.lo','mul.hi'):
            pr=rd(args[1])*rd(args[2]); vals[dst]=(pr&M32) if base=='mul.lo' else pr>>32
        elif base.startswith('mad') :
            hi='.hi' in base; carry_in=base.startswith('madc'); carry_out='.cc' in base
            pr=rd(args[1])*rd(args[2]); part=(pr>>32) if hi else (pr&M32)
            s=part+rd(args[3])+(cc if carry_in else 0)
            vals[dst]=s&M32
            if carry_out: cc=s>>32
            else: assert s>>32==0, (name,ins_,"carry lost")
        elif base.startswith('add'):
            carry_in=base.startswith('addc'); carry_out='.cc' in base
            s=rd(args[1])+rd(args[2])+(cc if carry_in else 0)
            vals[dst]=s&M32
            if carry_out: cc=s>>32
            else: assert s>>32==0, (name,ins_,"carry lost")
        else: raise SystemExit("unknown op "+opc)
    for o,v in zip(outs,vals): put(o,v)

# ---- parse Fe::sqr body: sequence of detail:: calls and m = X[0] * Cfg::INV --------------------------------------
body=re.search(r'static B2_D Fe sqr\(const Fe& a\) \{(.*?)Fe r;', src, re.S).group(1)
stmts=[s.strip() for s in re.sub(r'//[^\n]*','',body).split(';') if s.strip()]

def sqr_sim(a):
    al=[(a>>(32*i))&M32 for i in range(8)]
    env={'ev':[0]*8,'od':[0]*8}
    sc={}  # scalar values
    for j in range(1,8): sc[f'd{j}']=((al[j]<<1)|(al[j-1]>>31))&M32
    def ev_expr(e):
        e=e.strip()
        m=re.match(r'a\.v\[(\d)\]$',e)
        if m: return al[int(m.group(1))]
        m=re.match(r'(d\d) & ~1u$',e)
        if m: return sc[m.group(1)]&~1&M32
        m=re.match(r'Cfg::mod\((\d)\)$',e)
        if m: return pl[int(m.group(1))]
        if e in sc: return sc[e]
        if e=='m': return sc['m']
        raise SystemExit("expr? "+e)
    for s in stmts:
        if s.startswith('uint32_t') : continue
        m=re.match(r'm = (\w+)\[0\] \* Cfg::INV$',s)
        if m: sc['m']=(env[m.group(1)][0]*INV)&M32; continue
        m=re.match(r'detail::(\w+)\((.*)\)$',s,re.S)
        assert m, s
        name=m.group(1); args=[x.strip() for x in re.split(r',(?![^\[]*\])',m.group(2))]
        pnames=funcs[name][0]
        local={}
        for pn,arg in zip(pnames,args):
            if arg in('ev','od'): local[pn]=env[arg]
            elif re.match(r'(ev|od)\[\d\]$',arg):
                arr,idx=re.match(r'(\w+)\[(\d)\]',arg).groups(); local[pn]=('ref',env[arr],int(idx))
            else: local[pn]=ev_expr(arg)
        # materialise refs as boxes
        e2={}
        for k,v in local.items():
            if isinstance(v,tuple): e2[k]=[v[1][v[2]]]
            else: e2[k]=v
        run(name,e2)
        for k,v in local.items():
            if isinstance(v,tuple): v[1][v[2]]=e2[k][0]
    ev,od=env['ev'],env['od']
    r=sum(ev[i]<<(32*i) for i in range(8))+sum(od[i]<<(32*(i-1)) for i in range(1,8))
    if r>=P: r-=P
    return r
Rinv=pow(1<<256,-1,P)
random.seed(7)
N_RANDOM = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tests=[0,1,2,P-1,P-2,(P-1)>>1,0x80000000,sum(0x80000000<<(32*k) for k in range(8))%P,sum(0xffffffff<<(32*k) for k in range(7))]+[random.randrange(P) for _ in range(N_RANDOM)]
for a in tests:
    assert sqr_sim(a)==a*a*Rinv%P, hex(a)
print("asm-level simulation of Fe::sqr: %d inputs ok"%len(tests))
