"""CPU execution of DEVICE code: parses ethrex_b200/csrc/field.cuh, runs the inline-PTX carry chains (operand
numbering, carry flag, lost-carry assertions) in Python along the exact statement sequence of the function bodies
of Fe::mul, Fe::sqr, Fe::mul2_add and Fe::mul4_add (loops and conditions of the C++ are translated, not re-written),
and compares every result with big-integer arithmetic -- for both moduli.

    python tools/field_asm_sim.py [n_random]          (tests/test_field_asm_model.py runs it)

Test infrastructure: nothing in the product imports it.
"""
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "ethrex_b200", "csrc", "field.cuh")).read()
M32 = 0xFFFFFFFF


def _cfg(name):
    body = re.search(r"struct %s \{(.*?)\n\};" % name, SRC, re.S).group(1)
    mod = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", re.search(r"mod\(int i\) \{.*?\{(.*?)\}", body, re.S).group(1))]
    inv = int(re.search(r"INV = (0x[0-9a-f]+)u", body).group(1), 16)
    p = sum(v << (32 * i) for i, v in enumerate(mod))
    assert (p * ((-inv) % (1 << 32))) & M32 == 1 or (p * inv + 1) & M32 == 0
    return p, mod, inv


# ---- the asm primitives of namespace detail ----------------------------------------------------------------------
FUNCS = {}
for m in re.finditer(r"B2_D (?:void|uint32_t) (\w+)\(([^)]*)\)\s*\{(?:\s*uint32_t borrow;)?\s*asm\((.*?)\);\s*(?:return borrow;\s*)?\}", SRC, re.S):
    name, params, body = m.group(1), m.group(2), m.group(3)
    parts = re.split(r"\n\s*:\s", body)
    text = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0])).replace("\\n", "\n").replace("\\t", " ")
    outs = re.findall(r'"([+=]r)"\(([^)]*)\)', parts[1]) if len(parts) > 1 else []
    ins = re.findall(r'"(r)"\(([^)]*)\)', parts[2]) if len(parts) > 2 else []
    plist = [p.strip() for p in params.split(",")]
    pnames = [p.split()[-1].lstrip("*&") for p in plist]
    is_ref = [("&" in p and "const" not in p) for p in plist]
    FUNCS[name] = (pnames, is_ref, text, [o[1] for o in outs], [i[1] for i in ins])


def run_asm(name, env):
    pnames, _, text, outs, ins = FUNCS[name]
    ops = outs + ins

    def get(expr):
        mm = re.match(r"(\w+)\[(\d+)\]$", expr)
        if mm:
            return env[mm.group(1)][int(mm.group(2))]
        v = env[expr]
        return v[0] if isinstance(v, list) else v

    def put(expr, val):
        mm = re.match(r"(\w+)\[(\d+)\]$", expr)
        if mm:
            env[mm.group(1)][int(mm.group(2))] = val
        else:
            env[expr][0] = val

    vals = [get(o) for o in ops]
    cc = 0
    for ins_ in [i.strip() for i in text.replace("\n", ";").split(";") if i.strip()]:
        mm = re.match(r"([\w.]+)\s+(.*)$", ins_)
        opc, args = mm.group(1).replace(".u32", ""), [a.strip() for a in mm.group(2).split(",")]
        rd = lambda a: vals[int(a[1:])] if a.startswith("%") else int(a, 0)  # noqa: E731
        dst = int(args[0][1:])
        if opc in ("mul.lo", "mul.hi"):
            pr = rd(args[1]) * rd(args[2])
            vals[dst] = (pr & M32) if opc == "mul.lo" else pr >> 32
            continue
        if opc.startswith("mad"):
            pr = rd(args[1]) * rd(args[2])
            s = ((pr >> 32) if ".hi" in opc else (pr & M32)) + rd(args[3]) + (cc if opc.startswith("madc") else 0)
        elif opc.startswith("add"):
            s = rd(args[1]) + rd(args[2]) + (cc if opc.startswith("addc") else 0)
        elif opc.startswith("sub"):
            s = rd(args[1]) - rd(args[2]) - (cc if opc.startswith("subc") else 0)
            vals[dst] = s & M32
            if ".cc" in opc:
                cc = 1 if s < 0 else 0
            continue
        else:
            raise SystemExit("unknown opcode " + opc)
        vals[dst] = s & M32
        if ".cc" in opc:
            cc = s >> 32
        else:
            assert s >> 32 == 0, (name, ins_, "a carry is lost here")
    for o, v in zip(outs, vals):
        put(o, v)


# ---- translate a member function body (the part before `Fe r;`) to Python -----------------------------------------
def translate(fn_name, signature_re):
    m = re.search(signature_re + r" \{(.*?)\n    Fe r;", SRC, re.S)
    assert m, fn_name
    body = re.sub(r"//[^\n]*", "", m.group(1))
    body = re.sub(r"#pragma unroll", "", body)
    body = re.sub(r"= \{([^{}]*)\};", r"= LIST(\1);", body)  # array initialisers would confuse the brace tracker
    out, depth = [], 1
    # statement splitter that keeps for(...) headers intact
    toks = re.findall(r"\s*(for \([^)]*\) \{|if \([^)]*\) \{|\}|[^;{}]+;)", body)
    for t in toks:
        t = t.strip()
        ind = "    " * depth
        if t == "}":
            depth -= 1
            continue
        mm = re.match(r"for \(int (\w) = (\d+); \1 < (\d+); (?:\1 \+= (\d+)|\+\+\1)\) \{", t)
        if mm:
            out.append(f"{ind}for {mm.group(1)} in range({mm.group(2)}, {mm.group(3)}, {mm.group(4) or 1}):")
            depth += 1
            continue
        mm = re.match(r"if \((.*)\) \{", t)
        if mm:
            out.append(f"{ind}if {mm.group(1)}:")
            depth += 1
            continue
        t = t.rstrip(";").strip()
        if re.match(r"uint32_t (ev\[8\], od\[8\](, m)?|m)$", t):
            continue
        mm = re.match(r"const Fe\* const (\w)\[4\] = LIST\((.*)\)$", t)
        if mm:
            out.append(f"{ind}{mm.group(1)} = [{mm.group(2).replace('&', '')}]")
            continue
        mm = re.match(r"uint32_t (d1 = .*)$", t, re.S)
        if mm:  # the doubled limbs of sqr: d_j = funnelshift_l(a[j-1], a[j], 1)
            for dj, lo, hi in re.findall(r"(d\d) = __funnelshift_l\(a\.v\[(\d)\], a\.v\[(\d)\], 1\)", t):
                out.append(f"{ind}{dj} = ((a[{hi}] << 1) | (a[{lo}] >> 31)) & M32")
            continue
        mm = re.match(r"m = (\w+)\[0\] \* Cfg::INV$", t)
        if mm:
            out.append(f"{ind}m = ({mm.group(1)}[0] * INV) & M32")
            continue
        mm = re.match(r"detail::(\w+)\((.*)\)$", t, re.S)
        assert mm, (fn_name, t)
        args = [x.strip() for x in re.split(r",(?![^\[]*\])", mm.group(2))]
        py = []
        for a in args:
            a = re.sub(r"(\w)\.v\[([^\]]+)\]", r"\1[\2]", a)            # a.v[i] -> a[i]
            a = re.sub(r"(\w)\[(\w)\]->v\[([^\]]+)\]", r"\1[\2][\3]", a)  # x[k]->v[i] -> x[k][i]
            a = re.sub(r"Cfg::mod\((\d)\)", r"PL[\1]", a)
            a = re.sub(r"~1u", "0xFFFFFFFE", a)
            py.append(a)
        out.append(f"{ind}call({mm.group(1)!r}, {py!r}, locals())")
    return "def fn(a, b, c, d, e, f, g, h, PL, INV, call, M32):\n    ev = [0] * 8\n    od = [0] * 8\n" + "\n".join(out) + "\n    return ev, od\n"


def call(name, arg_src, scope):
    pnames, is_ref, _, _, _ = FUNCS[name]
    env, back = {}, []
    for pn, ref, src in zip(pnames, is_ref, arg_src):
        if src in ("ev", "od"):
            env[pn] = scope[src]
        elif ref:
            arr, idx = re.match(r"(\w+)\[(\d+)\]$", src).groups()
            env[pn] = [scope[arr][int(idx)]]
            back.append((pn, scope[arr], int(idx)))
        else:
            env[pn] = eval(src, {}, scope) & M32  # noqa: S307 (translated field.cuh expressions only)
    run_asm(name, env)
    for pn, arr, idx in back:
        arr[idx] = env[pn][0]


SIGS = {
    "mul": r"static B2_D Fe mul\(const Fe& a, const Fe& b\)",
    "sqr": r"static B2_D Fe sqr\(const Fe& a\)",
    "mul2_add": r"static B2_D Fe mul2_add\(const Fe& a, const Fe& b, const Fe& c, const Fe& d\)",
    "mul4_add": r"static B2_D Fe mul4_add\(const Fe& a, const Fe& b, const Fe& c, const Fe& d, const Fe& e, const Fe& f, const Fe& g, const Fe& h\)",
}
COMPILED = {}
for k, sig in SIGS.items():
    ns = {}
    exec(translate(k, sig), ns)  # noqa: S102 (source generated from field.cuh by translate())
    COMPILED[k] = ns["fn"]


def limbs(v):
    return [(v >> (32 * i)) & M32 for i in range(8)]


def run(fn, p, pl, inv, *vals):
    ops = [limbs(v) for v in vals] + [[0] * 8] * (8 - len(vals))
    ev, od = COMPILED[fn](*ops, pl, inv, call, M32)
    # both tails of field.cuh: r = ev + (od >> 32), then one conditional subtraction
    r = sum(ev[i] << (32 * i) for i in range(8)) + sum(od[i] << (32 * (i - 1)) for i in range(1, 8))
    assert r < 2 * p, (fn, "more than one subtraction needed")
    return r - p if r >= p else r


def main():
    n_random = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    rng = random.Random(7)
    total = 0
    for cfg in ("FqCfg", "FrCfg"):
        p, pl, inv = _cfg(cfg)
        rinv = pow(1 << 256, -1, p)
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) >> 1, 0x80000000, sum(0x80000000 << (32 * k) for k in range(8)) % p,
                sum(0xFFFFFFFF << (32 * k) for k in range(7)), (1 << 253) % p]
        cases = [(x, y) for x in edge for y in edge[:5]] + [(rng.randrange(p), rng.randrange(p)) for _ in range(n_random)]
        for x, y in cases:
            assert run("mul", p, pl, inv, x, y) == x * y * rinv % p, (cfg, "mul", hex(x), hex(y))
            assert run("sqr", p, pl, inv, x) == x * x * rinv % p, (cfg, "sqr", hex(x))
            total += 2
        quads = [(p - 1,) * 8] + [tuple(rng.choice(edge) for _ in range(8)) for _ in range(20)] + \
                [tuple(rng.randrange(p) for _ in range(8)) for _ in range(max(20, n_random // 4))]
        for q in quads:
            assert run("mul2_add", p, pl, inv, *q[:4]) == (q[0] * q[1] + q[2] * q[3]) * rinv % p, (cfg, "mul2_add")
            assert run("mul4_add", p, pl, inv, *q) == (q[0] * q[1] + q[2] * q[3] + q[4] * q[5] + q[6] * q[7]) * rinv % p, (cfg, "mul4_add")
            total += 2
    print(f"asm-level simulation of Fe::mul / sqr / mul2_add / mul4_add (Fq and Fr): {total} inputs ok")


if __name__ == "__main__":
    main()
