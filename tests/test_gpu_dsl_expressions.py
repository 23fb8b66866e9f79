"""Randomised DSL expression semantics on the device: typing (integer-valued literals are Int,
`+ - *` stay Int on Int operands, `/` and `^` are always Real, Bool -> {0,1}), intrinsics, conditional expressions
and comparison on reals — each random expression is printed as DSL text (fully parenthesised), compiled through
DSL -> CUDA C -> NVRTC, evaluated on the GPU as an output equation and compared with a Python evaluation of the
same tree under the reference's rules (pharmsol-dsl/src/analyze.rs:2751-2817, rust_backend.rs:277-466)."""
import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = {"a": 1.75, "b": -0.4, "c": 3.0, "d": 12.5}


class N:
    def __init__(self, kind, *args):
        self.kind, self.args = kind, args


def gen(rng, depth, want="num"):
    """Random tree; `want`: num (Int or Real) | bool."""
    if want == "bool":
        op = rng.choice(["<", "<=", ">", ">=", "==", "!=", "and", "or", "not"]) if depth > 0 else "<"
        if op in ("and", "or"):
            return N(op, gen(rng, depth - 1, "bool"), gen(rng, depth - 1, "bool"))
        if op == "not":
            return N("not", gen(rng, depth - 1, "bool"))
        return N(op, gen(rng, max(depth - 1, 0)), gen(rng, max(depth - 1, 0)))
    if depth <= 0:
        r = rng.random()
        if r < 0.35:
            return N("param", rng.choice(list(PARAMS)))
        if r < 0.65:
            return N("int", rng.randint(0, 6))
        if r < 0.75:
            return N("intlit_real_spelling", rng.randint(1, 5))     # written "3.0": still an Int constant
        return N("real", rng.choice([0.5, 1.25, 2.75, 0.1, 7.5]))
    r = rng.random()
    if r < 0.45:
        return N(rng.choice(["+", "-", "*"]), gen(rng, depth - 1), gen(rng, depth - 1))
    if r < 0.55:
        return N("/", gen(rng, depth - 1), gen(rng, depth - 1))
    if r < 0.60:
        return N("neg", gen(rng, depth - 1))
    if r < 0.70:
        return N("if", gen(rng, depth - 1, "bool"), gen(rng, depth - 1), gen(rng, depth - 1))
    if r < 0.80:
        return N(rng.choice(["min", "max"]), gen(rng, depth - 1), gen(rng, depth - 1))
    if r < 0.85:
        return N("abs", gen(rng, depth - 1))
    if r < 0.90:
        return N("^", gen(rng, depth - 1), N("int", rng.randint(0, 3)))
    f = rng.choice(["exp_s", "ln_s", "sqrt_s", "floor", "ceil", "round", "sin", "cos"])
    return N(f, gen(rng, depth - 1))


def text(n):
    k, a = n.kind, n.args
    if k == "param":
        return a[0]
    if k == "int":
        return str(a[0])
    if k == "intlit_real_spelling":
        return f"{a[0]}.0"
    if k == "real":
        return repr(a[0])
    if k in ("+", "-", "*", "/", "^", "<", "<=", ">", ">=", "==", "!="):
        return f"({text(a[0])} {k} {text(a[1])})"
    if k == "and":
        return f"({text(a[0])} && {text(a[1])})"
    if k == "or":
        return f"({text(a[0])} || {text(a[1])})"
    if k == "not":
        return f"(!{text(a[0])})"
    if k == "neg":
        return f"(-{text(a[0])})"
    if k == "if":
        return f"(if ({text(a[0])}) {{ {text(a[1])} }} else {{ {text(a[2])} }})"
    if k in ("min", "max", "abs", "floor", "ceil", "round", "sin", "cos"):
        return f"{k}({', '.join(text(x) for x in a)})"
    if k == "exp_s":                       # keep magnitudes sane: exp(min(x, 3))
        return f"exp(min({text(a[0])}, 3))"
    if k == "ln_s":
        return f"ln(abs({text(a[0])}) + 1)"
    if k == "sqrt_s":
        return f"sqrt(abs({text(a[0])}))"
    raise KeyError(k)


def ev(n):
    """-> (value, type) with type in {'int', 'real', 'bool'} under the reference's typing rules."""
    k, a = n.kind, n.args
    if k == "param":
        return PARAMS[a[0]], "real"
    if k in ("int", "intlit_real_spelling"):
        return int(a[0]), "int"
    if k == "real":
        return float(a[0]), "real"

    def num(x):
        v, t = ev(x)
        return (1.0 if v else 0.0, "real") if t == "bool" else (v, t)

    def boolean(x):
        v, t = ev(x)
        return bool(v) if t == "bool" else (v != 0)
    if k in ("+", "-", "*"):
        (x, tx), (y, ty) = num(a[0]), num(a[1])
        if tx == "int" and ty == "int":
            return (x + y if k == "+" else x - y if k == "-" else x * y), "int"
        x, y = float(x), float(y)
        return (x + y if k == "+" else x - y if k == "-" else x * y), "real"
    if k == "/":
        (x, _), (y, _) = num(a[0]), num(a[1])
        x, y = float(x), float(y)
        if y == 0.0:
            return (math.nan if x == 0.0 or x != x else math.copysign(math.inf, x) * math.copysign(1.0, y)), "real"
        return x / y, "real"
    if k == "^":
        (x, _), (y, _) = num(a[0]), num(a[1])
        try:
            return math.pow(float(x), float(y)), "real"
        except (OverflowError, ValueError):
            return math.nan, "real"
    if k == "neg":
        v, t = num(a[0])
        return -v, t
    if k in ("<", "<=", ">", ">=", "==", "!="):
        (x, tx), (y, ty) = ev(a[0]), ev(a[1])
        import operator
        op = {"<": operator.lt, "<=": operator.le, ">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne}[k]
        return op(float(x), float(y)), "bool"
    if k == "and":
        return boolean(a[0]) and boolean(a[1]), "bool"
    if k == "or":
        return boolean(a[0]) or boolean(a[1]), "bool"
    if k == "not":
        return not boolean(a[0]), "bool"
    if k == "if":
        c = boolean(a[0])
        (x, tx), (y, ty) = num(a[1]), num(a[2])
        if tx == "int" and ty == "int":
            return (x if c else y), "int"
        return (float(x) if c else float(y)), "real"
    if k in ("min", "max"):
        (x, tx), (y, ty) = num(a[0]), num(a[1])
        f = min if k == "min" else max
        if tx == "int" and ty == "int":
            return f(x, y), "int"
        x, y = float(x), float(y)
        if x != x:
            return y, "real"
        if y != y:
            return x, "real"
        return f(x, y), "real"
    if k == "abs":
        v, t = num(a[0])
        return abs(v), t
    v = float(num(a[0])[0])
    if v != v or math.isinf(v):
        return (math.nan if k in ("sin", "cos") or v != v else v), "real"
    if k == "floor":
        return float(math.floor(v)), "real"
    if k == "ceil":
        return float(math.ceil(v)), "real"
    if k == "round":                                       # half away from zero (Rust f64::round)
        return float(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)), "real"
    if k == "sin":
        return math.sin(v), "real"
    if k == "cos":
        return math.cos(v), "real"
    if k == "exp_s":
        return math.exp(min(v, 3.0)), "real"
    if k == "ln_s":
        return math.log(abs(v) + 1.0), "real"
    if k == "sqrt_s":
        return math.sqrt(abs(v)), "real"
    raise KeyError(k)


@pytest.mark.parametrize("seed", range(20))
def test_random_expressions_match_python_semantics(ps, seed):
    rng = random.Random(1000 + seed)
    exprs = [gen(rng, rng.randint(2, 4)) for _ in range(12)]
    outs = [f"y{k}" for k in range(len(exprs))]
    src = ["name = expr_fuzz_%d" % seed, "kind = ode", "params = " + ", ".join(PARAMS), "states = x", "outputs = " + ", ".join(outs),
           "bolus(iv) -> x", "dx(x) = -a * x"]
    src += [f"out({o}) = {text(e)} ~ continuous()" for o, e in zip(outs, exprs)]
    eq = ps.Equation.from_dsl("\n".join(src) + "\n")
    ops = [("bolus", 0.0, 1.0, "iv")] + [("missing_observation", 1.0, o) for o in outs]
    got = np.array(eq.estimate_predictions(ps.Subject("s", ops), list(PARAMS.values())).flat_predictions())
    for k, e in enumerate(exprs):
        v, t = ev(e)
        want = float(v) if t != "bool" else (1.0 if v else 0.0)
        if want != want:
            assert got[k] != got[k], (text(e), got[k], want)
        elif math.isinf(want):
            assert got[k] == want, (text(e), got[k], want)
        else:
            assert got[k] == pytest.approx(want, rel=1e-12, abs=1e-12), (text(e), got[k], want)
