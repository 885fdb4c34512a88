"""Tiny Fortran-subset -> Python transpiler used ONLY to generate golden vectors.

The container has no Fortran compiler, so the reference's numerical kernels
(src/derive.f90, src/filters.f90, src/schemes.f90, parts of src/poisson.f90,
src/tools.f90, src/stretching.f90) cannot be compiled.  Their bodies are plain
F77-style loops, so this module reads the reference source text where it lies
(/root/reference, never copied into the repo), rewrites each subroutine into a
Python function operating on 1-based numpy wrappers, and executes it.  The
numbers that come out are outputs of the reference's own statements; they are
committed as fixtures under tests/golden/ by make_golden.py.

Loops over the two non-line directions are executed as whole-array numpy
operations (the loop variable is bound to a full slice) -- legal because those
loops carry no dependence; loops along the line direction stay sequential so
the recurrences are evaluated in the reference's order.

This file is test infrastructure.  It is not imported by the product.
"""
from __future__ import annotations

import re
import numpy as np

ALL = slice(None)


class FArr:
    """1-based (or arbitrary lower bound) Fortran-like view of a numpy array."""

    __array_priority__ = 100

    def __init__(self, a, lb=None):
        self.a = a
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim

    def _ix(self, idx):
        if idx is Ellipsis:
            return Ellipsis
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for d, v in enumerate(idx):
            if isinstance(v, slice):
                lo = None if v.start is None else v.start - self.lb[d]
                hi = None if v.stop is None else v.stop - self.lb[d] + 1
                out.append(slice(lo, hi, v.step))
            else:
                out.append(int(v) - self.lb[d])
        return tuple(out)

    def __getitem__(self, idx):
        return self.a[self._ix(idx)]

    def __setitem__(self, idx, val):
        self.a[self._ix(idx)] = val


def farr(shape, dtype=float, lb=None):
    return FArr(np.zeros(shape, dtype=dtype, order="F"), lb)


# ---------------------------------------------------------------------------
_OPS = [
    (r"\.eq\.", "=="), (r"\.ne\.", "!="), (r"\.ge\.", ">="), (r"\.le\.", "<="),
    (r"\.gt\.", ">"), (r"\.lt\.", "<"), (r"\.and\.", " and "), (r"\.or\.", " or "),
    (r"\.not\.", " not "), (r"\.true\.", "True"), (r"\.false\.", "False"),
    (r"/=", "!="),
]


def _strip_comment(line):
    out, instr = [], None
    for ch in line:
        if instr:
            out.append(ch)
            if ch == instr:
                instr = None
        elif ch in "'\"":
            instr = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def _logical_lines(text):
    """comment-stripped, continuation-joined, ';'-split, lower-cased lines"""
    raw = []
    skip = 0
    for ln in text.split("\n"):
        s = ln.strip()
        if s.startswith("#if"):
            skip += 1
            continue
        if s.startswith("#endif"):
            skip -= 1
            continue
        if s.startswith("#else"):
            continue
        if skip:
            continue
        raw.append(_strip_comment(ln).rstrip())
    joined, cur = [], ""
    for ln in raw:
        s = ln.strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].strip()
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        joined.append(cur)
        cur = ""
    out = []
    for ln in joined:
        for part in ln.split(";"):
            part = part.strip()
            if part:
                out.append(part.lower())
    return out


def _match_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced: " + s)


def _split_args(s):
    args, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


_FUNCS = {"sin", "cos", "tan", "exp", "sqrt", "abs", "atan", "acos", "asin", "log",
          "max", "min", "real", "rl", "iy", "cx", "cmplx", "aimag", "mod", "int",
          "dble", "float", "sinh", "cosh", "tanh", "conjg", "nint", "sign", "dacos"}


class Transpiler:
    def __init__(self, arrays_hint=()):
        self.arrays_hint = set(arrays_hint)

    # -- expressions ------------------------------------------------------
    def expr(self, s, arrays, intctx=False):
        """intctx: the expression is an array subscript or a do-loop bound, where Fortran's `/` between
        integers is an integer division"""
        s = s.strip()
        for pat, rep in _OPS:
            s = re.sub(pat, rep, s)
        if intctx:
            s = re.sub(r"(?<![/!=<>])/(?![/=])", "//", s)
        # kind suffixes and d-exponents
        s = re.sub(r"(\d\.?)_mytype\b", r"\1", s)
        s = re.sub(r"(\d\.?\d*)d([+-]?\d+)", r"\1e\2", s)
        s = re.sub(r"(?<![\w.])(\d+)\.(?![\d\w])", r"\1.0", s)
        out, i = "", 0
        while i < len(s):
            m = re.match(r"[a-z_][a-z0-9_%]*", s[i:])
            if m and (i == 0 or not (s[i - 1].isalnum() or s[i - 1] in "_.")):
                name = m.group(0)
                j = i + len(name)
                k = j
                while k < len(s) and s[k] == " ":
                    k += 1
                if k < len(s) and s[k] == "(" and name not in ("and", "or", "not"):
                    e = _match_paren(s, k)
                    inner = s[k + 1:e]
                    args = [self.index_arg(a, arrays, intctx=(name in arrays)) for a in _split_args(inner)]
                    pyname = name.replace("%", "__")
                    if name in arrays:
                        out += f"{pyname}[{', '.join(args)}]"
                    elif name == "real":
                        out += f"_real({args[0]})"
                    elif name in ("cmplx",):
                        out += f"cx({args[0]}, {args[1]})"
                    else:
                        out += f"{pyname}({', '.join(args)})"
                    i = e + 1
                    continue
                out += name.replace("%", "__")
                i = j
                continue
            out += s[i]
            i += 1
        return out

    def index_arg(self, a, arrays, intctx=False):
        a = a.strip()
        if a == ":":
            return "ALL"
        m = re.match(r"^(.*?):(.*)$", a)
        if m and "(" not in a:
            lo = self.expr(m.group(1), arrays, intctx) if m.group(1).strip() else "None"
            hi = self.expr(m.group(2), arrays, intctx) if m.group(2).strip() else "None"
            return f"slice({lo}, {hi})"
        a = re.sub(r"^kind\s*=\s*mytype$", "None", a)
        if a in ("mytype",):
            return "None"
        return self.expr(a, arrays, intctx)

    # -- a subroutine -----------------------------------------------------
    def subroutine(self, text, vector_vars=(), extra_arrays=()):
        lines = _logical_lines(text)
        head = lines[0]
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)", head)
        if m:
            name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
        else:
            m = re.match(r"subroutine\s+(\w+)", head)
            name, args = m.group(1), []
        arrays = set(extra_arrays) | self.arrays_hint
        body = []
        local_arrays = []
        int_scalars = set()
        for ln in lines[1:]:
            if re.match(r"(integer|real|complex|logical|character|type)\b", ln) and "::" in ln:
                decl, names = ln.split("::", 1)
                if decl.strip().startswith("integer") and "dimension" not in decl and "parameter" not in decl:
                    for item in _split_args(names):
                        if "(" not in item and "=" not in item:
                            int_scalars.add(item.strip())
                dim = re.search(r"dimension\s*\(", decl)
                dimspec = None
                if dim:
                    k0 = decl.index("(", dim.start())
                    dimspec = decl[k0 + 1:_match_paren(decl, k0)]
                is_cplx = "int" if decl.strip().startswith("integer") else decl.strip().startswith("complex")
                for item in _split_args(names):
                    nm = re.match(r"(\w+)", item).group(1)
                    spec = dimspec
                    if "(" in item:
                        k0 = item.index("(")
                        spec = item[k0 + 1:_match_paren(item, k0)]
                    if spec is not None:
                        arrays.add(nm)
                        if nm not in args and ":" not in spec.replace("::", ""):
                            local_arrays.append((nm, spec, is_cplx))
                        elif nm not in args and "allocatable" not in decl:
                            local_arrays.append((nm, spec, is_cplx))
                continue
            if re.match(r"(use|implicit|private|public|intent|contains|save)\b", ln):
                continue
            if re.match(r"\d+\s+format", ln):
                continue
            body.append(ln)
        py = [f"def {name}({', '.join(args)}):"]
        ind = 1
        vec = set(vector_vars)

        def emit(s):
            py.append("    " * ind + s)

        def stmt(ln):
            nonlocal ind
            if re.match(r"end\s*subroutine", ln):
                return
            if ln == "return":
                emit("return locals()")
                return
            if ln.startswith("stop"):
                emit("raise RuntimeError('stop')")
                return
            m = re.match(r"call\s+(\w+)\s*(\((.*)\))?$", ln)
            if m:
                cname = m.group(1)
                if cname in ("prepare", "inversion5_v1", "inversion5_v2", "matrice_refinement",
                             "stretching", "stretching_full", "abxyz", "waves"):
                    a = [self.index_arg(x, arrays) for x in _split_args(m.group(3) or "")]
                    emit(f"{cname}({', '.join(a)})")
                else:
                    emit(f"pass  # call {cname}")
                return
            if re.match(r"(write|print|flush|open|close|read|deallocate)\b", ln):
                emit("pass")
                return
            m = re.match(r"allocate\s*\((.*)\)$", ln)
            if m:
                for item in _split_args(m.group(1)):
                    k0 = item.index("(")
                    nm = item[:k0].strip()
                    dims = [self.expr(d, arrays, True) for d in _split_args(item[k0 + 1:_match_paren(item, k0)])]
                    emit(f"{nm} = farr(({', '.join(dims)},))")
                return
            m = re.match(r"do\s+while\s*\((.*)\)$", ln)
            if m:
                emit(f"while {self.expr(m.group(1), arrays)}:")
                ind += 1
                emit("pass")
                return
            m = re.match(r"do\s+(\w+)\s*=\s*(.*)$", ln)
            if m:
                var = m.group(1)
                parts = _split_args(m.group(2))
                lo, hi = self.expr(parts[0], arrays, True), self.expr(parts[1], arrays, True)
                if var in vec:
                    emit(f"for {var} in (ALL,):")
                elif len(parts) == 3:
                    st = self.expr(parts[2], arrays)
                    emit(f"for {var} in _frange({lo}, {hi}, {st}):")
                else:
                    emit(f"for {var} in range({lo}, ({hi})+1):")
                ind += 1
                emit("pass")
                return
            if re.match(r"end\s*do$", ln):
                ind -= 1
                return
            m = re.match(r"(else\s*if|elseif|if)\s*\(", ln)
            if m:
                k = ln.index("(", m.start())
                e = _match_paren(ln, k)
                cond = self.expr(ln[k + 1:e], arrays)
                rest = ln[e + 1:].strip()
                kw = "if" if m.group(1) == "if" else "elif"
                if rest == "then":
                    if kw == "elif":
                        ind -= 1
                    emit(f"{kw} {cond}:")
                    ind += 1
                    emit("pass")
                else:
                    emit(f"if {cond}:")
                    ind += 1
                    stmt(rest)
                    ind -= 1
                return
            if ln == "else":
                ind -= 1
                emit("else:")
                ind += 1
                emit("pass")
                return
            if re.match(r"end\s*if$", ln):
                ind -= 1
                return
            # assignment
            depth, pos = 0, -1
            for q, ch in enumerate(ln):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "=" and depth == 0 and ln[q - 1] not in "<>=/!" and ln[q + 1:q + 2] != "=":
                    pos = q
                    break
            if pos < 0:
                raise ValueError("cannot parse: " + ln)
            lhs, rhs = ln[:pos].strip(), ln[pos + 1:].strip()
            r = self.expr(rhs, arrays)
            if lhs in arrays and rhs in arrays:
                emit(f"{lhs}.a[...] = {rhs}.a")
            elif lhs in arrays:
                emit(f"{lhs}[...] = {r}")
            elif lhs in int_scalars and ("/" in rhs or "." in rhs):
                emit(f"{lhs} = _fint({r})")   # real -> integer assignment truncates (Fortran)
            else:
                emit(f"{self.expr(lhs, arrays)} = {r}")

        for nm, spec, is_cplx in local_arrays:
            dims = _split_args(spec)
            if any(d.strip() == ":" for d in dims):
                continue
            shp, lbs = [], []
            for d in dims:
                if ":" in d:
                    lo, hi = d.split(":", 1)
                    lo, hi = self.expr(lo, arrays, True), self.expr(hi, arrays, True)
                    shp.append(f"({hi})-({lo})+1")
                    lbs.append(lo)
                else:
                    shp.append(self.expr(d, arrays, True))
                    lbs.append("1")
            emit(f"{nm} = farr(({', '.join(shp)},), {'int' if is_cplx == 'int' else ('complex' if is_cplx else 'float')}, ({', '.join(lbs)},))")
        for ln in body:
            stmt(ln)
        emit("return locals()")
        return name, "\n".join(py)


def _frange(lo, hi, st):
    return range(lo, hi + (1 if st > 0 else -1), st)


def _fint(x):
    return x.astype(int) if isinstance(x, np.ndarray) else int(x)


def _real(x, *a):
    if isinstance(x, complex) or (isinstance(x, np.ndarray) and np.iscomplexobj(x)):
        return x.real
    return float(x) if np.isscalar(x) else x


def base_namespace():
    ns = {"ALL": ALL, "np": np, "farr": farr, "FArr": FArr, "_frange": _frange, "_real": _real, "_fint": _fint,
          "sin": np.sin, "cos": np.cos, "tan": np.tan, "exp": np.exp, "sqrt": np.sqrt,
          "abs": np.abs, "atan": np.arctan, "acos": np.arccos, "asin": np.arcsin,
          "log": np.log, "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh,
          "dacos": np.arccos, "dble": float, "int": int, "nint": lambda x: int(round(x)),
          "max": lambda *a: np.maximum.reduce(a) if any(isinstance(x, np.ndarray) for x in a) else max(a),
          "min": lambda *a: np.minimum.reduce(a) if any(isinstance(x, np.ndarray) for x in a) else min(a),
          "mod": lambda a, b: a % b,
          "present": lambda x: x is not None, "allocated": lambda x: True,
          "rl": lambda z: np.real(z), "iy": lambda z: np.imag(z), "aimag": lambda z: np.imag(z),
          "cx": lambda a, b: a + 1j * b if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else complex(a, b),
          "conjg": np.conj}
    return ns


def module_parameters(text):
    """named constants `real(mytype),parameter :: name=value` of a module"""
    ns = {}
    tr = Transpiler()
    env = base_namespace()
    for ln in _logical_lines(text):
        if "parameter" in ln and "::" in ln and re.match(r"(real|integer|complex)", ln):
            rest = ln.split("::", 1)[1]
            for item in _split_args(rest):
                if "=" not in item:
                    continue
                k, v = item.split("=", 1)
                try:
                    ns[k.strip()] = eval(tr.expr(v, set()), env, ns)
                except Exception:
                    pass
    return ns


def extract_subroutines(text):
    out = {}
    for m in re.finditer(r"^[ \t]*subroutine[ \t]+(\w+).*?^[ \t]*end[ \t]*subroutine[ \t]*\w*",
                         text, re.S | re.M | re.I):
        out[m.group(1).lower()] = m.group(0)
    return out
