#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ from the reference source.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It executes the reference's own statements (via f90mini.py) for
  * src/schemes.f90   prepare / first_derivative / second_derivative / interpolation
  * src/filters.f90   set_filter_coefficients + filx/fily/filz_{00,11,12,21,22}
  * src/derive.f90    all 42 operators
on small seeded random fields and writes
  tests/golden/schemes.npz, tests/golden/operators.npz
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import os
import re
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402

REF = os.environ.get("X3D_REFERENCE", "/root/reference")

NX, NY, NZ = 12, 11, 10
SEED = 20261017


def load_sources():
    src = {}
    for f in ("schemes", "derive", "filters", "module_param"):
        src[f] = open(os.path.join(REF, "src", f + ".f90")).read()
    return src


def dummy_args(text):
    lines = fm._logical_lines(text)
    m = re.match(r"subroutine\s+\w+\s*\((.*)\)", lines[0])
    return [a.strip() for a in m.group(1).split(",")]


def call_sites(text, sub):
    out = []
    for ln in fm._logical_lines(text):
        m = re.match(r"call\s+%s\s*\((.*)\)$" % sub, ln)
        if m:
            out.append([a.strip() for a in fm._split_args(m.group(1))])
    return out


class Ref:
    """Executes the reference's schemes()/filter() for one configuration."""

    def __init__(self, src, nx, ny, nz, ncl, ifirstder=4, isecondder=4, ipinter=3,
                 lengths=(2 * np.pi, 2 * np.pi, 2 * np.pi), nu0nu=4.0, cnu=0.44,
                 istret=0, af=0.45):
        self.src = src
        self.n = (nx, ny, nz)
        self.ncl = ncl  # ((nclx1,nclxn),(ncly1,nclyn),(nclz1,nclzn))
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(src["module_param"]))
        ns.update(dict(ifirstder=ifirstder, isecondder=isecondder, ipinter=ipinter,
                       nu0nu=nu0nu, cnu=cnu, nrank=1, iibm=0, istret=istret, iimplicit=0))
        self.ns = ns
        subs = fm.extract_subroutines(src["schemes"])
        tr = fm.Transpiler()
        for name in ("prepare", "first_derivative", "second_derivative", "interpolation"):
            _, code = tr.subroutine(subs[name])
            exec(code, ns)
        fsubs = fm.extract_subroutines(src["filters"])
        _, code = tr.subroutine(fsubs["set_filter_coefficients"])
        exec(code, ns)
        self.names = [set(), set(), set()]
        self.mod = {}  # module variables (derivX/Y/Z, parfiX/Y/Z, coefficient arrays)
        nm = []
        d = []
        for a in range(3):
            per = ncl[a] == (0, 0)
            nm.append(self.n[a] if per else self.n[a] - 1)
            d.append(lengths[a] / nm[a])
            self.mod["ncl" + "xyz"[a]] = per
        self.nm, self.d = nm, d
        sch = subs["schemes"]
        self._run_calls(sch, subs, "first_derivative",
                        lambda a: dict(d=d[a], n=self.n[a], ncl1=ncl[a][0], ncln=ncl[a][1]))
        self._run_calls(sch, subs, "second_derivative",
                        lambda a: dict(d2=d[a] * d[a], n=self.n[a], ncl1=ncl[a][0], ncln=ncl[a][1]))
        self._run_calls(sch, subs, "interpolation",
                        lambda a: dict(dx=d[a], nxm=nm[a], nx=self.n[a], nclx1=ncl[a][0], nclxn=ncl[a][1]),
                        stag=True)
        fil = fsubs["filter"]
        self._run_calls(fil, fsubs, "set_filter_coefficients",
                        lambda a: dict(af=af, n=self.n[a], ncl1=ncl[a][0], ncln=ncl[a][1]))
        ns.update(self.mod)

    def _run_calls(self, caller_text, subs, sub, inputs, stag=False):
        dums = dummy_args(subs[sub])
        sites = call_sites(caller_text, sub)[:3]  # velocity sets only (x, y, z)
        # array dummies and their extents
        decl = {}
        for ln in fm._logical_lines(subs[sub]):
            m = re.match(r"real\(mytype\)\s*,\s*dimension\((\w+)\).*::(.*)", ln)
            if m:
                for nm_ in fm._split_args(m.group(2)):
                    decl[nm_.strip()] = m.group(1)
        for a, actual in enumerate(sites):
            inp = inputs(a)
            args = []
            for dn, an in zip(dums, actual):
                if dn in inp:
                    args.append(inp[dn])
                elif dn in decl:
                    ext = inp[decl[dn]]
                    arr = fm.farr((ext,))
                    self.mod[an] = arr
                    args.append(arr)
                else:
                    args.append(None)
            out = self.ns[sub](*args)
            self.names[a].update(actual)
            for dn, an in zip(dums, actual):
                if dn not in inp and dn not in decl:
                    self.mod[an] = out[dn]

    def operator(self, name):
        if name in self.ns and callable(self.ns[name]):
            return self.ns[name]
        for f in ("derive", "filters"):
            subs = fm.extract_subroutines(self.src[f])
            if name in subs:
                axis = re.match(r"(?:inter|der|fil)([xyz])", name).group(1)
                vec = {"x": ("j", "k"), "y": ("i", "k"), "z": ("i", "j")}[axis]
                _, code = fm.Transpiler().subroutine(subs[name], vector_vars=vec)
                exec(code, self.ns)
                return self.ns[name]
        raise KeyError(name)


BCS = ["00", "11", "12", "21", "22"]


def ncl_for(axis, bc):
    ncl = [(0, 0), (0, 0), (0, 0)]
    ncl[axis] = (int(bc[0]), int(bc[1]))
    return tuple(ncl)


def main():
    src = load_sources()
    rng = np.random.default_rng(SEED)
    ops = {}
    sch = {}
    n = (NX, NY, NZ)
    u_full = rng.uniform(-1.0, 1.0, size=n)
    u_full = np.asfortranarray(u_full)
    ppy = rng.uniform(0.5, 1.5, size=NY)
    ops["u"] = u_full
    ops["ppy"] = ppy
    lengths = (2 * np.pi, 3.0, 1.7)
    for second in (4, 5):
        for axis in range(3):
            ax = "xyz"[axis]
            for bc in BCS:
                for istret in ((0, 2) if (axis == 1 and bc != "00") else (0,)):
                    ref = Ref(src, NX, NY, NZ, ncl_for(axis, bc), isecondder=second,
                              lengths=lengths, istret=istret)
                    m = ref.mod
                    tagc = f"{ax}{bc}_s{second}"
                    if istret == 0:
                        # coefficient arrays and scalars of the varied axis (schemes.npz)
                        for k in sorted(ref.names[axis]):
                            v = m.get(k)
                            if isinstance(v, fm.FArr):
                                sch[f"{tagc}/{k}"] = v.a.copy()
                            elif isinstance(v, (float, np.floating)):
                                sch[f"{tagc}/{k}"] = np.float64(v)
                    npaires = (1, 0) if bc != "00" and bc != "22" else (1,)
                    for npaire in npaires:
                        sfx = "p" if npaire == 1 else ""
                        fam = []
                        if second == 4:
                            fam.append((f"der{ax}_{bc}", "ff", "fs", "fw", True))
                            fam.append((f"fil{ax}_{bc}", "fiff", "fifs", "fifw", False))
                        fam.append((f"der{ax}{ax}_{bc}", "sf", "ss", "sw", False))
                        for name, f1, f2, f3, isd1 in fam:
                            if istret and not isd1:
                                continue
                            if name.startswith("der" + ax + ax):
                                # second derivative: the p-arrays go with npaire=1 (transeq.f90:442-444)
                                cf = [m[f"{f1}{ax}{sfx}"], m[f"{f2}{ax}{sfx}"], m[f"{f3}{ax}{sfx}"]]
                            else:
                                cf = [m[f"{f1}{ax}{sfx}"], m[f"{f2}{ax}{sfx}"], m[f"{f3}{ax}{sfx}"]]
                            t = fm.farr(n)
                            t.a[...] = -777.0
                            r = fm.farr(n)
                            s = fm.farr([n[d] for d in range(3) if d != axis])
                            u = fm.FArr(u_full.copy(order="F"))
                            fn = ref.operator(name)
                            if isd1 and axis == 1:
                                fn(t, u, r, s, cf[0], cf[1], cf[2], fm.FArr(ppy.copy()), NX, NY, NZ, npaire, 0.0)
                            else:
                                fn(t, u, r, s, cf[0], cf[1], cf[2], NX, NY, NZ, npaire, 0.0)
                            ops[f"{name}/np{npaire}/s{second}/st{istret}"] = t.a.copy()
    # staggered operators: periodic and non-periodic (closures shared by ncl=1 and 2)
    for axis in range(3):
        ax = "xyz"[axis]
        for bc in ("00", "11", "22", "12"):
            for istret in ((0, 2) if (axis == 1 and bc != "00") else (0,)):
                ref = Ref(src, NX, NY, NZ, ncl_for(axis, bc), lengths=lengths, istret=istret)
                m = ref.mod
                nn, nm_ = n[axis], ref.nm[axis]
                sh_v = list(n)
                sh_p = list(n)
                sh_p[axis] = nm_
                uv = np.asfortranarray(u_full.copy())
                up = np.asfortranarray(u_full[tuple(slice(0, s) for s in sh_p)].copy())
                six = {"x": "x6", "y": "y6", "z": "z6"}[ax]
                i6 = {"x": "i6", "y": "i6y", "z": "i6z"}[ax]
                ppyi = rng.uniform(0.5, 1.5, size=nm_)
                ops[f"ppyi/{bc}"] = ppyi if axis == 1 else ops.get(f"ppyi/{bc}", ppyi)
                for npaire in (1, 0):
                    pf = "p" if npaire == 1 else ""
                    other = [s for d, s in enumerate(n) if d != axis]
                    cases = []
                    # (name, input, out shape, args builder)
                    cases.append((f"der{ax}vp", uv, sh_p, [m[f"cf{six}"], m[f"cs{six}"], m[f"cw{six}"]]))
                    cases.append((f"inter{ax}vp", uv, sh_p,
                                  [m[f"cif{ax}p6"], m[f"cis{ax}p6"], m[f"ciw{ax}p6"]]))
                    cases.append((f"der{ax}pv", up, sh_v,
                                  [m[f"cfip6{i6[2:]}"], m[f"csip6{i6[2:]}"], m[f"cwip6{i6[2:]}"],
                                   m[f"cf{six}"], m[f"cs{six}"], m[f"cw{six}"]]))
                    cases.append((f"inter{ax}pv", up, sh_v,
                                  [m[f"cifip6{i6[2:]}"], m[f"cisip6{i6[2:]}"], m[f"ciwip6{i6[2:]}"],
                                   m[f"cif{six}"], m[f"cis{six}"], m[f"ciw{six}"]]))
                    for name, uin, sho, cf in cases:
                        if istret and name not in ("deryvp", "derypv"):
                            continue
                        t = fm.farr(sho)
                        t.a[...] = -777.0
                        r = fm.farr(sh_v)
                        s = fm.farr(other)
                        fn = ref.operator(name)
                        u = fm.FArr(uin.copy(order="F"))
                        if ax == "x":
                            dims = ([NX, nm_, NY, NZ] if name.endswith("vp") else [nm_, NX, NY, NZ])
                            fn(t, u, r, s, *cf, *dims, npaire)
                        elif ax == "y":
                            if name == "interyvp":
                                fn(t, u, r, s, *cf, NX, NY, nm_, NZ, npaire)
                            elif name == "deryvp":
                                fn(t, u, r, s, *cf, fm.FArr(ppyi.copy()), NX, NY, nm_, NZ, npaire)
                            elif name == "interypv":
                                fn(t, u, r, s, *cf, NX, nm_, NY, NZ, npaire)
                            else:
                                fn(t, u, r, s, *cf, fm.FArr(ppy.copy()), NX, nm_, NY, NZ, npaire)
                        else:
                            dims = ([NX, NY, NZ, nm_] if name.endswith("vp") else [NX, NY, nm_, NZ])
                            fn(t, u, r, s, *cf, *dims, npaire)
                        ops[f"{name}/bc{bc}/np{npaire}/st{istret}"] = t.a.copy()
    # scheme option sweep (scalars only) for x
    for fd, sd, ip in ((1, 1, 1), (4, 4, 1), (4, 4, 2), (4, 5, 3)):
        ref = Ref(src, NX, NY, NZ, ((2, 1), (0, 0), (1, 2)), ifirstder=fd, isecondder=sd, ipinter=ip,
                  lengths=lengths)
        for a in range(3):
            for k in sorted(ref.names[a]):
                v = ref.mod.get(k)
                key = f"opt_{fd}{sd}{ip}/{k}"
                if isinstance(v, fm.FArr):
                    sch[key] = v.a.copy()
                elif isinstance(v, (float, np.floating)):
                    sch[key] = np.float64(v)
    meta = dict(nx=NX, ny=NY, nz=NZ, seed=SEED, lengths=np.array(lengths), af=0.45,
                nu0nu=4.0, cnu=0.44)
    np.savez_compressed(os.path.join(HERE, "operators.npz"), **ops, **{"meta/" + k: v for k, v in meta.items()})
    # pack the scalars of each configuration into one (names, values) pair
    packed, scal = {}, {}
    for k, v in sch.items():
        tag, nm_ = k.split("/", 1)
        if np.ndim(v) == 0:
            scal.setdefault(tag, []).append((nm_, float(v)))
        else:
            packed[k] = v
    for tag, items in scal.items():
        packed[tag + "/scalar_names"] = np.array([a for a, _ in items])
        packed[tag + "/scalar_values"] = np.array([b for _, b in items])
    sch = packed
    np.savez_compressed(os.path.join(HERE, "schemes.npz"), **sch)
    print("operators:", len(ops), "entries; schemes:", len(sch), "entries")


if __name__ == "__main__":
    main()
