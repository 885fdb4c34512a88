"""Mapping between the reference's per-axis module names (src/module_param.f90:559-656,
src/variables.f90) and the axis-free names used by the oracle / C ABI structs."""
import re


def canon_array(name, ax):
    """ffx -> ff, sfyp -> sfp, cfy6 -> cfx6, cfi6z -> cfi6, fiffzp -> fiffp"""
    n = name
    if re.search(r"i6[yz]?$|ip6[yz]?$", n):          # velocity-sized staggered arrays: cfi6, cfip6y, ...
        return re.sub(r"[yz]$", "", n)
    m = re.match(r"^(c\w*?)%s(p?6)$" % ax, n)          # pressure-sized: cfx6, cifyp6, ...
    if m:
        return m.group(1) + "x" + m.group(2)
    m = re.match(r"^(\w+?)%s(p?)$" % ax, n)            # ffx, ffxp, sfy, fiffzp
    if m:
        return m.group(1) + m.group(2)
    return n


def canon_scalar(name, ax):
    n = name
    m = re.match(r"^(\w+?)%s(6?)$" % ax, n)
    if not m:
        return n
    base, six = m.group(1), m.group(2)
    base = re.sub(r"^(alfa|af|bf|alsa|as|bs|cs|ds|fial|fia|fib|fic|fid|alca|ac|bc|ailca|aic|bic|cic|dic)[ijk]$",
                  r"\1i", base)
    out = base + six
    return "asm_" if out == "asm" else out
