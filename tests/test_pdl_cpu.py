"""Programmatic dependent launch invariant (msx_common.cuh): a kernel launched through msx_launch() may start before its
stream predecessor has finished, so its FIRST statement must be pdl_entry() (griddepcontrol.wait).  Checked on the
sources: every kernel name passed to msx_launch() is a __global__ function whose body begins with pdl_entry()."""
import glob
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "musicstyletransfer_b200", "csrc")


def _kernels(src):
    """name -> True when the body of the __global__ function starts with pdl_entry();"""
    out, pos = {}, 0
    while True:
        g = src.find("__global__", pos)
        if g < 0:
            return out
        p = src.find("(", g)
        while True:                      # the parameter list is the last (...) group before '{' or ';'
            depth, m = 0, p
            while True:
                if src[m] == "(":
                    depth += 1
                elif src[m] == ")":
                    depth -= 1
                    if depth == 0:
                        break
                m += 1
            nxt = m + 1
            while src[nxt] in " \n\t\\":
                nxt += 1
            if src[nxt] in "{;":
                break
            p = src.find("(", m)
        name = re.findall(r"([A-Za-z_0-9]+)\s*$", src[g:p])[0]
        if src[nxt] == "{":
            out[name] = out.get(name, True) and src[nxt + 1:].lstrip(" \n\\").startswith("pdl_entry();")
        pos = nxt + 1


def test_every_pdl_launched_kernel_waits_first():
    launched_total = 0
    for path in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        src = open(path).read()
        launched = set(re.findall(r"msx_launch\(\s*([A-Za-z_0-9]+)", src))
        kernels = _kernels(src)
        for k in launched:
            assert k in kernels, (os.path.basename(path), k, "launched through msx_launch but not defined in this file")
            assert kernels[k], (os.path.basename(path), k, "must call pdl_entry() first")
        launched_total += len(launched)
    assert launched_total >= 30          # the train step's kernels all take the programmatic launch


def test_pdl_switch_is_exported():
    import ctypes
    from musicstyletransfer_b200 import lib
    l = lib.load()
    prev = l.msx_get_pdl()
    l.msx_set_pdl(ctypes.c_int(0))
    assert l.msx_get_pdl() == 0
    l.msx_set_pdl(ctypes.c_int(1))
    assert l.msx_get_pdl() == 1
    l.msx_set_pdl(ctypes.c_int(prev))
