"""ctypes front end of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

liboracle.so is the CPU restatement (oracle/skirt_oracle.cpp) of the reference's DustGrid::path() family,
DustGridPath and optical-depth integration on the flattened tables of include/skirtgpu.h.  It is pinned
against the golden vectors under tests/golden/ (generated from the reference's own code) and, where
oracle/_ref exists, against that library directly (tests/test_oracle.py).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        for name in ("orc_grid_cartesian", "orc_grid_tree", "orc_grid_amesh", "orc_grid_voronoi", "orc_grid_sphere1d", "orc_grid_sphere2d", "orc_grid_cylinder2d",
                     "orc_medium"):
            getattr(L, name).restype = C.c_void_p
        L.orc_path_batch.restype = C.c_long
        L.orc_stuck.restype = C.c_long
        L.orc_pathlength.restype = C.c_double
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """grid + medium in the table layout of include/skirtgpu.h (the dicts the golden fixtures hold)"""

    def __init__(self, tables, medium=None):
        L = lib(); t = tables; kind = t["kind"]
        self._keep = []
        if kind == "cartesian":
            xv, yv, zv = _f64(t["xv"]), _f64(t["yv"]), _f64(t["zv"])
            self.g = L.orc_grid_cartesian(_p(xv), len(xv) - 1, _p(yv), len(yv) - 1, _p(zv), len(zv) - 1)
        elif kind in ("octtree", "bintree"):
            box = _f64(t["box"]); child0 = _i32(t["child0"]); parent = _i32(t["parent"]); cell = _i32(t["cell"])
            d = _i32(t["dir"]) if t.get("dir") is not None else None
            ns = _i32(t["nbrStart"]) if t.get("nbrStart") is not None else None
            ni = _i32(t["nbrIds"] if len(t["nbrIds"]) else [0]) if t.get("nbrIds") is not None else None
            self.g = L.orc_grid_tree(0 if kind == "octtree" else 1, int(t["search"]), len(child0), _p(box), _p(child0), _p(parent),
                                     _p(cell), _p(d), _p(ns), _p(ni))
        elif kind == "amesh":
            box = _f64(t["box"]); nxyz = _i32(t["nxyz"]); child0 = _i32(t["child0"]); cell = _i32(t["cell"]); wall = _i32(t["wallNbr"])
            self.g = L.orc_grid_amesh(len(child0), _p(box), _p(nxyz), _p(child0), _p(cell), _p(wall))
        elif kind == "voronoi":
            pad = lambda v: _i32(v if len(v) else [0])
            part = _f64(t["particles"]); ext = _f64(t["extent"]); cb = _f64(t["cellBox"]) if t.get("cellBox") is not None else None
            self.g = L.orc_grid_voronoi(len(part), _p(part), _p(_i32(t["nbrStart"])), _p(pad(t["nbrIds"])), _p(ext), int(t["nb"]),
                                        _p(_i32(t["blkStart"])), _p(pad(t["blkIds"])), _p(_i32(t["blkTree"])), len(t["kdM"]),
                                        _p(pad(t["kdM"])), _p(pad(t["kdAxis"])), _p(pad(t["kdUp"])), _p(pad(t["kdLeft"])),
                                        _p(pad(t["kdRight"])), _p(cb))
        elif kind == "sphere1d":
            rv = _f64(t["rv"]); self.g = L.orc_grid_sphere1d(len(rv) - 1, _p(rv))
        elif kind == "sphere2d":
            rv, tv, cv = _f64(t["rv"]), _f64(t["thetav"]), _f64(t["cv"])
            self.g = L.orc_grid_sphere2d(len(rv) - 1, _p(rv), len(tv) - 1, _p(tv), _p(cv))
        elif kind == "cylinder2d":
            Rv, zv = _f64(t["Rv"]), _f64(t["zv"]); self.g = L.orc_grid_cylinder2d(len(Rv) - 1, _p(Rv), len(zv) - 1, _p(zv))
        else:
            raise ValueError(kind)
        self.g = C.c_void_p(self.g)
        self.m = None
        if medium is not None:
            rho = _f64(medium["rho"]); rho = rho[:, None] if rho.ndim == 1 else rho
            kext = _f64(np.atleast_2d(medium["kext"]))
            ksca = _f64(np.atleast_2d(medium["ksca"])) if medium.get("ksca") is not None else None
            g = _f64(np.atleast_2d(medium["g"])) if medium.get("g") is not None else None
            self.m = C.c_void_p(L.orc_medium(rho.shape[0], rho.shape[1], kext.shape[1], _p(rho), _p(kext), _p(ksca), _p(g)))
            self.Nlambda = kext.shape[1]

    def close(self):
        if getattr(self, "g", None):
            lib().orc_grid_destroy(self.g); self.g = None
        if getattr(self, "m", None):
            lib().orc_medium_destroy(self.m); self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def Ncells(self):
        return lib().orc_num_cells(self.g)

    def stuck(self):
        return lib().orc_stuck(self.g)

    def path_batch(self, r, k, ell=None, nthreads=None):
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        nthreads = nthreads or min(os.cpu_count() or 1, 16)
        ella, stride = None, 0
        if ell is not None:
            ella = _i32(np.atleast_1d(ell)); stride = 1 if len(ella) == n and n > 1 else 0
        off = np.zeros(n + 1, np.int64)
        args = (self.g, self.m, _p(r), _p(k), C.c_long(n), _p(ella), stride)
        total = lib().orc_path_batch(*args, C.c_long(0), _p(off), None, None, None, None, None, nthreads)
        t = max(total, 1)
        m = np.zeros(t, np.int32); ds = np.zeros(t); s = np.zeros(t); dtau = np.zeros(t); tau = np.zeros(t)
        lib().orc_path_batch(*args, C.c_long(t), _p(off), _p(m), _p(ds), _p(s), _p(dtau), _p(tau), nthreads)
        return dict(offsets=off, m=m[:total], ds=ds[:total], s=s[:total], dtau=dtau[:total], tau=tau[:total])

    def whichcell(self, r):
        r = _f64(r).reshape(-1, 3); m = np.zeros(len(r), np.int32)
        lib().orc_whichcell(self.g, _p(r), C.c_long(len(r)), _p(m))
        return m

    def opticaldepth(self, r, k, ell, distance=None):
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        ella = _i32(np.atleast_1d(ell)); stride = 1 if len(ella) == n and n > 1 else 0
        d = None if distance is None else _f64(distance)
        tau = np.zeros(n)
        lib().orc_opticaldepth(self.g, self.m, _p(r), _p(k), C.c_long(n), _p(ella), stride, _p(d), _p(tau))
        return tau

    def pathlength(self, r, k, ell, tau):
        r = _f64(r); k = _f64(k)
        return lib().orc_pathlength(self.g, self.m, _p(r), _p(k), int(ell), C.c_double(tau))

    def random_positions(self, m, seed, n):
        xyz = np.zeros((n, 3))
        lib().orc_random_positions(self.g, int(m), C.c_ulong(seed), C.c_long(n), _p(xyz))
        return xyz


def uniforms(seed, n):
    u = np.zeros(n)
    lib().orc_uniforms(C.c_ulong(seed), C.c_long(n), _p(u))
    return u
