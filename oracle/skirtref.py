"""ctypes front end for oracle/_ref/libskirtref.so -- TEST INFRASTRUCTURE ONLY.

The library holds the reference's OWN translation units (compiled in place from /root/reference by
oracle/Makefile) behind the small C harness oracle/ref_harness.cpp.  Only tests/, the golden-vector
generator, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libskirtref.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.skr_error.restype = C.c_char_p
        L.skr_create.restype = C.c_void_p
        L.skr_create.argtypes = [C.c_char_p]
        L.skr_packages_per_lambda.restype = C.c_double
        L.skr_path_batch.restype = C.c_long
        L.skr_warnings.restype = C.c_long
        L.skr_saved_image.restype = C.c_long
        L.skr_saved_table.restype = C.c_long
        _lib = L
    return _lib


class RefError(RuntimeError):
    pass


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefSim:
    """One reference simulation hierarchy (MonteCarloSimulation + children) built from a spec."""

    def __init__(self, spec, luminosities=None, mixes=None, particles=None, amesh=None, mueller=None):
        L = lib()
        self.h = C.c_void_p(L.skr_create(spec.encode()))
        if not self.h:
            raise RefError(L.skr_error().decode())
        self.spec = spec
        for i, lum in enumerate(luminosities or []):
            lum = _f64(lum)
            self._chk(L.skr_set_luminosities(self.h, i, lum.ctypes.data_as(C.c_void_p), len(lum)))
        for i, (kabs, ksca, g) in enumerate(mixes or []):
            kabs, ksca, g = _f64(kabs), _f64(ksca), _f64(g)
            self._chk(L.skr_set_mix(self.h, i, kabs.ctypes.data_as(C.c_void_p), ksca.ctypes.data_as(C.c_void_p),
                                    g.ctypes.data_as(C.c_void_p), len(g)))
        for i, mu in enumerate(mueller or []):
            if mu is None:
                continue
            S11, S12, S33, S34 = (_f64(v) for v in mu)       # each [Nlambda, Ntheta]
            self._chk(L.skr_set_mueller(self.h, i, S11.shape[1], S11.shape[0], S11.ctypes.data_as(C.c_void_p), S12.ctypes.data_as(C.c_void_p),
                                        S33.ctypes.data_as(C.c_void_p), S34.ctypes.data_as(C.c_void_p)))
        if particles is not None:
            p = _f64(particles)
            self._chk(L.skr_set_particles(self.h, p.ctypes.data_as(C.c_void_p), len(p)))
        if amesh is not None:
            nxyz = np.ascontiguousarray(amesh[0], dtype=np.int32)
            val = _f64(amesh[1])
            self._chk(L.skr_set_amesh(self.h, nxyz.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), len(val)))
        self._setup = False

    def _chk(self, rc):
        if rc:
            raise RefError(lib().skr_error().decode())

    def close(self):
        if self.h:
            lib().skr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self):
        self._chk(lib().skr_setup(self.h))
        self._setup = True
        return self

    # ---- scalars -------------------------------------------------------------------------------
    @property
    def Nlambda(self):
        return lib().skr_num_lambda(self.h)

    @property
    def Ncells(self):
        return lib().skr_num_cells(self.h)

    @property
    def Ncomp(self):
        return lib().skr_num_comp(self.h)

    @property
    def Nstellar(self):
        return lib().skr_num_stellar(self.h)

    @property
    def grid_kind(self):
        return lib().skr_grid_kind(self.h)

    def packages_per_lambda(self):
        return lib().skr_packages_per_lambda(self.h)

    # ---- tables --------------------------------------------------------------------------------
    def wavelengths(self):
        lam = np.zeros(self.Nlambda); dlam = np.zeros(self.Nlambda)
        lib().skr_get_lambda(self.h, lam.ctypes.data_as(C.c_void_p), dlam.ctypes.data_as(C.c_void_p))
        return lam, dlam

    def interstellar_mix(self):
        n = self.Nlambda
        a, s, g = np.zeros(n), np.zeros(n), np.zeros(n)
        self._chk(lib().skr_interstellar_mix(self.h, a.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                             g.ctypes.data_as(C.c_void_p)))
        return a, s, g

    def medium(self):
        N, Cn, Ln = self.Ncells, self.Ncomp, self.Nlambda
        rho = np.zeros((N, Cn)); kext = np.zeros((Cn, Ln)); ksca = np.zeros((Cn, Ln)); g = np.zeros((Cn, Ln))
        lib().skr_get_rho(self.h, rho.ctypes.data_as(C.c_void_p))
        lib().skr_get_opt(self.h, kext.ctypes.data_as(C.c_void_p), ksca.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
        return dict(rho=rho, kext=kext, ksca=ksca, g=g)

    def albedo(self):
        a = np.zeros((self.Ncomp, self.Nlambda))
        lib().skr_get_albedo(self.h, a.ctypes.data_as(C.c_void_p))
        return a

    def volumes(self):
        v = np.zeros(self.Ncells)
        lib().skr_get_volumes(self.h, v.ctypes.data_as(C.c_void_p))
        return v

    def luminosities(self):
        Lm = np.zeros((self.Nstellar, self.Nlambda))
        lib().skr_get_luminosities(self.h, Lm.ctypes.data_as(C.c_void_p))
        return Lm

    def grid_tables(self):
        """Flattened grid in the layout of include/skirtgpu.h (dict with 'kind')."""
        L = lib(); k = self.grid_kind
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        if k == 0:
            n = np.zeros(3, dtype=np.int32); L.skr_cart_dims(self.h, vp(n))
            xv, yv, zv = np.zeros(n[0]+1), np.zeros(n[1]+1), np.zeros(n[2]+1)
            L.skr_cart_axes(self.h, vp(xv), vp(yv), vp(zv))
            return dict(kind="cartesian", xv=xv, yv=yv, zv=zv)
        if k in (1, 2, 5, 6):           # 5, 6: ParticleTreeDustGrid (octree / binary tree) -- same node tables, its own traversal
            nn = C.c_int(); nb = C.c_int(); eps = C.c_double()
            L.skr_tree_sizes(self.h, C.byref(nn), C.byref(nb), C.byref(eps))
            N = nn.value
            box = np.zeros((N, 6)); child0 = np.zeros(N, np.int32); parent = np.zeros(N, np.int32)
            cell = np.zeros(N, np.int32); sdir = np.zeros(N, np.int32)
            nbrStart = np.zeros(6*N+1, np.int32); nbrIds = np.zeros(max(nb.value, 1), np.int32)
            self._chk(L.skr_tree_tables(self.h, vp(box), vp(child0), vp(parent), vp(cell), vp(sdir), vp(nbrStart), vp(nbrIds)))
            search = 3 if k >= 5 else int(self.spec_value("grid")[3])
            return dict(kind="octtree" if k in (1, 5) else "bintree", search=search, eps=eps.value, box=box, child0=child0,
                        parent=parent, cell=cell, dir=sdir, nbrStart=nbrStart, nbrIds=nbrIds[:nb.value])
        if k == 4:
            nn = C.c_int(); eps = C.c_double()
            L.skr_amesh_sizes(self.h, C.byref(nn), C.byref(eps))
            N = nn.value
            box = np.zeros((N, 6)); nxyz = np.zeros((N, 3), np.int32); child0 = np.zeros(N, np.int32)
            cell = np.zeros(N, np.int32); wall = np.zeros((N, 6), np.int32)
            L.skr_amesh_tables(self.h, vp(box), vp(nxyz), vp(child0), vp(cell), vp(wall))
            return dict(kind="amesh", eps=eps.value, box=box, nxyz=nxyz, child0=child0, cell=cell, wallNbr=wall)
        if k == 3:
            sizes = np.zeros(5, np.int64); eps = C.c_double()
            L.skr_voro_sizes(self.h, vp(sizes), C.byref(eps))
            N, nn, nb, nr, nkd = [int(v) for v in sizes]
            t = dict(kind="voronoi", eps=eps.value, nb=nb, extent=np.array(self.spec_value("box"), dtype=np.float64),
                     particles=np.zeros((N, 3)), cellBox=np.zeros((N, 6)), nbrStart=np.zeros(N+1, np.int32),
                     nbrIds=np.zeros(max(nn, 1), np.int32), blkStart=np.zeros(nb**3+1, np.int32),
                     blkIds=np.zeros(max(nr, 1), np.int32), blkTree=np.zeros(nb**3, np.int32),
                     kdM=np.zeros(max(nkd, 1), np.int32), kdAxis=np.zeros(max(nkd, 1), np.int32),
                     kdUp=np.zeros(max(nkd, 1), np.int32), kdLeft=np.zeros(max(nkd, 1), np.int32),
                     kdRight=np.zeros(max(nkd, 1), np.int32))
            L.skr_voro_tables(self.h, vp(t["particles"]), vp(t["cellBox"]), vp(t["nbrStart"]), vp(t["nbrIds"]),
                              vp(t["blkStart"]), vp(t["blkIds"]), vp(t["blkTree"]), vp(t["kdM"]), vp(t["kdAxis"]),
                              vp(t["kdUp"]), vp(t["kdLeft"]), vp(t["kdRight"]))
            for key in ("kdM", "kdAxis", "kdUp", "kdLeft", "kdRight"):
                t[key] = t[key][:nkd]
            t["nbrIds"] = t["nbrIds"][:nn]; t["blkIds"] = t["blkIds"][:nr]
            return t
        if k in (7, 8, 9):
            sizes = np.zeros(2, np.int32); L.skr_sym_sizes(self.h, vp(sizes))
            v1 = np.zeros(sizes[0] + 1); v2 = np.zeros(sizes[1] + 1); cv = np.zeros(sizes[1] + 1)
            L.skr_sym_tables(self.h, vp(v1), vp(v2), vp(cv))
            if k == 7:
                return dict(kind="sphere1d", rv=v1)
            if k == 8:
                return dict(kind="sphere2d", rv=v1, thetav=v2, cv=cv)
            return dict(kind="cylinder2d", Rv=v1, zv=v2)
        raise RefError("no grid")

    def spec_value(self, key):
        for line in self.spec.splitlines():
            w = line.split()
            if w and w[0] == key:
                return w[1:]
        return None

    # ---- deterministic geometry ------------------------------------------------------------------
    def path_batch(self, r, k, ell=-1, nthreads=1):
        """DustGrid::path (+fillOpticalDepth when ell>=0) for fixed rays -> CSR dict."""
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        off = np.zeros(n+1, np.int64)
        total = lib().skr_path_batch(self.h, vp(r), vp(k), C.c_long(n), ell, C.c_long(0), vp(off), None, None, None, None, None, nthreads)
        if total < 0:
            raise RefError(lib().skr_error().decode())
        m = np.zeros(total, np.int32); ds = np.zeros(total); s = np.zeros(total); dtau = np.zeros(total); tau = np.zeros(total)
        total = lib().skr_path_batch(self.h, vp(r), vp(k), C.c_long(n), ell, C.c_long(max(total, 1)), vp(off), vp(m), vp(ds), vp(s),
                                     vp(dtau), vp(tau), nthreads)
        if total < 0:
            raise RefError(lib().skr_error().decode())
        return dict(offsets=off, m=m, ds=ds, s=s, dtau=dtau, tau=tau)

    def whichcell(self, r):
        r = _f64(r).reshape(-1, 3); m = np.zeros(len(r), np.int32)
        self._chk(lib().skr_whichcell(self.h, r.ctypes.data_as(C.c_void_p), C.c_long(len(r)), m.ctypes.data_as(C.c_void_p)))
        return m

    def opticaldepth_batch(self, r, k, ell, distance=None):
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); tau = np.zeros(len(r))
        d = None if distance is None else _f64(distance)
        self._chk(lib().skr_opticaldepth_batch(self.h, r.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), C.c_long(len(r)),
                                               ell, None if d is None else d.ctypes.data_as(C.c_void_p), tau.ctypes.data_as(C.c_void_p)))
        return tau

    # ---- Monte Carlo -----------------------------------------------------------------------------
    def reset(self, seed):
        self._chk(lib().skr_reset(self.h, int(seed)))

    def set_packages(self, n):
        lib().skr_set_packages(self.h, C.c_double(n))

    def run_stellar(self):
        sec = C.c_double()
        self._chk(lib().skr_run_stellar(self.h, C.byref(sec)))
        return sec.value

    def instruments(self):
        """list of dicts with raw (uncalibrated) detector arrays, as accumulated by detect()"""
        out = []
        for i in range(lib().skr_num_instruments(self.h)):
            nf = C.c_long(); ns = C.c_long()
            lib().skr_instrument_sizes(self.h, i, C.byref(nf), C.byref(ns))
            frame = np.zeros(nf.value); sed = np.zeros(ns.value)
            lib().skr_get_instrument(self.h, i, frame.ctypes.data_as(C.c_void_p) if nf.value else None,
                                     sed.ctypes.data_as(C.c_void_p) if ns.value else None)
            geo = np.zeros(16)
            lib().skr_get_instrument_geometry(self.h, i, geo.ctypes.data_as(C.c_void_p))
            out.append(dict(frame=frame, sed=sed, geometry=geo))
        return out

    def multiframe(self, i, which, ell):
        """raw array of one frame of a MultiFrameInstrument (which = -1: total flux, k >= 0: stellar component k)"""
        L = lib(); L.skr_get_multiframe.restype = C.c_long
        n = L.skr_get_multiframe(self.h, int(i), int(which), int(ell), None)
        if n < 0:
            raise RefError("no such multi-frame array")
        a = np.zeros(n)
        L.skr_get_multiframe(self.h, int(i), int(which), int(ell), a.ctypes.data_as(C.c_void_p))
        return a

    def full_channel(self, i, c, nframe):
        """raw (uncalibrated) data cube and SED of one FullInstrument channel"""
        frame = np.zeros(nframe); sed = np.zeros(self.Nlambda)
        if lib().skr_get_full_channel(self.h, int(i), int(c), frame.ctypes.data_as(C.c_void_p), sed.ctypes.data_as(C.c_void_p)):
            raise RefError("not a FullInstrument channel")
        return frame, sed

    def labs(self):
        a = np.zeros((self.Ncells, self.Nlambda))
        self._chk(lib().skr_get_labs(self.h, a.ctypes.data_as(C.c_void_p)))
        return a

    # ---- dust emission phases -------------------------------------------------------------------
    def prepare_dust(self, ynstellar=True):
        """calculatedustemission + bolometric absorbed luminosities; returns Lv[Nlambda, Ncells]"""
        Lv = np.zeros((self.Nlambda, self.Ncells))
        self._chk(lib().skr_prepare_dust(self.h, int(bool(ynstellar)), Lv.ctypes.data_as(C.c_void_p)))
        return Lv

    def run_dust(self, selfabs, factor=1.0):
        sec = C.c_double()
        self._chk(lib().skr_run_dust(self.h, int(bool(selfabs)), C.c_double(factor), C.byref(sec)))
        return sec.value

    def labs_dust(self):
        a = np.zeros((self.Ncells, self.Nlambda))
        self._chk(lib().skr_get_labs_dust(self.h, a.ctypes.data_as(C.c_void_p)))
        return a

    def labs_bol(self):
        a = np.zeros(self.Ncells)
        self._chk(lib().skr_get_labs_bol(self.h, a.ctypes.data_as(C.c_void_p)))
        return a

    def random_positions(self, m, n):
        xyz = np.zeros((n, 3))
        self._chk(lib().skr_random_positions(self.h, int(m), C.c_long(n), xyz.ctypes.data_as(C.c_void_p)))
        return xyz

    # ---- output ------------------------------------------------------------------------------------
    def write_instruments(self):
        """Instrument::write() for every instrument (calibration included); note that it calibrates the detector
        arrays IN PLACE, like the reference does at the end of a simulation"""
        self._chk(lib().skr_write_instruments(self.h))

    def saved_image(self, name):
        n = lib().skr_saved_image(name.encode(), None, C.c_long(0))
        if n < 0:
            raise RefError(f"no image named {name}")
        a = np.zeros(n)
        lib().skr_saved_image(name.encode(), a.ctypes.data_as(C.c_void_p), C.c_long(n))
        return a

    def saved_table(self, name):
        nc = C.c_int()
        n = lib().skr_saved_table(name.encode(), None, C.c_long(0), C.byref(nc))
        if n < 0:
            raise RefError(f"no table named {name}")
        a = np.zeros(n)
        lib().skr_saved_table(name.encode(), a.ctypes.data_as(C.c_void_p), C.c_long(n), C.byref(nc))
        return a.reshape(-1, max(nc.value, 1))

    def sample_launch(self, ell, n):
        r = np.zeros((n, 3)); k = np.zeros((n, 3)); Lw = np.zeros(n)
        self._chk(lib().skr_sample_launch(self.h, ell, C.c_long(n), r.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                          Lw.ctypes.data_as(C.c_void_p)))
        return r, k, Lw

    def uniforms(self, n):
        u = np.zeros(n)
        lib().skr_uniforms(self.h, C.c_long(n), u.ctypes.data_as(C.c_void_p))
        return u
