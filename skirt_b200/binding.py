"""ctypes binding of include/skirtgpu.h (the drop-in C ABI).  Host arrays are numpy; device arrays are
passed as raw pointers (e.g. torch.Tensor.data_ptr()).  Every non-zero status becomes EngineError, the
Python analogue of the reference adapter's `throw FATALERROR(msg)` (FatalError.hpp:47)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SKG_LIBRARY selects another build of the same library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get("SKG_LIBRARY") or os.path.join(_HERE, "libskirtgpu.so")
_lib = None

SKG_HOST, SKG_DEVICE = 0, 1
GEOM_EXPDISK, GEOM_SERSIC = 1, 2
INSTR_FRAME, INSTR_SED, INSTR_SIMPLE, INSTR_FULL, INSTR_MULTIFRAME, INSTR_PERSPECTIVE = 1, 2, 3, 4, 5, 6
# FullInstrument channels (include/skirtgpu.h SKG_CHAN_*): scattering level n is channel CHAN_LEVEL1 + n - 1
CHAN_TRANSPARENT, CHAN_STELLAR_DIRECT, CHAN_STELLAR_SCATTERED, CHAN_DUST_DIRECT, CHAN_DUST_SCATTERED, CHAN_LEVEL1 = 0, 1, 2, 3, 4, 5
PHASE_STELLAR, PHASE_DUST_SELFABS, PHASE_DUST_EMISSION = 0, 1, 2
REDUCE_LABS_STELLAR, REDUCE_LABS_DUST, REDUCE_INSTRUMENTS, REDUCE_ALL = 1, 2, 4, 7
# skg_segment == DustGridPath::Segment (DustGridPath.hpp:161-167)
SEGMENT = np.dtype([("m", np.int32), ("reserved", np.int32), ("ds", np.float64), ("s", np.float64), ("dtau", np.float64), ("tau", np.float64)])


class EngineError(RuntimeError):
    pass


class SkgSource(C.Structure):
    _fields_ = [("geometry", C.c_int), ("p", C.c_double * 8),
                ("spiral_arms", C.c_int), ("spiral_index", C.c_int),
                ("spiral_pitch", C.c_double), ("spiral_radius", C.c_double),
                ("spiral_phase", C.c_double), ("spiral_weight", C.c_double),
                ("ntab", C.c_int), ("rv", C.c_void_p), ("Xv", C.c_void_p), ("Sv", C.c_void_p)]


class SkgInstrumentFrame(C.Structure):
    _fields_ = [("Nxp", C.c_int), ("Nyp", C.c_int), ("fovxp", C.c_double), ("fovyp", C.c_double), ("xpc", C.c_double), ("ypc", C.c_double)]


class SkgInstrument(C.Structure):
    _fields_ = [("kind", C.c_int), ("distance", C.c_double), ("inclination", C.c_double),
                ("azimuth", C.c_double), ("positionAngle", C.c_double),
                ("Nxp", C.c_int), ("Nyp", C.c_int),
                ("fovxp", C.c_double), ("fovyp", C.c_double), ("xpc", C.c_double), ("ypc", C.c_double),
                ("scatteringLevels", C.c_int), ("writeTotal", C.c_int), ("writeStellarComps", C.c_int),
                ("frames", C.POINTER(SkgInstrumentFrame)),
                ("viewX", C.c_double), ("viewY", C.c_double), ("viewZ", C.c_double), ("crossX", C.c_double), ("crossY", C.c_double),
                ("crossZ", C.c_double), ("upX", C.c_double), ("upY", C.c_double), ("upZ", C.c_double), ("focal", C.c_double)]


class SkgMcParams(C.Structure):
    _fields_ = [("packages", C.c_double), ("luminosityScale", C.c_double), ("minWeightReduction", C.c_double),
                ("minScattEvents", C.c_double), ("scattBias", C.c_double), ("storeAbsorption", C.c_int),
                ("seed", C.c_uint64), ("streamOffset", C.c_uint64), ("ellBegin", C.c_int), ("ellEnd", C.c_int),
                ("poolPackets", C.c_int), ("continuousScattering", C.c_int)]


class SkgMcStats(C.Structure):
    _fields_ = [("packets", C.c_uint64), ("pathSegments", C.c_uint64), ("paths", C.c_uint64),
                ("scatterings", C.c_uint64), ("kernel_ms", C.c_double),
                ("absorbSegments", C.c_uint64), ("detections", C.c_uint64),
                ("launch_ms", C.c_double), ("peel_ms", C.c_double), ("absorb_ms", C.c_double), ("propagate_ms", C.c_double),
                ("iterations", C.c_uint64), ("peelSegments", C.c_uint64), ("propagateSegments", C.c_uint64)]


def _stats(st):
    return {name: getattr(st, name) for name, _ in SkgMcStats._fields_}


def lib_available():
    return os.path.exists(LIB_PATH)


def load_library():
    """Loads libskirtgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: build it with `make` or __graft_entry__.build(); "
                              "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.skg_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))      # raw device pointer


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """One engine per GPU (skg_engine).  Mirrors the call sequence of the reference's setup:
    grid -> medium (DustSystem) -> sources (StellarSystem) -> instruments -> run."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        self._chk(self._lib.skg_engine_create(int(device), C.byref(h)))
        self.h = h
        self.device = device
        self.Nlambda = 0
        self._keep = []

    def _chk(self, rc):
        if rc:
            raise EngineError(self._lib.skg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self._lib.skg_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        """cudaStream_t of the engine as an integer (wrap with torch.cuda.ExternalStream to record events on it)"""
        p = C.c_void_p()
        self._chk(self._lib.skg_stream(self.h, C.byref(p)))
        return p.value or 0

    @property
    def launch_count(self):
        n = C.c_uint64()
        self._chk(self._lib.skg_launch_count(self.h, C.byref(n)))
        return n.value

    # ---- grids ---------------------------------------------------------------------------------
    def grid_cartesian(self, xv, yv, zv):
        xv, yv, zv = _f64(xv), _f64(yv), _f64(zv)
        self._chk(self._lib.skg_grid_cartesian(self.h, _vp(xv), len(xv) - 1, _vp(yv), len(yv) - 1, _vp(zv), len(zv) - 1))

    def grid_tree(self, kind, search, box, child0, parent, cell, dir=None, nbrStart=None, nbrIds=None):
        box = _f64(box); child0 = _i32(child0); parent = _i32(parent); cell = _i32(cell)
        dir = None if dir is None else _i32(dir)
        nbrStart = None if nbrStart is None else _i32(nbrStart)
        nbrIds = None if nbrIds is None else _i32(nbrIds if len(nbrIds) else [0])
        self._chk(self._lib.skg_grid_tree(self.h, int(kind), int(search), len(child0), _vp(box), _vp(child0), _vp(parent),
                                          _vp(cell), _vp(dir), _vp(nbrStart), _vp(nbrIds)))

    def grid_amesh(self, box, nxyz, child0, cell, wallNbr):
        box = _f64(box); nxyz = _i32(nxyz); child0 = _i32(child0); cell = _i32(cell); wallNbr = _i32(wallNbr)
        self._chk(self._lib.skg_grid_amesh(self.h, len(child0), _vp(box), _vp(nxyz), _vp(child0), _vp(cell), _vp(wallNbr)))

    def grid_voronoi(self, t):
        ints = ("nbrStart", "nbrIds", "blkStart", "blkIds", "blkTree", "kdM", "kdAxis", "kdUp", "kdLeft", "kdRight")
        a = {k: (_f64(v) if k in ("particles", "cellBox", "extent") else _i32(v) if k in ints else v) for k, v in t.items()}
        nkd = len(a["kdM"])
        pad = lambda v: v if len(v) else np.zeros(1, np.int32)
        self._chk(self._lib.skg_grid_voronoi(
            self.h, len(a["particles"]), _vp(a["particles"]), _vp(a["nbrStart"]), _vp(pad(a["nbrIds"])), _vp(a["extent"]),
            int(a["nb"]), _vp(a["blkStart"]), _vp(pad(a["blkIds"])), _vp(a["blkTree"]), nkd, _vp(pad(a["kdM"])),
            _vp(pad(a["kdAxis"])), _vp(pad(a["kdUp"])), _vp(pad(a["kdLeft"])), _vp(pad(a["kdRight"])), _vp(a.get("cellBox"))))

    def set_grid(self, t):
        """t: dict in the layout produced by the host-side grid builders (kind + tables)."""
        kind = t["kind"]
        if kind == "cartesian":
            self.grid_cartesian(t["xv"], t["yv"], t["zv"])
        elif kind in ("octtree", "bintree"):
            self.grid_tree(0 if kind == "octtree" else 1, t["search"], t["box"], t["child0"], t["parent"], t["cell"],
                           t.get("dir"), t.get("nbrStart"), t.get("nbrIds"))
        elif kind == "amesh":
            self.grid_amesh(t["box"], t["nxyz"], t["child0"], t["cell"], t["wallNbr"])
        elif kind == "voronoi":
            self.grid_voronoi(t)
        elif kind == "sphere1d":
            rv = _f64(t["rv"]); self._keep = (rv,)
            self._chk(self._lib.skg_grid_sphere1d(self.h, len(rv) - 1, _vp(rv)))
        elif kind == "sphere2d":
            rv, th, cv = _f64(t["rv"]), _f64(t["thetav"]), _f64(t["cv"]); self._keep = (rv, th, cv)
            self._chk(self._lib.skg_grid_sphere2d(self.h, len(rv) - 1, _vp(rv), len(th) - 1, _vp(th), _vp(cv)))
        elif kind == "cylinder2d":
            Rv, zv = _f64(t["Rv"]), _f64(t["zv"]); self._keep = (Rv, zv)
            self._chk(self._lib.skg_grid_cylinder2d(self.h, len(Rv) - 1, _vp(Rv), len(zv) - 1, _vp(zv)))
        else:
            raise EngineError(f"unknown grid kind {kind}")

    @property
    def Ncells(self):
        return self._lib.skg_num_cells(self.h)

    def medium(self, rho, kext, ksca=None, g=None):
        rho = _f64(rho); kext = _f64(kext)
        if rho.ndim == 1:
            rho = rho[:, None]
        kext = np.atleast_2d(kext)
        Ncells, Ncomp = rho.shape
        Nlambda = kext.shape[1]
        ksca = None if ksca is None else _f64(np.atleast_2d(ksca))
        g = None if g is None else _f64(np.atleast_2d(g))
        self._chk(self._lib.skg_medium(self.h, Ncells, Ncomp, Nlambda, _vp(rho), _vp(kext), _vp(ksca), _vp(g)))
        self.Nlambda = Nlambda
        self.Ncomp = Ncomp
        self.polarized = False

    def medium_polarization(self, S11, S12, S33, S34):
        """skg_medium_polarization: Mueller matrix coefficients [Ncomp, Nlambda, Ntheta] of every dust component"""
        a = [np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(self.Ncomp, self.Nlambda, -1)) for v in (S11, S12, S33, S34)]
        self._chk(self._lib.skg_medium_polarization(self.h, a[0].shape[2], _vp(a[0]), _vp(a[1]), _vp(a[2]), _vp(a[3])))
        self.polarized = True

    # ---- deterministic geometry ------------------------------------------------------------------
    def path_batch(self, r, k, ell=None):
        """Batched DustGrid::path()+fillOpticalDepth() for host rays -> CSR dict (numpy)."""
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        off = np.zeros(n + 1, np.int64); total = C.c_int64()
        self._chk(self._lib.skg_path_count(self.h, SKG_HOST, C.c_int64(n), _vp(r), _vp(k), _vp(off), C.byref(total)))
        seg = np.zeros(max(total.value, 1), dtype=SEGMENT)
        ellp, stride = None, 0
        if ell is not None:
            ella = _i32(np.atleast_1d(ell)); ellp = _vp(ella); stride = 1 if len(ella) == n and n > 1 else 0
            if len(ella) not in (1, n):
                raise EngineError("ell must be a scalar or have one entry per ray")
        self._chk(self._lib.skg_path_fill(self.h, SKG_HOST, C.c_int64(n), _vp(r), _vp(k), ellp, stride, _vp(off), _vp(seg)))
        seg = seg[:total.value]
        return dict(offsets=off, **{f: np.ascontiguousarray(seg[f]) for f in ("m", "ds", "s", "dtau", "tau")})

    def path_batch_onepass(self, r, k, ell=None):
        """skg_path_batch for host rays: one traversal per ray into slabs; returns the same CSR dict as path_batch (the slabs
        compacted on the host) plus the raw starts / lengths and the number of records the slabs took"""
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        starts = np.zeros(n + 1, np.int64); lengths = np.zeros(max(n, 1), np.int32); needed = C.c_int64()
        ellp, stride = None, 0
        if ell is not None:
            ella = _i32(np.atleast_1d(ell)); ellp = _vp(ella); stride = 1 if len(ella) == n and n > 1 else 0
        args = (self.h, SKG_HOST, C.c_int64(n), _vp(r), _vp(k), ellp, stride, _vp(starts), _vp(lengths))
        self._chk(self._lib.skg_path_batch(*args, None, C.c_int64(0), C.byref(needed)))          # size query
        seg = np.zeros(max(needed.value, 1), dtype=SEGMENT)
        self._chk(self._lib.skg_path_batch(*args, _vp(seg), C.c_int64(len(seg)), C.byref(needed)))
        lengths = lengths[:n]
        off = np.zeros(n + 1, np.int64); np.cumsum(lengths, out=off[1:])
        idx = np.repeat(starts[:n] - off[:n], lengths) + np.arange(off[n]) if n else np.zeros(0, np.int64)
        out = seg[idx]
        return dict(offsets=off, starts=starts, lengths=lengths, slab_records=int(needed.value),
                    **{f: np.ascontiguousarray(out[f]) for f in ("m", "ds", "s", "dtau", "tau")})

    def path_batch_device(self, n, d_r, d_k, d_ell, ell_stride, d_starts, d_lengths, d_segments, capacity):
        """skg_path_batch on device arrays; returns the number of records needed (query with d_segments = None)"""
        needed = C.c_int64()
        self._chk(self._lib.skg_path_batch(self.h, SKG_DEVICE, C.c_int64(n), _vp(d_r), _vp(d_k), _vp(d_ell), int(ell_stride), _vp(d_starts),
                                           _vp(d_lengths), _vp(d_segments), C.c_int64(int(capacity)), C.byref(needed)))
        return needed.value

    def path_count_device(self, n, d_r, d_k, d_offsets):
        total = C.c_int64()
        self._chk(self._lib.skg_path_count(self.h, SKG_DEVICE, C.c_int64(n), _vp(d_r), _vp(d_k), _vp(d_offsets), C.byref(total)))
        return total.value

    def path_fill_device(self, n, d_r, d_k, d_ell, ell_stride, d_offsets, d_segments):
        """d_segments: device array of total x 40-byte skg_segment records (32-byte aligned)"""
        self._chk(self._lib.skg_path_fill(self.h, SKG_DEVICE, C.c_int64(n), _vp(d_r), _vp(d_k), _vp(d_ell), int(ell_stride),
                                          _vp(d_offsets), _vp(d_segments)))

    def opticaldepth(self, r, k, ell, distance=None, mc_walker=False):
        """DustSystem::opticaldepth for host rays; mc_walker=True: with the walker of the shooting stages (skg_opticaldepth_mc)"""
        r = _f64(r).reshape(-1, 3); k = _f64(k).reshape(-1, 3); n = len(r)
        ella = _i32(np.atleast_1d(ell)); stride = 1 if len(ella) == n and n > 1 else 0
        d = None if distance is None else _f64(distance)
        tau = np.zeros(n)
        fn = self._lib.skg_opticaldepth_mc if mc_walker else self._lib.skg_opticaldepth
        self._chk(fn(self.h, SKG_HOST, C.c_int64(n), _vp(r), _vp(k), _vp(ella), stride, _vp(d), _vp(tau)))
        return tau

    def whichcell(self, r):
        r = _f64(r).reshape(-1, 3); m = np.zeros(len(r), np.int32)
        self._chk(self._lib.skg_whichcell(self.h, SKG_HOST, C.c_int64(len(r)), _vp(r), _vp(m)))
        return m

    def selftest_division(self, n, seed=1):
        bad = C.c_uint64()
        self._chk(self._lib.skg_selftest_division(self.h, C.c_uint64(int(n)), C.c_uint64(int(seed)), C.byref(bad)))
        return bad.value

    def selftest_atomics(self, n=1 << 30, cells=1000000):
        """fp64 atomicAdds per second to random cells of a table (skg_selftest_atomics)"""
        r = C.c_double()
        self._chk(self._lib.skg_selftest_atomics(self.h, C.c_uint64(int(n)), int(cells), C.byref(r)))
        return r.value

    def stuck_counts(self):
        a = C.c_int64(); b = C.c_int64()
        self._chk(self._lib.skg_stuck_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- Monte Carlo -----------------------------------------------------------------------------
    def _source_array(self, comps):
        arr = (SkgSource * len(comps))()
        keep = []
        for i, c in enumerate(comps):
            s = arr[i]
            s.geometry = int(c["geometry"])
            p = list(c.get("p", [])) + [0.0] * 8
            for j in range(8):
                s.p[j] = float(p[j])
            sp = c.get("spiral")
            if sp:
                s.spiral_arms = int(sp["arms"]); s.spiral_index = int(sp["index"]); s.spiral_pitch = float(sp["pitch"])
                s.spiral_radius = float(sp["radius"]); s.spiral_phase = float(sp["phase"]); s.spiral_weight = float(sp["weight"])
            if c.get("rv") is not None:
                rv = _f64(c["rv"]); Xv = _f64(c["Xv"]); keep += [rv, Xv]
                s.ntab = len(rv); s.rv = rv.ctypes.data; s.Xv = Xv.ctypes.data
                if c.get("Sv") is not None:
                    Sv = _f64(c["Sv"]); keep.append(Sv); s.Sv = Sv.ctypes.data
        return arr, keep

    def sources(self, comps, L, emission_bias=0.5):
        """comps: list of dicts(geometry=..., p=[...], spiral=dict|None, rv=..., Xv=...); L[Ncomp, Nlambda]."""
        L = _f64(np.atleast_2d(L))
        arr, keep = self._source_array(comps)
        self._chk(self._lib.skg_sources(self.h, len(comps), arr, L.shape[1], _vp(L), C.c_double(emission_bias)))

    def sample_density(self, geometries, norm, sample_count=100, seed=4357):
        """DustSystem::setSampleDensityBody on the device -> rho[Ncells, Ncomp]"""
        arr, keep = self._source_array(geometries)
        nrm = _f64(norm)
        rho = np.zeros((self.Ncells, len(geometries)))
        self._chk(self._lib.skg_sample_density(self.h, len(geometries), arr, _vp(nrm), int(sample_count), C.c_uint64(int(seed)), _vp(rho)))
        return rho

    def sample_boxes(self, boxes, geometries, norm, sample_count=100, seed=4357, dispersion=False):
        """skg_sample_boxes: dust mass in every box (TreeNodeSampleDensityCalculator on the device); boxes[n, 6] =
        xmin,ymin,zmin,xmax,ymax,zmax.  dispersion=True: skg_sample_boxes_dispersion, returns (mass, densityDispersion)"""
        arr, keep = self._source_array(geometries)
        nrm = _f64(norm); b = _f64(boxes).reshape(-1, 6)
        mass = np.zeros(len(b))
        if dispersion:
            disp = np.zeros(len(b))
            self._chk(self._lib.skg_sample_boxes_dispersion(self.h, C.c_int64(len(b)), _vp(b), len(geometries), arr, _vp(nrm),
                                                            int(sample_count), C.c_uint64(int(seed)), _vp(mass), _vp(disp)))
            return mass, disp
        self._chk(self._lib.skg_sample_boxes(self.h, C.c_int64(len(b)), _vp(b), len(geometries), arr, _vp(nrm), int(sample_count),
                                             C.c_uint64(int(seed)), _vp(mass)))
        return mass

    def sample_launch(self, ell, n, seed=1):
        r = np.zeros((n, 3)); k = np.zeros((n, 3)); L = np.zeros(n)
        self._chk(self._lib.skg_sample_launch(self.h, int(ell), int(n), C.c_uint64(int(seed)), _vp(r), _vp(k), _vp(L)))
        return r, k, L

    def instruments(self, instr):
        arr = (SkgInstrument * len(instr))(); keep = []
        for i, d in enumerate(instr):
            a = arr[i]
            a.kind = int(d["kind"]); a.distance = float(d.get("distance", 0.0)); a.inclination = float(d.get("inclination", 0.0))
            for key in ("viewX", "viewY", "viewZ", "crossX", "crossY", "crossZ", "upX", "upY", "upZ", "focal"):      # PerspectiveInstrument
                setattr(a, key, float(d.get(key, 0.0)))
            a.azimuth = float(d.get("azimuth", 0.0)); a.positionAngle = float(d.get("positionAngle", 0.0))
            a.Nxp = int(d.get("Nxp", 0)); a.Nyp = int(d.get("Nyp", 0))
            a.fovxp = float(d.get("fovxp", 0.0)); a.fovyp = float(d.get("fovyp", 0.0))
            a.xpc = float(d.get("xpc", 0.0)); a.ypc = float(d.get("ypc", 0.0))
            a.scatteringLevels = int(d.get("scatteringLevels", 0))
            a.writeTotal = int(bool(d.get("writeTotal", True))); a.writeStellarComps = int(bool(d.get("writeStellarComps", False)))
            if d.get("frames") is not None:         # MultiFrameInstrument: one dict(Nxp, Nyp, fovxp, fovyp[, xpc, ypc]) per wavelength
                fr = (SkgInstrumentFrame * len(d["frames"]))()
                if len(d["frames"]) != self.Nlambda:
                    raise EngineError("Number of instrument frames must equal number of wavelengths")
                for q, f in enumerate(d["frames"]):
                    fr[q].Nxp = int(f["Nxp"]); fr[q].Nyp = int(f["Nyp"]); fr[q].fovxp = float(f["fovxp"]); fr[q].fovyp = float(f["fovyp"])
                    fr[q].xpc = float(f.get("xpc", 0.0)); fr[q].ypc = float(f.get("ypc", 0.0))
                keep.append(fr); a.frames = fr
        self._instr = list(instr)
        self._chk(self._lib.skg_instruments(self.h, len(instr), arr))

    def run_stellar(self, packages, total_packages=None, min_weight_reduction=1e4, min_scatt_events=0.0, scatt_bias=0.5,
                    store_absorption=False, seed=4357, stream_offset=0, ell_begin=0, ell_end=None, pool_packets=0, continuous_scattering=False):
        p = SkgMcParams()
        p.continuousScattering = int(bool(continuous_scattering))
        p.packages = float(packages)
        p.luminosityScale = float(total_packages if total_packages is not None else packages)
        p.minWeightReduction = float(min_weight_reduction); p.minScattEvents = float(min_scatt_events)
        p.scattBias = float(scatt_bias); p.storeAbsorption = int(bool(store_absorption))
        p.seed = int(seed); p.streamOffset = int(stream_offset)
        p.poolPackets = int(pool_packets)
        p.ellBegin = int(ell_begin); p.ellEnd = int(self.Nlambda if ell_end is None else ell_end)
        st = SkgMcStats()
        self._chk(self._lib.skg_run_stellar(self.h, C.byref(p), C.byref(st)))
        return _stats(st)

    def _params(self, packages, total_packages, min_weight_reduction, min_scatt_events, scatt_bias, store_absorption, seed,
                stream_offset, ell_begin, ell_end, pool_packets):
        p = SkgMcParams()
        p.packages = float(packages)
        p.luminosityScale = float(total_packages if total_packages is not None else packages)
        p.minWeightReduction = float(min_weight_reduction); p.minScattEvents = float(min_scatt_events)
        p.scattBias = float(scatt_bias); p.storeAbsorption = int(bool(store_absorption))
        p.seed = int(seed); p.streamOffset = int(stream_offset); p.poolPackets = int(pool_packets)
        p.ellBegin = int(ell_begin); p.ellEnd = int(self.Nlambda if ell_end is None else ell_end)
        return p

    def run_dust(self, phase, Lcell, packages, total_packages=None, emission_bias=0.5, min_weight_reduction=1e4,
                 min_scatt_events=0.0, scatt_bias=0.5, seed=4357, stream_offset=0, ell_begin=0, ell_end=None, pool_packets=0):
        """PanMonteCarloSimulation::dodustselfabsorptionchunk / dodustemissionchunk for all wavelengths;
        Lcell[Nlambda, Ncells] = Labsbol[m] * dustluminosity(m, ell)"""
        Lc = _f64(Lcell)
        if Lc.shape != (self.Nlambda, self.Ncells):
            raise EngineError(f"Lcell must have shape (Nlambda, Ncells) = ({self.Nlambda}, {self.Ncells})")
        p = self._params(packages, total_packages, min_weight_reduction, min_scatt_events, scatt_bias, False, seed,
                         stream_offset, ell_begin, ell_end, pool_packets)
        st = SkgMcStats()
        self._chk(self._lib.skg_run_dust(self.h, C.byref(p), int(phase), C.c_double(emission_bias), SKG_HOST, _vp(Lc), C.byref(st)))
        return _stats(st)

    def dust_library(self, volumes, kappaabs, lambdav, dlambdav):
        v = _f64(volumes); k = _f64(np.atleast_2d(kappaabs)); lam = _f64(lambdav); dl = _f64(dlambdav)
        self._chk(self._lib.skg_dust_library(self.h, _vp(v), _vp(k), _vp(lam), _vp(dl)))

    def dust_cell_luminosities(self):
        """device pointer of Lcell[Nlambda, Ncells] computed from the device-resident absorption tables"""
        p = C.c_void_p()
        self._chk(self._lib.skg_dust_cell_luminosities(self.h, C.byref(p)))
        return p.value

    def run_dust_device(self, phase, d_Lcell, packages, total_packages=None, emission_bias=0.5, min_weight_reduction=1e4,
                        min_scatt_events=0.0, scatt_bias=0.5, seed=4357, stream_offset=0, pool_packets=0):
        p = self._params(packages, total_packages, min_weight_reduction, min_scatt_events, scatt_bias, False, seed,
                         stream_offset, 0, None, pool_packets)
        st = SkgMcStats()
        self._chk(self._lib.skg_run_dust(self.h, C.byref(p), int(phase), C.c_double(emission_bias), SKG_DEVICE, C.c_void_p(int(d_Lcell)), C.byref(st)))
        return _stats(st)

    def copy_from_device(self, d_ptr, shape):
        out = np.zeros(shape)
        self._chk(self._lib.skg_copy_to_host(self.h, C.c_void_p(int(d_ptr)), _vp(out), C.c_size_t(out.nbytes)))
        return out

    def reset_labs_dust(self):
        self._chk(self._lib.skg_reset_labs_dust(self.h))

    def fetch_labs_dust(self):
        a = np.zeros((self.Ncells, self.Nlambda))
        self._chk(self._lib.skg_fetch_labs_dust(self.h, _vp(a), 0))
        return a

    def labs_bolometric(self):
        a = np.zeros(self.Ncells)
        self._chk(self._lib.skg_labs_bolometric(self.h, _vp(a)))
        return a

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array in page-locked host memory (skg_host_alloc); freed when the array is garbage collected"""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._chk(self._lib.skg_host_alloc(C.c_size_t(n), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        lib = self._lib
        import weakref
        weakref.finalize(buf, lambda ptr=p.value: lib.skg_host_free(C.c_void_p(ptr)))
        return arr

    def reset_results(self):
        self._chk(self._lib.skg_reset_results(self.h))

    def fetch_frame(self, i, out=None):
        d = self._instr[i]
        a = np.zeros(int(d["Nxp"]) * int(d["Nyp"]) * self.Nlambda) if out is None else out
        self._chk(self._lib.skg_fetch_frame(self.h, i, _vp(a), 0))
        return a

    def fetch_sed(self, i, out=None):
        a = np.zeros(self.Nlambda) if out is None else out
        self._chk(self._lib.skg_fetch_sed(self.h, i, _vp(a), 0))
        return a

    def fetch_frame_channel(self, i, channel, out=None):
        """data cube of one FullInstrument channel (CHAN_*)"""
        d = self._instr[i]
        a = np.zeros(int(d["Nxp"]) * int(d["Nyp"]) * self.Nlambda) if out is None else out
        self._chk(self._lib.skg_fetch_frame_channel(self.h, i, int(channel), _vp(a), 0))
        return a

    def fetch_sed_channel(self, i, channel, out=None):
        a = np.zeros(self.Nlambda) if out is None else out
        self._chk(self._lib.skg_fetch_sed_channel(self.h, i, int(channel), _vp(a), 0))
        return a

    def fetch_labs(self, out=None):
        a = np.zeros((self.Ncells, self.Nlambda)) if out is None else out
        self._chk(self._lib.skg_fetch_labs(self.h, _vp(a), 0))
        return a

    def fetch_multiframe(self, i, which, ell):
        """one frame of a MultiFrameInstrument: which = -1 the total flux, k >= 0 stellar component k; [Nyp, Nxp] of frame ell"""
        f = self._instr[i]["frames"][ell]
        out = np.zeros((int(f["Nyp"]), int(f["Nxp"])))
        self._chk(self._lib.skg_fetch_multiframe(self.h, int(i), int(which), int(ell), _vp(out), 0))
        return out

    def results_snapshot(self):
        """skg_results_snapshot: shadow copies of every accumulator, to be fetched while the engine goes on"""
        self._chk(self._lib.skg_results_snapshot(self.h))

    def fetch_snapshot_async(self, which, part, out=None):
        """starts the transfer of one shadow array into `out` (page-locked, see pinned_empty); returns the element count"""
        n = C.c_int64()
        self._chk(self._lib.skg_fetch_snapshot_async(self.h, int(which), int(part), _vp(out), C.byref(n)))
        return n.value

    def fetch_snapshot_wait(self):
        self._chk(self._lib.skg_fetch_snapshot_wait(self.h))

    def device_accumulator(self, which, part=0):
        p = C.c_void_p(); n = C.c_int64()
        self._chk(self._lib.skg_device_accumulators(self.h, int(which), int(part), C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- multi-GPU -----------------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = np.zeros(128, np.uint8)
        self._chk(self._lib.skg_comm_unique_id(_vp(buf)))
        return buf

    def comm_init(self, rank, nranks, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._chk(self._lib.skg_comm_init(self.h, int(rank), int(nranks), _vp(uid)))

    def allreduce(self, which=REDUCE_ALL):
        """skg_allreduce: in-place sum over the ranks of the selected accumulators (each at most once, see skirtgpu.h);
        returns the device time of the collective in ms"""
        ms = C.c_double()
        self._chk(self._lib.skg_allreduce(self.h, int(which), C.byref(ms)))
        return ms.value

    def allreduce_results(self):
        self._chk(self._lib.skg_allreduce_results(self.h))

    def labs_dust_total(self):
        """PanDustSystem::Labsdusttot(): over all cells, wavelengths and processes; identical on every rank"""
        t = C.c_double()
        self._chk(self._lib.skg_labs_dust_total(self.h, C.byref(t)))
        return t.value

    def labs_stellar_total(self):
        t = C.c_double()
        self._chk(self._lib.skg_labs_stellar_total(self.h, C.byref(t)))
        return t.value
