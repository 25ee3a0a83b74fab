"""ctypes binding of include/skirthost.h (libskirthost.so): host-side construction of the tree, adaptive-mesh and Voronoi
grids as the flat tables of include/skirtgpu.h.  Native C++ (skirt_b200/host/GridBuilders.cpp); there is no Python
fallback: a missing library raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskirthost.so")
_lib = None


class HostError(RuntimeError):
    pass


def lib_available():
    return os.path.exists(LIB_PATH)


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} is missing: build it with `make skirt_b200/libskirthost.so`")
        L = C.CDLL(LIB_PATH)
        L.skh_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _chk(rc):
    if rc:
        raise HostError(load_library().skh_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def voronoi_available():
    return bool(load_library().skh_voronoi_available())


class TreeBuilder:
    """skh_tree_*: the tree grows one level per subdivide() call; `decide(boxes) -> bool[n]` is asked for the levels
    between minLevel and maxLevel (TreeDustGrid::subdivide, TreeDustGrid.cpp:168-233)"""

    def __init__(self, kind, extent, minLevel, maxLevel):
        self.L = load_library()
        self.h = C.c_void_p()
        ext = np.ascontiguousarray(extent, dtype=np.float64)
        _chk(self.L.skh_tree_create(int(kind), _p(ext), int(minLevel), int(maxLevel), C.byref(self.h)))
        self.kind = int(kind)

    def close(self):
        if getattr(self, "h", None):
            self.L.skh_tree_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frontier(self):
        level = C.c_int(); size = C.c_int64(); need = C.c_int()
        _chk(self.L.skh_tree_frontier(self.h, C.byref(level), C.byref(size), C.byref(need)))
        return level.value, size.value, bool(need.value)

    def frontier_boxes(self):
        _, n, _ = self.frontier()
        box = np.zeros((n, 6))
        if n:
            _chk(self.L.skh_tree_frontier_boxes(self.h, _p(box)))
        return box

    def subdivide(self, flags=None, barycenters=None):
        f = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        if barycenters is None:
            _chk(self.L.skh_tree_subdivide(self.h, _p(f)))
        else:
            b = np.ascontiguousarray(barycenters, dtype=np.float64).reshape(-1, 3)
            if f is not None and len(b) != len(f):
                raise HostError("one barycentre per frontier node is needed")
            _chk(self.L.skh_tree_subdivide_at(self.h, _p(f), _p(b)))

    def grow(self, decide):
        """runs the whole subdivision loop; decide(level, boxes) returns the flags, or (flags, barycenters[n, 3]) for
        barycentric subdivision"""
        while True:
            level, n, need = self.frontier()
            if n == 0:
                return
            d = decide(level, self.frontier_boxes()) if need else None
            if isinstance(d, tuple):
                self.subdivide(d[0], d[1])
            else:
                self.subdivide(d)

    def finish(self, search=1):
        nn = C.c_int(); nc = C.c_int(); nb = C.c_int64()
        _chk(self.L.skh_tree_finish(self.h, int(search), C.byref(nn), C.byref(nc), C.byref(nb)))
        N = nn.value
        t = dict(kind="octtree" if self.kind == 0 else "bintree", search=int(search), box=np.zeros(6 * N), child0=np.zeros(N, np.int32),
                 parent=np.zeros(N, np.int32), cell=np.zeros(N, np.int32), dir=np.zeros(N, np.int32), level=np.zeros(N, np.int32))
        if search == 1:
            t["nbrStart"] = np.zeros(6 * N + 1, np.int32); t["nbrIds"] = np.zeros(max(nb.value, 1), np.int32)
        _chk(self.L.skh_tree_tables(self.h, _p(t["box"]), _p(t["child0"]), _p(t["parent"]), _p(t["cell"]), _p(t["dir"]), _p(t["level"]),
                                    _p(t.get("nbrStart")), _p(t.get("nbrIds"))))
        if search == 1:
            t["nbrIds"] = t["nbrIds"][:nb.value]
        t["Ncells"] = nc.value
        return t


def build_particle_tree(kind, extent, particles, extraLevels=0):
    """skh_ptree_build: the tree of ParticleTreeDustGrid (ParticleTreeDustGrid.cpp:76-152) around particles[n,3]; kind 0 octree, 1 binary
    tree; tables for skg_grid_tree with search = 3"""
    L = load_library()
    ext = np.ascontiguousarray(extent, dtype=np.float64)
    pts = np.ascontiguousarray(particles, dtype=np.float64).reshape(-1, 3)
    h = C.c_void_p(); nn = C.c_int(); nc = C.c_int()
    _chk(L.skh_ptree_build(int(kind), _p(ext), _p(pts), C.c_int64(len(pts)), int(extraLevels), C.byref(h), C.byref(nn), C.byref(nc)))
    try:
        N = nn.value
        t = dict(kind="octtree" if int(kind) == 0 else "bintree", search=3, box=np.zeros(6 * N), child0=np.zeros(N, np.int32),
                 parent=np.zeros(N, np.int32), cell=np.zeros(N, np.int32), dir=np.zeros(N, np.int32), level=np.zeros(N, np.int32))
        _chk(L.skh_tree_tables(h, _p(t["box"]), _p(t["child0"]), _p(t["parent"]), _p(t["cell"]), _p(t["dir"]), _p(t["level"]), None, None))
    finally:
        L.skh_tree_destroy(h)
    t["Ncells"] = nc.value
    return t


def build_adaptive_mesh(extent, nxyz):
    """skh_amesh_*: tables of skg_grid_amesh + cell volumes and, per cell, the index of its line in the input sequence"""
    L = load_library()
    ext = np.ascontiguousarray(extent, dtype=np.float64)
    nx = np.ascontiguousarray(nxyz, dtype=np.int32).reshape(-1, 3)
    h = C.c_void_p(); nn = C.c_int(); nc = C.c_int()
    _chk(L.skh_amesh_build(_p(ext), _p(nx), C.c_int64(len(nx)), C.byref(h), C.byref(nn), C.byref(nc)))
    try:
        N, M = nn.value, nc.value
        t = dict(kind="amesh", box=np.zeros(6 * N), nxyz=np.zeros(3 * N, np.int32), child0=np.zeros(N, np.int32), cell=np.zeros(N, np.int32),
                 wallNbr=np.zeros(6 * N, np.int32), volume=np.zeros(M), fileIndex=np.zeros(M, np.int32))
        _chk(L.skh_amesh_tables(h, _p(t["box"]), _p(t["nxyz"]), _p(t["child0"]), _p(t["cell"]), _p(t["wallNbr"]), _p(t["volume"]), _p(t["fileIndex"])))
    finally:
        L.skh_amesh_destroy(h)
    t["Ncells"] = M
    return t


def build_voronoi_mesh(extent, particles):
    """skh_voronoi_*: tables of skg_grid_voronoi (+ cell volumes and centroids) from the particle positions, via Voro++"""
    L = load_library()
    ext = np.ascontiguousarray(extent, dtype=np.float64)
    pts = np.ascontiguousarray(particles, dtype=np.float64).reshape(-1, 3)
    h = C.c_void_p(); sizes = np.zeros(5, np.int64)
    _chk(L.skh_voronoi_build(_p(ext), _p(pts), C.c_int64(len(pts)), C.byref(h), _p(sizes)))
    try:
        N, nn, nb, nr, nk = (int(v) for v in sizes)
        i32 = lambda n: np.zeros(max(n, 1), np.int32)
        t = dict(kind="voronoi", particles=pts, extent=ext, nb=nb, cellBox=np.zeros(6 * N), volume=np.zeros(N), centroid=np.zeros(3 * N),
                 nbrStart=i32(N + 1), nbrIds=i32(nn), blkStart=i32(nb ** 3 + 1), blkIds=i32(nr), blkTree=i32(nb ** 3),
                 kdM=i32(nk), kdAxis=i32(nk), kdUp=i32(nk), kdLeft=i32(nk), kdRight=i32(nk))
        _chk(L.skh_voronoi_tables(h, _p(t["cellBox"]), _p(t["volume"]), _p(t["centroid"]), _p(t["nbrStart"]), _p(t["nbrIds"]), _p(t["blkStart"]),
                                  _p(t["blkIds"]), _p(t["blkTree"]), _p(t["kdM"]), _p(t["kdAxis"]), _p(t["kdUp"]), _p(t["kdLeft"]), _p(t["kdRight"])))
        for k, n in (("nbrIds", nn), ("blkIds", nr), ("kdM", nk), ("kdAxis", nk), ("kdUp", nk), ("kdLeft", nk), ("kdRight", nk)):
            t[k] = t[k][:n]
    finally:
        L.skh_voronoi_destroy(h)
    t["Ncells"] = N
    return t
