// Grid traversal device functions: one stepping walker per grid type, each a restructured-for-GPU rendition of
// the reference's DustGrid::path() that yields one segment per step instead of filling a std::vector.
//
// Bit-exactness contract: compiled with -fmad=false, IEEE division, the reference's
// operand order, comparison asymmetries and eps conventions, so that cell sequences are identical and
// ds/s/dtau/tau agree to the last bit with the reference built without FMA contraction.
//
//   Cartesian      CartesianDustGrid.cpp:136-283 (+ NR::locate_clip, NR.hpp:99-112,146-151)
//   Tree           TreeDustGrid.cpp:390-662 (+ DustGridPath::moveInside DustGridPath.cpp:57-150,
//                  TreeNode::whichnode TreeNode.cpp:70-93, OctTreeNode::child OctTreeNode.cpp:184-189,
//                  BinTreeNode::child BinTreeNode.cpp:326-335)
//   AdaptiveMesh   AdaptiveMesh.cpp:297-367 (+ AdaptiveMeshNode.cpp:109-151, Box::cellindices Box.hpp:134-139)
//   Voronoi        VoronoiMesh.cpp:749-844 (+ cellIndex :512-541, kd Node::nearest :180-225)
//
// Every grid type provides a stepping walker: start() = entry (moveInside) + point location, step() = one
// crossing, returning whether a segment (m, ds) is to be added (DustGridPath::addSegment drops ds <= 0,
// DustGridPath.cpp:46-53).  The scheduler in wavefront.cuh drives them and hands the segments to a job.
#pragma once
#include <cfloat>
#include <cmath>
#include "tables.h"

namespace skg
{

#define SKG_DBL_MAX 1.7976931348623157e308

// ---------------------------------------------------------------------------------------------------
// pending "outside" segments: the reference adds up to three m=-1 segments while moving a ray into
// the grid and clears them again if the ray turns out to miss the grid; we hold them back until the
// entry is confirmed.
struct Entry
{
    double ds[3];
    int n;
};

__device__ __forceinline__ bool finite3(double a, double b, double c)
{
    return isfinite(a) && isfinite(b) && isfinite(c);
}

// a / b for a divisor that stays the same over a whole path (the direction cosines), given rb = 1.0 / b computed
// once with an IEEE division.  q0 = a*rb is within 2 ulp of a/b; one Newton correction with the exactly computed
// residual (FMA) makes it a faithful rounding, and by Markstein's theorem (Markstein 1990; Muller et al., Handbook of
// Floating-Point Arithmetic, 2nd ed., section 4.7) a second one with rb = RN(1/b) yields the correctly rounded
// quotient RN(a/b) -- bit for bit what the reference's `/` gives, at 5 fp64 operations instead of the ~14-instruction
// general division sequence with its slow-path branch.  Verified against `/` on the device by skg_selftest_division.
__device__ __forceinline__ double divInvariant(double a, double b, double rb)
{
    double q = a * rb;
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, rb, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, rb, q);
}

// NR::locate_basic_impl, NR.hpp:99-112
__device__ __forceinline__ int locateBasic(const double* xv, double x, int n)
{
    int jl = -1, ju = n;
    while (ju - jl > 1)
    {
        int jm = (ju + jl) >> 1;
        if (x < xv[jm]) ju = jm; else jl = jm;
    }
    return jl;
}
// NR::locate_clip over an array of n values, NR.hpp:146-151
__device__ __forceinline__ int locateClip(const double* xv, double x, int n)
{
    if (x < xv[0]) return 0;
    return locateBasic(xv, x, n - 1);
}
// NR::locate_fail, NR.hpp:155-160
__device__ __forceinline__ int locateFail(const double* xv, double x, int n)
{
    if (x > xv[n - 1]) return -1;
    return locateBasic(xv, x, n - 1);
}

// ---------------------------------------------------------------------------------------------------
// Cartesian
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cartWhichCell(const CartGrid& g, double x, double y, double z)
{
    // CartesianDustGrid::whichcell, CartesianDustGrid.cpp:109-118
    int i = locateFail(g.xv, x, g.Nx + 1);
    int j = locateFail(g.yv, y, g.Ny + 1);
    int k = locateFail(g.zv, z, g.Nz + 1);
    if (i < 0 || j < 0 || k < 0) return -1;
    return k + g.Nz * j + g.Nz * g.Ny * i;
}

// 8-byte load from the shared window (the staged Cartesian borders): an explicit LDS instead of a generic load
__device__ __forceinline__ double ldsF64(unsigned addr)
{ double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }

// One crossing at a time: CartesianDustGrid::path (CartesianDustGrid.cpp:136-283) as a state machine, so that a
// warp can refill finished lanes with new rays instead of waiting for its longest path.
//
// Per-ray invariants are hoisted out of the crossing: the walker tracks, per axis, the BYTE OFFSET of the border the
// ray leaves through (xv[i+1] for k >= 0, xv[i] for k < 0, :236-238), the signed offset / cell-number increments
// of a crossing, and the offset at which the ray leaves the grid.  A crossing is then three shared-memory loads,
// three invariant-divisor divisions, the reference's three-way exit test as selects, and predicated increments.
#ifndef SKG_CART_REGBORDERS
#define SKG_CART_REGBORDERS 0
#endif
// REGB: keep the three exit borders in registers and re-read only the axis that was crossed (fewer shared-memory
// wavefronts: for kernels whose load/store pipe is the bottleneck); otherwise read all three every crossing (the
// loads are issued first and overlap, shortest dependency chain).
// TINYSEL: how direction components with |k| <= 1e-15 are handled: as unconditional selects (peel-off rays towards
// an observer at azimuth 0/90/... have an exactly zero component in EVERY lane) or as a branch that random rays
// essentially never take.
template<bool REGB, bool TINYSEL> struct CartWalkerT
{
    static constexpr int kStepUnroll = 4;       // crossings of a batch unrolled in the scheduler (wavefront.cuh): the step is ~100 instructions
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;       // 1/k per axis (see divInvariant)
    double xE, yE, zE;          // the exit borders themselves: only the axis that was crossed is re-read
    int ox, oy, oz;             // byte offsets of the exit borders in xv / yv / zv
    int stx, sty, stz;          // +-8: offset increment of a crossing along each axis
    int dmx, dmy, dmz;          // cell number increments (m = k + Nz*j + Nz*Ny*i, :326-329)
    int m, tiny;                // tiny: bit a set when |k_a| <= 1e-15 (that axis is never crossed, :240-242)
    bool alive;

    // entry part, :151-230.  Returns false when the ray misses the grid (the reference clears the path);
    // otherwise `en` holds the up to three "outside" segments (m = -1) that precede the first cell.
    __device__ __forceinline__ bool start(const CartGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        const double* xv = g.xv; const double* yv = g.yv; const double* zv = g.zv;
        const int Nx = g.Nx, Ny = g.Ny, Nz = g.Nz;
        const double xmin = g.ext[0], xmax = g.ext[1], ymin = g.ext[2], ymax = g.ext[3], zmin = g.ext[4], zmax = g.ext[5];
        double ds;
        if (x < xmin)
        {
            if (kx <= 0.0) return false;
            ds = (xmin - x) / kx; en.ds[en.n++] = ds;
            x = xmin + 1e-8 * (xv[1] - xv[0]); y += ky * ds; z += kz * ds;
        }
        else if (x > xmax)
        {
            if (kx >= 0.0) return false;
            ds = (xmax - x) / kx; en.ds[en.n++] = ds;
            x = xmax - 1e-8 * (xv[Nx] - xv[Nx - 1]); y += ky * ds; z += kz * ds;
        }
        if (y < ymin)
        {
            if (ky <= 0.0) return false;
            ds = (ymin - y) / ky; en.ds[en.n++] = ds;
            x += kx * ds; y = ymin + 1e-8 * (yv[1] - yv[0]); z += kz * ds;
        }
        else if (y > ymax)
        {
            if (ky >= 0.0) return false;
            ds = (ymax - y) / ky; en.ds[en.n++] = ds;
            x += kx * ds; y = ymax - 1e-8 * (yv[Ny] - yv[Ny - 1]); z += kz * ds;
        }
        if (z < zmin)
        {
            if (kz <= 0.0) return false;
            ds = (zmin - z) / kz; en.ds[en.n++] = ds;
            x += kx * ds; y += ky * ds; z = zmin + 1e-8 * (zv[1] - zv[0]);
        }
        else if (z > zmax)
        {
            if (kz >= 0.0) return false;
            ds = (zmax - z) / kz; en.ds[en.n++] = ds;
            x += kx * ds; y += ky * ds; z = zmax - 1e-8 * (zv[Nz] - zv[Nz - 1]);
        }
        if (x < xmin || x > xmax || y < ymin || y > ymax || z < zmin || z > zmax) return false;     // :224
        const int i = locateClip(xv, x, Nx + 1);
        const int j = locateClip(yv, y, Ny + 1);
        const int k = locateClip(zv, z, Nz + 1);
        m = k + Nz * j + Nz * Ny * i;
        const bool nx = kx < 0.0, ny = ky < 0.0, nz = kz < 0.0;
        ox = 8 * (i + (nx ? 0 : 1)); oy = 8 * (j + (ny ? 0 : 1)); oz = 8 * (k + (nz ? 0 : 1));
        stx = nx ? -8 : 8; sty = ny ? -8 : 8; stz = nz ? -8 : 8;
        dmx = nx ? -Ny * Nz : Ny * Nz; dmy = ny ? -Nz : Nz; dmz = nz ? -1 : 1;
        tiny = (fabs(kx) > 1e-15 ? 0 : 1) | (fabs(ky) > 1e-15 ? 0 : 2) | (fabs(kz) > 1e-15 ? 0 : 4);
        rkx = 1.0 / kx; rky = 1.0 / ky; rkz = 1.0 / kz;
        if (REGB) { xE = ldsF64(g.sx + ox); yE = ldsF64(g.sy + oy); zE = ldsF64(g.sz + oz); }
        alive = true;
        return true;
    }

    // one pass of the loop :234-282.  Returns true when segment (mseg, ds) is to be added (addSegment drops ds <= 0).
    __device__ __forceinline__ bool step(const CartGrid& g, Counters*, int& mseg, double& ds)
    {
        // the borders are always staged in shared memory by the kernels that walk (stageCart)
        if (!REGB) { xE = ldsF64(g.sx + ox); yE = ldsF64(g.sy + oy); zE = ldsF64(g.sz + oz); }
        double dsx = divInvariant(xE - x, kx, rkx);
        double dsy = divInvariant(yE - y, ky, rky);
        double dsz = divInvariant(zE - z, kz, rkz);
        // direction components with |k| <= 1e-15 never cross a border: ds = DBL_MAX (:240-242)
        if (TINYSEL)
        {
            dsx = (tiny & 1) ? SKG_DBL_MAX : dsx; dsy = (tiny & 2) ? SKG_DBL_MAX : dsy; dsz = (tiny & 4) ? SKG_DBL_MAX : dsz;
        }
        else if (tiny)
        {
            // written as a loop over the set bits so that it stays a branch instead of predicated selects
            int t = tiny;
            do
            {
                const int bit = t & -t;
                if (bit == 1) dsx = SKG_DBL_MAX; else if (bit == 2) dsy = SKG_DBL_MAX; else dsz = SKG_DBL_MAX;
                t ^= bit;
            } while (t);
        }
        mseg = m;
        // the reference's three branches (X if dsx<=dsy&&dsx<=dsz, Y if dsy<dsx&&dsy<=dsz, Z if dsz<dsx&&dsz<dsy) as selects,
        // so that the lanes of a warp do not diverge on the exit face: the hit coordinate snaps to the face, the other two
        // advance by k*ds -- identical values, branch free
        const bool bx = dsx <= dsy && dsx <= dsz;
        const bool by = !bx && dsy < dsx && dsy <= dsz;
        const bool bz = !bx && !by && dsz < dsx && dsz < dsy;
        if (!(bx || by || bz)) { alive = false; return false; }     // unreachable for finite input (the reference would spin forever)
        ds = bx ? dsx : (by ? dsy : dsz);
        const double xa = x + kx * ds, ya = y + ky * ds, za = z + kz * ds;
        x = bx ? xE : xa; y = by ? yE : ya; z = bz ? zE : za;
        // only the crossed axis gets a new exit border (one shared-memory read per crossing, a third of the lanes
        // per table: few bank conflicts); the staged arrays carry one pad element on either side, so that the read
        // is harmless when the ray has just left the grid
        if (bx) { ox += stx; m += dmx; if (REGB) xE = ldsF64(g.sx + ox); }
        if (by) { oy += sty; m += dmy; if (REGB) yE = ldsF64(g.sy + oy); }
        if (bz) { oz += stz; m += dmz; if (REGB) zE = ldsF64(g.sz + oz); }
        // left the grid: the offset stepped below 0 (wraps) or beyond the last border
        alive = (unsigned)ox <= 8u * g.Nx && (unsigned)oy <= 8u * g.Ny && (unsigned)oz <= 8u * g.Nz;
        return ds > 0;
    }
};

// ---------------------------------------------------------------------------------------------------
// DustGridPath::moveInside, DustGridPath.cpp:57-150.  box = xmin,ymin,zmin,xmax,ymax,zmax.
// Returns false for the reference's OUTSIDE position.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool moveInside(const double* box, double eps, double& rx, double& ry, double& rz,
                                           double kx, double ky, double kz, Entry& en)
{
    const double xmin = box[0], ymin = box[1], zmin = box[2], xmax = box[3], ymax = box[4], zmax = box[5];
    en.n = 0;
    if (rx <= xmin)
    {
        if (kx <= 0.0) return false;
        double ds = (xmin - rx) / kx; en.ds[en.n++] = ds;
        rx = xmin + eps; ry += ky * ds; rz += kz * ds;
    }
    else if (rx >= xmax)
    {
        if (kx >= 0.0) return false;
        double ds = (xmax - rx) / kx; en.ds[en.n++] = ds;
        rx = xmax - eps; ry += ky * ds; rz += kz * ds;
    }
    if (ry <= ymin)
    {
        if (ky <= 0.0) return false;
        double ds = (ymin - ry) / ky; en.ds[en.n++] = ds;
        rx += kx * ds; ry = ymin + eps; rz += kz * ds;
    }
    else if (ry >= ymax)
    {
        if (ky >= 0.0) return false;
        double ds = (ymax - ry) / ky; en.ds[en.n++] = ds;
        rx += kx * ds; ry = ymax - eps; rz += kz * ds;
    }
    if (rz <= zmin)
    {
        if (kz <= 0.0) return false;
        double ds = (zmin - rz) / kz; en.ds[en.n++] = ds;
        rx += kx * ds; ry += ky * ds; rz = zmin + eps;
    }
    else if (rz >= zmax)
    {
        if (kz >= 0.0) return false;
        double ds = (zmax - rz) / kz; en.ds[en.n++] = ds;
        rx += kx * ds; ry += ky * ds; rz = zmax - eps;
    }
    return true;
}

// Box::contains (closed on all faces), Box.hpp:94-95
__device__ __forceinline__ bool boxContains(const double* b, double x, double y, double z)
{
    return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5];
}

// ---------------------------------------------------------------------------------------------------
// Tree (octree / binary tree)
// ---------------------------------------------------------------------------------------------------
// TreeNode::whichnode(Vec), TreeNode.cpp:70-80, starting from the root
__device__ __forceinline__ int treeWhichNode(const TreeGrid& g, double x, double y, double z)
{
    if (!boxContains(g.box, x, y, z)) return -1;
    int node = 0;
    if (g.lookupG > 1)
    {
        // start below the root when the lattice cell's node strictly contains the point: the descent from the root
        // necessarily passes through that node (all comparisons against its ancestors' split planes agree)
        const int G = g.lookupG;
        const int i = max(0, min(G - 1, (int)((x - g.box[0]) * g.lookupInv[0])));
        const int j = max(0, min(G - 1, (int)((y - g.box[1]) * g.lookupInv[1])));
        const int k = max(0, min(G - 1, (int)((z - g.box[2]) * g.lookupInv[2])));
        const int cand = __ldg(g.lookup + ((size_t)i * G + j) * G + k);
        const double* b = g.box + 6 * (size_t)cand;
        if (x > __ldg(b) && x < __ldg(b + 3) && y > __ldg(b + 1) && y < __ldg(b + 4) && z > __ldg(b + 2) && z < __ldg(b + 5)) node = cand;
    }
    int c0;
    while ((c0 = __ldg(g.child0 + node)) >= 0)
    {
        const double* cb = g.box + 6 * (size_t)c0;
        if (g.kind == 0)
        {
            // OctTreeNode::child, OctTreeNode.cpp:184-189
            int l = (x < cb[3] ? 0 : 1) + (y < cb[4] ? 0 : 2) + (z < cb[5] ? 0 : 4);
            node = c0 + l;
        }
        else
        {
            // BinTreeNode::child, BinTreeNode.cpp:326-335
            int d = __ldg(g.dir + node);
            double v = d == 0 ? x : (d == 1 ? y : z);
            node = (v < cb[3 + d]) ? c0 : c0 + 1;
        }
    }
    return node;
}

__device__ __forceinline__ double nextAfterAlong(double v, double k)
{
    return nextafter(v, (k < 0.0) ? -SKG_DBL_MAX : SKG_DBL_MAX);
}

// TreeDustGrid::path (TreeDustGrid.cpp:390-662) one crossing at a time
struct TreeWalker
{
    static constexpr int kStepUnroll = 1;       // large step body: a plain loop (unrolling it costs more in instruction fetch than it gains)
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;
    double bx[6];               // box of the current node, carried over from the neighbour test that selected it
    int nb[7];                  // neighbour list offsets of the current node's six walls (Neighbor search)
    int node, cellv;
    bool alive;

    // everything a crossing needs from the node tables, fetched as soon as the node is known
    __device__ __forceinline__ void loadNode(const TreeGrid& g, bool withBox)
    {
        if (withBox) { const double* b = g.box + 6 * (size_t)node; for (int c = 0; c < 6; c++) bx[c] = __ldg(b + c); }
        cellv = __ldg(g.cell + node);
        if (g.search == 1)
        {
            const int* s = g.nbrStart + 6 * (size_t)node; for (int w = 0; w < 7; w++) nb[w] = __ldg(s + w);
        }
    }
    // one expanded neighbour record: 96 bytes as three 256-bit reads
    static __device__ __forceinline__ void loadRec(const TreeNbrRec* rec, double (&w)[12])
    {
        const double* rp = reinterpret_cast<const double*>(rec);
#pragma unroll
        for (int u = 0; u < 3; u++)
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(w[4 * u]), "=d"(w[4 * u + 1]), "=d"(w[4 * u + 2]), "=d"(w[4 * u + 3]) : "l"(rp + 4 * u));
    }
    // continue from the neighbour whose record has been read
    __device__ __forceinline__ void adopt(const double (&w)[12])
    {
        for (int c = 0; c < 6; c++) bx[c] = w[c];
        const long long w6 = __double_as_longlong(w[6]), w7 = __double_as_longlong(w[7]), w8 = __double_as_longlong(w[8]),
                        w9 = __double_as_longlong(w[9]), w10 = __double_as_longlong(w[10]);
        node = (int)(w6 & 0xffffffffll); cellv = (int)(w6 >> 32);
        nb[0] = (int)(w7 & 0xffffffffll); nb[1] = (int)(w7 >> 32); nb[2] = (int)(w8 & 0xffffffffll); nb[3] = (int)(w8 >> 32);
        nb[4] = (int)(w9 & 0xffffffffll); nb[5] = (int)(w9 >> 32); nb[6] = (int)(w10 & 0xffffffffll);
    }

    __device__ __forceinline__ bool start(const TreeGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        if (!moveInside(g.box, g.eps, x, y, z, kx, ky, kz, en)) return false;
        node = treeWhichNode(g, x, y, z);
        if (node < 0) return false;
        loadNode(g, true);
        rkx = 1.0 / kx; rky = 1.0 / ky; rkz = 1.0 / kz;
        alive = true;
        return true;
    }

    __device__ __forceinline__ bool step(const TreeGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        const bool nx = kx < 0.0, ny = ky < 0.0, nz = kz < 0.0;
        const double xnext = nx ? bx[0] : bx[3];
        const double ynext = ny ? bx[1] : bx[4];
        const double znext = nz ? bx[2] : bx[5];
        const double dsx = (fabs(kx) > 1e-15) ? divInvariant(xnext - x, kx, rkx) : SKG_DBL_MAX;
        const double dsy = (fabs(ky) > 1e-15) ? divInvariant(ynext - y, ky, rky) : SKG_DBL_MAX;
        const double dsz = (fabs(kz) > 1e-15) ? divInvariant(znext - z, kz, rkz) : SKG_DBL_MAX;
        mseg = cellv;
        if (g.search != 2)
        {
            // TopDown (TreeDustGrid.cpp:412-456) and Neighbor (:460-521)
            const double eps = g.eps;
            int wall;
            if (dsx <= dsy && dsx <= dsz) { ds = dsx; wall = nx ? 0 : 1; }
            else if (dsy <= dsx && dsy <= dsz) { ds = dsy; wall = ny ? 2 : 3; }
            else { ds = dsz; wall = nz ? 4 : 5; }
            x += (ds + eps) * kx;
            y += (ds + eps) * ky;
            z += (ds + eps) * kz;

            const int oldnode = node;
            bool haveBox = false, haveAll = false;
            if (g.search == 1)
            {
                // TreeNode::whichnode(wall, r), TreeNode.cpp:84-93: first neighbour whose closed box contains r
                const int beg = wall == 0 ? nb[0] : wall == 1 ? nb[1] : wall == 2 ? nb[2] : wall == 3 ? nb[3] : wall == 4 ? nb[4] : nb[5];
                const int end = wall == 0 ? nb[1] : wall == 1 ? nb[2] : wall == 2 ? nb[3] : wall == 3 ? nb[4] : wall == 4 ? nb[5] : nb[6];
                node = -1;
                if (g.nbrRec)
                {
                    // expanded records: box, cell and the neighbour offsets of the candidate arrive in one 96-byte read.
                    int q = beg;
                    for (; q < end; q++)
                    {
                        double w[12]; loadRec(g.nbrRec + q, w);
                        if (x >= w[0] && x <= w[3] && y >= w[1] && y <= w[4] && z >= w[2] && z <= w[5]) { adopt(w); haveAll = true; break; }
                    }
                    if (!haveAll) node = -1;
                }
                else for (int q = beg; q < end; q++)
                {
                    const int cand = __ldg(g.nbrIds + q);
                    const double* cb = g.box + 6 * (size_t)cand;
                    const double c0 = __ldg(cb), c1 = __ldg(cb + 1), c2 = __ldg(cb + 2), c3 = __ldg(cb + 3), c4 = __ldg(cb + 4), c5 = __ldg(cb + 5);
                    if (x >= c0 && x <= c3 && y >= c1 && y <= c4 && z >= c2 && z <= c5)
                    { node = cand; bx[0] = c0; bx[1] = c1; bx[2] = c2; bx[3] = c3; bx[4] = c4; bx[5] = c5; haveBox = true; break; }
                }
                if (node < 0) node = treeWhichNode(g, x, y, z);
            }
            else node = treeWhichNode(g, x, y, z);

            if (node == oldnode)
            {
                atomicAdd(&ctr->stuckEscaped, 1ull);
                x = nextAfterAlong(x, kx); y = nextAfterAlong(y, ky); z = nextAfterAlong(z, kz);
                node = treeWhichNode(g, x, y, z);
                haveBox = false; haveAll = false;
                if (node == oldnode) { atomicAdd(&ctr->stuckTerminated, 1ull); node = -1; }
            }
            if (node < 0) alive = false;
            else if (!haveAll) loadNode(g, !haveBox);
            return ds > 0;
        }

        // Bookkeeping (octree only), TreeDustGrid.cpp:527-659
        int l = node;
        if (dsx <= dsy && dsx <= dsz)
        {
            ds = dsx;
            x = xnext; y += ky * dsx; z += kz * dsx;
            while (true)
            {
                int oct = ((l - 1) % 8) + 1;
                bool place = nx ? (oct % 2 == 1) : (oct % 2 == 0);
                if (!place) break;
                l = __ldg(g.parent + l);
                if (l == 0) { alive = false; return ds > 0; }
            }
            l += nx ? -1 : 1;
            while (__ldg(g.cell + l) == -1)
            {
                int c0 = __ldg(g.child0 + l);
                const double* cb = g.box + 6 * (size_t)c0;
                double yM = cb[4], zM = cb[5];
                if (nx) l = (y <= yM) ? ((z <= zM) ? c0 + 1 : c0 + 5) : ((z <= zM) ? c0 + 3 : c0 + 7);
                else    l = (y <= yM) ? ((z <= zM) ? c0 + 0 : c0 + 4) : ((z <= zM) ? c0 + 2 : c0 + 6);
            }
        }
        else if (dsy < dsx && dsy <= dsz)
        {
            ds = dsy;
            x += kx * dsy; y = ynext; z += kz * dsy;
            while (true)
            {
                bool place = ny ? ((l - 1) % 4 < 2) : ((l - 1) % 4 > 1);
                if (!place) break;
                l = __ldg(g.parent + l);
                if (l == 0) { alive = false; return ds > 0; }
            }
            l += ny ? -2 : 2;
            while (__ldg(g.cell + l) == -1)
            {
                int c0 = __ldg(g.child0 + l);
                const double* cb = g.box + 6 * (size_t)c0;
                double xM = cb[3], zM = cb[5];
                if (ny) l = (x <= xM) ? ((z <= zM) ? c0 + 2 : c0 + 6) : ((z <= zM) ? c0 + 3 : c0 + 7);
                else    l = (x <= xM) ? ((z <= zM) ? c0 + 0 : c0 + 4) : ((z <= zM) ? c0 + 1 : c0 + 5);
            }
        }
        else if (dsz < dsx && dsz < dsy)
        {
            ds = dsz;
            x += kx * dsz; y += ky * dsz; z = znext;
            while (true)
            {
                int oct = ((l - 1) % 8) + 1;
                bool place = nz ? (oct < 5) : (oct > 4);
                if (!place) break;
                l = __ldg(g.parent + l);
                if (l == 0) { alive = false; return ds > 0; }
            }
            l += nz ? -4 : 4;
            while (__ldg(g.cell + l) == -1)
            {
                int c0 = __ldg(g.child0 + l);
                const double* cb = g.box + 6 * (size_t)c0;
                double xM = cb[3], yM = cb[4];
                if (nz) l = (x <= xM) ? ((y <= yM) ? c0 + 4 : c0 + 6) : ((y <= yM) ? c0 + 5 : c0 + 7);
                else    l = (x <= xM) ? ((y <= yM) ? c0 + 0 : c0 + 2) : ((y <= yM) ? c0 + 1 : c0 + 3);
            }
        }
        else { alive = false; return false; }
        node = l;
        loadNode(g, true);
        return ds > 0;
    }
};

// ---------------------------------------------------------------------------------------------------
// Adaptive mesh
// ---------------------------------------------------------------------------------------------------
// Box::cellindices for one axis, Box.hpp:134-139
__device__ __forceinline__ int cellIndex1(double v, double vmin, double vmax, int n)
{
    double q = n * (v - vmin) / (vmax - vmin);
    int i = (q >= 2147483648.0 || q <= -2147483649.0 || q != q) ? INT_MIN : (int)q;   // x86 cvttsd2si behaviour
    return max(0, min(n - 1, i));
}

// AdaptiveMeshNode::whichnode(Vec) from the root, AdaptiveMeshNode.cpp:132-142 (+ child :109-128).
// Returns -1 when outside, -2 when the reference would throw "Can't locate the appropriate child node".
__device__ __forceinline__ int ameshWhichNode(const AMeshGrid& g, double x, double y, double z)
{
    if (!boxContains(g.box, x, y, z)) return -1;
    int node = 0;
    int c0;
    while ((c0 = __ldg(g.child0 + node)) >= 0)
    {
        const double* b = g.box + 6 * (size_t)node;
        int Nx = __ldg(g.nxyz + 3 * (size_t)node), Ny = __ldg(g.nxyz + 3 * (size_t)node + 1), Nz = __ldg(g.nxyz + 3 * (size_t)node + 2);
        int i = cellIndex1(x, b[0], b[3], Nx);
        int j = cellIndex1(y, b[1], b[4], Ny);
        int k = cellIndex1(z, b[2], b[5], Nz);
        int child = c0 + (k * Ny + j) * Nx + i;
        const double* cb = g.box + 6 * (size_t)child;
        if (!boxContains(cb, x, y, z))
        {
            if (x < cb[0]) i--; else if (x > cb[3]) i++;
            if (y < cb[1]) j--; else if (y > cb[4]) j++;
            if (z < cb[2]) k--; else if (z > cb[5]) k++;
            if (i < 0 || i >= Nx || j < 0 || j >= Ny || k < 0 || k >= Nz) return -2;
            child = c0 + (k * Ny + j) * Nx + i;
            if (!boxContains(g.box + 6 * (size_t)child, x, y, z)) return -2;
        }
        node = child;
    }
    return node;
}

// AdaptiveMesh::path (AdaptiveMesh.cpp:297-367) one crossing at a time
struct AMeshWalker
{
    static constexpr int kStepUnroll = 1;
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;
    double bx[6];               // box of the current node (carried over from the neighbour test that selected it)
    int wn[6];                  // the node beyond each wall of the current leaf
    int node, cellv;
    bool alive;

    __device__ __forceinline__ void loadNode(const AMeshGrid& g, bool withBox)
    {
        if (withBox) { const double* b = g.box + 6 * (size_t)node; for (int c = 0; c < 6; c++) bx[c] = __ldg(b + c); }
        cellv = __ldg(g.cell + node);
        const int* w = g.wallNbr + 6 * (size_t)node;
        for (int c = 0; c < 6; c++) wn[c] = __ldg(w + c);
    }

    __device__ __forceinline__ bool start(const AMeshGrid& g, Counters* ctr, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        if (!moveInside(g.box, g.eps, x, y, z, kx, ky, kz, en)) return false;
        node = ameshWhichNode(g, x, y, z);
        if (node < 0) { if (node == -2) atomicAdd(&ctr->errors, 1ull); return false; }
        loadNode(g, true);
        rkx = 1.0 / kx; rky = 1.0 / ky; rkz = 1.0 / kz;
        alive = true;
        return true;
    }

    __device__ __forceinline__ bool step(const AMeshGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        const bool nx = kx < 0.0, ny = ky < 0.0, nz = kz < 0.0;
        const double eps = g.eps;
        const double xnext = nx ? bx[0] : bx[3];
        const double ynext = ny ? bx[1] : bx[4];
        const double znext = nz ? bx[2] : bx[5];
        const double dsx = (fabs(kx) > 1e-15) ? divInvariant(xnext - x, kx, rkx) : SKG_DBL_MAX;
        const double dsy = (fabs(ky) > 1e-15) ? divInvariant(ynext - y, ky, rky) : SKG_DBL_MAX;
        const double dsz = (fabs(kz) > 1e-15) ? divInvariant(znext - z, kz, rkz) : SKG_DBL_MAX;
        int wall;
        if (dsx <= dsy && dsx <= dsz) { ds = dsx; wall = nx ? 0 : 1; }
        else if (dsy <= dsx && dsy <= dsz) { ds = dsy; wall = ny ? 2 : 3; }
        else { ds = dsz; wall = nz ? 4 : 5; }
        mseg = cellv;
        // r += (ds+eps)*k   (Vec operator*(double,Vec), Vec.hpp)
        x += (ds + eps) * kx;
        y += (ds + eps) * ky;
        z += (ds + eps) * kz;

        const int oldnode = node;
        const int cand = wall == 0 ? wn[0] : wall == 1 ? wn[1] : wall == 2 ? wn[2] : wall == 3 ? wn[3] : wall == 4 ? wn[4] : wn[5];
        bool haveBox = false;
        node = -3;
        if (cand >= 0)
        {
            const double* cb = g.box + 6 * (size_t)cand;
            const double c0 = __ldg(cb), c1 = __ldg(cb + 1), c2 = __ldg(cb + 2), c3 = __ldg(cb + 3), c4 = __ldg(cb + 4), c5 = __ldg(cb + 5);
            if (x >= c0 && x <= c3 && y >= c1 && y <= c4 && z >= c2 && z <= c5)
            { node = cand; bx[0] = c0; bx[1] = c1; bx[2] = c2; bx[3] = c3; bx[4] = c4; bx[5] = c5; haveBox = true; }
        }
        if (node == -3) node = ameshWhichNode(g, x, y, z);
        if (node == -2) { atomicAdd(&ctr->errors, 1ull); alive = false; return ds > 0; }

        if (node == oldnode)
        {
            atomicAdd(&ctr->stuckEscaped, 1ull);
            x = nextAfterAlong(x, kx); y = nextAfterAlong(y, ky); z = nextAfterAlong(z, kz);
            node = ameshWhichNode(g, x, y, z);
            haveBox = false;
            if (node == -2) { atomicAdd(&ctr->errors, 1ull); alive = false; return ds > 0; }
            if (node == oldnode) { atomicAdd(&ctr->stuckTerminated, 1ull); node = -1; }
        }
        if (node < 0) alive = false;
        else loadNode(g, !haveBox);
        return ds > 0;
    }
};

// ---------------------------------------------------------------------------------------------------
// Voronoi
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double voroSD(const VoroGrid& g, int m, double x, double y, double z)
{
    // VoronoiCell::squaredDistanceTo: (r-_r).norm2(), Vec.hpp
    const double* p = g.particles + 3 * (size_t)m;
    double dx = x - p[0], dy = y - p[1], dz = z - p[2];
    return dx * dx + dy * dy + dz * dz;
}

// lessthan(p1, p2, axis), VoronoiMesh.cpp:77-105
__device__ __forceinline__ bool voroLess(double x1, double y1, double z1, const double* p2, int axis)
{
    double a1, a2, b1, b2, c1, c2;
    if (axis == 0) { a1 = x1; a2 = p2[0]; b1 = y1; b2 = p2[1]; c1 = z1; c2 = p2[2]; }
    else if (axis == 1) { a1 = y1; a2 = p2[1]; b1 = z1; b2 = p2[2]; c1 = x1; c2 = p2[0]; }
    else if (axis == 2) { a1 = z1; a2 = p2[2]; b1 = x1; b2 = p2[0]; c1 = y1; c2 = p2[1]; }
    else return false;
    if (a1 < a2) return true;
    if (a1 > a2) return false;
    if (b1 < b2) return true;
    if (b1 > b2) return false;
    if (c1 < c2) return true;
    return false;
}

// Node::nearest (VoronoiMesh.cpp:180-225) made iterative: an explicit stack holds the frames of the
// reference's recursion (each frame = one nearest() invocation on a subtree).
#define SKG_KD_STACK 40
static __device__ int voroKdNearest(const VoroGrid& g, int root, double x, double y, double z)
{
    int fRoot[SKG_KD_STACK], fCur[SKG_KD_STACK], fBest[SKG_KD_STACK];
    double fBestSD[SKG_KD_STACK];
    int sp = 0;
    int result = -1;            // value "returned" by the most recently finished frame
    bool entering = true;       // true: start a new frame on `root`; false: resume frame sp-1 after a return
    while (true)
    {
        int cur, best; double bestSD; int froot;
        if (entering)
        {
            if (sp >= SKG_KD_STACK) return -1;
            froot = root;
            cur = root;
            while (true)
            {
                int mm = __ldg(g.kdM + cur), axis = __ldg(g.kdAxis + cur);
                bool less = voroLess(x, y, z, g.particles + 3 * (size_t)mm, axis % 3);
                int child = less ? __ldg(g.kdLeft + cur) : __ldg(g.kdRight + cur);
                if (child < 0) break;
                cur = child;
            }
            best = cur; bestSD = voroSD(g, __ldg(g.kdM + best), x, y, z);
        }
        else
        {
            sp--;
            froot = fRoot[sp]; cur = fCur[sp]; best = fBest[sp]; bestSD = fBestSD[sp];
            // combine with the result of the recursive call on the other child
            double oSD = voroSD(g, __ldg(g.kdM + result), x, y, z);
            if (oSD < bestSD) { best = result; bestSD = oSD; }
            // move up
            if (cur == froot) { result = best; if (sp == 0) return __ldg(g.kdM + result); entering = false; continue; }
            cur = __ldg(g.kdUp + cur);
        }
        // climbing loop
        bool pushed = false;
        while (true)
        {
            int mm = __ldg(g.kdM + cur), axis = __ldg(g.kdAxis + cur) % 3;
            double curSD = voroSD(g, mm, x, y, z);
            if (curSD < bestSD) { best = cur; bestSD = curSD; }
            const double* p = g.particles + 3 * (size_t)mm;
            double d = (axis == 0) ? (p[0] - x) : (axis == 1 ? (p[1] - y) : (p[2] - z));
            double splitSD = d * d;
            if (splitSD < bestSD)
            {
                bool less = voroLess(x, y, z, p, axis);
                int other = less ? __ldg(g.kdRight + cur) : __ldg(g.kdLeft + cur);
                if (other >= 0)
                {
                    fRoot[sp] = froot; fCur[sp] = cur; fBest[sp] = best; fBestSD[sp] = bestSD; sp++;
                    root = other; entering = true; pushed = true;
                    break;
                }
            }
            if (cur == froot) break;
            cur = __ldg(g.kdUp + cur);
        }
        if (pushed) continue;
        result = best;
        if (sp == 0) return __ldg(g.kdM + result);
        entering = false;
    }
}

// VoronoiMesh::cellIndex, VoronoiMesh.cpp:512-541
__device__ __forceinline__ int voroCellIndex(const VoroGrid& g, double x, double y, double z)
{
    if (!boxContains(g.ext, x, y, z)) return -1;
    int nb = g.nb;
    int i = cellIndex1(x, g.ext[0], g.ext[3], nb);
    int j = cellIndex1(y, g.ext[1], g.ext[4], nb);
    int k = cellIndex1(z, g.ext[2], g.ext[5], nb);
    size_t b = (size_t)i * nb * nb + (size_t)j * nb + k;
    int tree = __ldg(g.blkTree + b);
    if (tree >= 0) return voroKdNearest(g, tree, x, y, z);
    int beg = __ldg(g.blkStart + b), end = __ldg(g.blkStart + b + 1);
    int m = -1;
    double mdist = SKG_DBL_MAX;
    for (int q = beg; q < end; q++)
    {
        int id = __ldg(g.blkIds + q);
        double idist = voroSD(g, id, x, y, z);
        if (idist < mdist) { m = id; mdist = idist; }
    }
    return m;
}

// VoronoiMesh::path (VoronoiMesh.cpp:749-844) one crossing at a time
struct VoroWalker
{
    static constexpr int kStepUnroll = 1;
    double x, y, z, kx, ky, kz;
    int mr;
    int guard;
    bool alive;

    __device__ __forceinline__ bool start(const VoroGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en)
    {
        alive = false; en.n = 0; guard = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        if (!moveInside(g.ext, g.eps, x, y, z, kx, ky, kz, en)) return false;
        mr = voroCellIndex(g, x, y, z);
        if (mr < 0) return false;
        alive = true;
        return true;
    }

    __device__ __forceinline__ bool step(const VoroGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        const double eps = g.eps;
        const double* pr = g.particles + 3 * (size_t)mr;
        const double prx = pr[0], pry = pr[1], prz = pr[2];
        double sq = SKG_DBL_MAX;
        const int NO_INDEX = -99;
        int mq = NO_INDEX;
        const int beg = __ldg(g.nbrStart + mr), end = __ldg(g.nbrStart + mr + 1);
        // the neighbour loop of VoronoiMesh.cpp:777-828, four neighbours at a time: ids and particle positions of a
        // group are fetched together (twelve independent loads in flight), then evaluated in list order so that the
        // first smallest intersection still wins
        for (int q0 = beg; q0 < end; q0 += 4)
        {
            int mi[4]; double px[4], py[4], pz[4];
#pragma unroll
            for (int u = 0; u < 4; u++) mi[u] = (q0 + u < end) ? __ldg(g.nbrIds + q0 + u) : -99;
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const double* pi = g.particles + 3 * (size_t)max(mi[u], 0);
                px[u] = __ldg(pi); py[u] = __ldg(pi + 1); pz[u] = __ldg(pi + 2);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                if (q0 + u >= end) break;
                double si = 0;
                if (mi[u] >= 0)
                {
                    const double pix = px[u], piy = py[u], piz = pz[u];
                    double nxv = pix - prx, nyv = piy - pry, nzv = piz - prz;        // n = pi - pr
                    double ndotk = nxv * kx + nyv * ky + nzv * kz;                   // Vec::dot(n,bfk)
                    if (ndotk > 0)
                    {
                        double qx = 0.5 * (pix + prx), qy = 0.5 * (piy + pry), qz = 0.5 * (piz + prz);   // p = 0.5*(pi+pr)
                        si = (nxv * (qx - x) + nyv * (qy - y) + nzv * (qz - z)) / ndotk;                // dot(n,p-r)/ndotk
                    }
                }
                else
                {
                    switch (mi[u])
                    {
                    case -1: si = (g.ext[0] - x) / kx; break;
                    case -2: si = (g.ext[3] - x) / kx; break;
                    case -3: si = (g.ext[1] - y) / ky; break;
                    case -4: si = (g.ext[4] - y) / ky; break;
                    case -5: si = (g.ext[2] - z) / kz; break;
                    case -6: si = (g.ext[5] - z) / kz; break;
                    default: atomicAdd(&ctr->errors, 1ull); alive = false; return false;
                    }
                }
                if (si > 0 && si < sq) { sq = si; mq = mi[u]; }
            }
        }
        if (mq == NO_INDEX)
        {
            // r += bfk*_eps  (Vec operator*(Vec,double))
            x += kx * eps; y += ky * eps; z += kz * eps;
            mr = voroCellIndex(g, x, y, z);
            if (++guard > 1000000) { atomicAdd(&ctr->errors, 1ull); mr = -1; }
            if (mr < 0) alive = false;
            return false;
        }
        mseg = mr; ds = sq;                     // sq > 0 by construction
        x += (sq + eps) * kx; y += (sq + eps) * ky; z += (sq + eps) * kz;
        mr = mq;
        if (mr < 0) alive = false;
        return true;
    }
};

}   // namespace skg
