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
// pulls the line holding *p into L1 (no register, no scoreboard)
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ double ldsF64(unsigned addr)
{ double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }

// entry part of CartesianDustGrid::path, :151-230: moves a ray that starts outside the grid to its boundary.  Returns false
// when the ray misses the grid (the reference clears the path); otherwise `en` holds the up to three "outside" segments
// (m = -1) that precede the first cell, (x, y, z) the position inside the grid and (i, j, k) the indices of the first cell.
__device__ __forceinline__ bool cartEnter(const CartGrid& g, double& x, double& y, double& z, double kx, double ky, double kz,
                                          Entry& en, int& i, int& j, int& k)
{
    en.n = 0;
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
    i = locateClip(xv, x, Nx + 1);
    j = locateClip(yv, y, Ny + 1);
    k = locateClip(zv, z, Nz + 1);
    return true;
}

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
template<bool REGB, bool TINYSEL, bool AHEAD = false> struct CartWalkerT
{
    static constexpr bool kPredicated = false;
    static constexpr int kStepUnroll = 4;       // crossings of a batch unrolled in the scheduler (wavefront.cuh): the step is ~100 instructions
    static constexpr bool kSplitStep = false;
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;       // 1/k per axis (see divInvariant)
    double xE, yE, zE;          // the exit borders themselves: only the axis that was crossed is re-read
    int ox, oy, oz;             // byte offsets of the exit borders in xv / yv / zv
    int stx, sty, stz;          // +-8: offset increment of a crossing along each axis
    int dmx, dmy, dmz;          // cell number increments (m = k + Nz*j + Nz*Ny*i, :326-329)
    int m, tiny;                // tiny: bit a set when |k_a| <= 1e-15 (that axis is never crossed, :240-242)
    bool alive;

    // entry part, :151-230 (cartEnter), then the per-ray invariants of the crossing loop
    __device__ __forceinline__ int locator() const { return -1; }       // point location is cheap on this grid: nothing to remember
    __device__ __forceinline__ bool start(const CartGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en, int = -1)
    {
        alive = false;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        int i, j, k;
        if (!cartEnter(g, x, y, z, kx, ky, kz, en, i, j, k)) return false;
        const int Ny = g.Ny, Nz = g.Nz;
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
        // the cell entered now is the one whose density the NEXT crossing gathers (after its three divisions): start pulling
        // it into L1 here, half a crossing earlier than the gather itself (no register, no scoreboard)
        if (AHEAD && g.rhoAhead && alive) prefetchL1(g.rhoAhead + (size_t)m * g.rhoAheadStride);
        return ds > 0;
    }
};

// The walker of the photon SHOOTING stages (peel-off, escape + absorption, propagation).  Their outputs are Monte Carlo
// estimates gated at 3 sigma (BASELINE.json north_star), not the bit-exact path records, so the crossing does not have to
// reproduce the reference's rounding: the same DDA is carried in the path-length parameter t measured from the entry
// point -- per axis the value of t at which the ray leaves the current cell, t_a = (border_a - r0_a) / k_a evaluated as
// fma(border_a, 1/k_a, -r0_a/k_a) -- so that a crossing is three compares, one subtraction (ds = t_min - t), and for the
// crossed axis alone one shared-memory read of the next border and one FMA.  No division, no position update (the
// position is r0 + t k whenever it is needed), 6 fp64 instructions instead of ~35.  The same exit-face rule as
// CartesianDustGrid.cpp:243-269 (X if tx<=ty&&tx<=tz, else Y if ty<=tz, else Z), the same entry code (cartEnter), the
// same treatment of |k_a| <= 1e-15 (that axis is never crossed: t_a = DBL_MAX).  Differences to the exact walker are of
// the order of an ulp of t per segment (they do not accumulate: every t_a is computed from the border itself) and, at a
// crossing through a cell edge within that rounding, the order of two zero-length neighbours; skg_opticaldepth_mc exposes
// this walker so that tests can bound the deviation from the exact one.
// UNIFORM (every axis a LinMesh): consecutive borders of an axis are one bin width apart, so the exit parameter of
// the crossed axis advances by the constant |width / k_a| -- no shared-memory read at all in the crossing, and the walker
// only counts the cells left along each axis.  (The additions drift by an ulp of t per crossing, far inside the 1e-10 the
// walker is held to.)  Otherwise the next border is read from the staged arrays and t_a = fma(border, 1/k_a, -r0_a/k_a).
#ifndef SKG_FAST_UNROLL
#define SKG_FAST_UNROLL 4
#endif
template<bool UNIFORM> struct CartFastWalkerT
{
    static constexpr int kStepUnroll = SKG_FAST_UNROLL;
    static constexpr bool kSplitStep = false;
    static constexpr bool kPredicated = true;       // provides stepLive(): the scheduler runs batches of crossings branch-free
    double tx, ty, tz, t;
    double rkx, rky, rkz, cx, cy, cz;               // UNIFORM: rk* hold the constant steps |w_a / k_a|, c* are unused
    int ox, oy, oz, stx, sty, stz;                  // UNIFORM: o* count the cells left along each axis (current one included)
    int dmx, dmy, dmz, m;
    bool alive;

    __device__ __forceinline__ int locator() const { return -1; }
    __device__ __forceinline__ bool start(const CartGrid& g, Counters*, double x, double y, double z, double kx, double ky, double kz, Entry& en, int = -1)
    {
        alive = false;
        int i, j, k;
        if (!cartEnter(g, x, y, z, kx, ky, kz, en, i, j, k)) return false;
        const int Ny = g.Ny, Nz = g.Nz;
        m = k + Nz * j + Nz * Ny * i;
        const bool nx = kx < 0.0, ny = ky < 0.0, nz = kz < 0.0;
        dmx = nx ? -Ny * Nz : Ny * Nz; dmy = ny ? -Nz : Nz; dmz = nz ? -1 : 1;
        const bool ax = fabs(kx) > 1e-15, ay = fabs(ky) > 1e-15, az = fabs(kz) > 1e-15;
        const double ikx = ax ? 1.0 / kx : 0.0, iky = ay ? 1.0 / ky : 0.0, ikz = az ? 1.0 / kz : 0.0;
        const double ccx = ax ? -x * ikx : SKG_DBL_MAX, ccy = ay ? -y * iky : SKG_DBL_MAX, ccz = az ? -z * ikz : SKG_DBL_MAX;
        const int ex = 8 * (i + (nx ? 0 : 1)), ey = 8 * (j + (ny ? 0 : 1)), ez = 8 * (k + (nz ? 0 : 1));      // exit borders of the first cell
        tx = __fma_rn(ldsF64(g.sx + ex), ikx, ccx); ty = __fma_rn(ldsF64(g.sy + ey), iky, ccy); tz = __fma_rn(ldsF64(g.sz + ez), ikz, ccz);
        if (UNIFORM)
        {
            rkx = fabs(g.wx * ikx); rky = fabs(g.wy * iky); rkz = fabs(g.wz * ikz);
            ox = nx ? i + 1 : g.Nx - i; oy = ny ? j + 1 : Ny - j; oz = nz ? k + 1 : Nz - k;
        }
        else
        {
            rkx = ikx; rky = iky; rkz = ikz; cx = ccx; cy = ccy; cz = ccz;
            ox = ex; oy = ey; oz = ez; stx = nx ? -8 : 8; sty = ny ? -8 : 8; stz = nz ? -8 : 8;
        }
        t = 0.0;
        alive = true;
        return true;
    }

    // One crossing, fully predicated: `live` lanes advance, the others keep their state (so that the scheduler can run a
    // batch of crossings without a branch around each one).  Only the crossed axis is touched -- three small predicated
    // blocks instead of selects over the per-axis constants.  Returns true when segment (mseg, ds) is to be added; clears
    // `live` when the ray has left the grid (the staged arrays carry one pad element on either side: the last read is harmless).
    __device__ __forceinline__ bool stepLive(const CartGrid& g, bool& live, int& mseg, double& ds)
    {
        const bool bx = live && tx <= ty && tx <= tz;
        const bool by = live && !bx && ty <= tz;
        const bool bz = live && !bx && !by;
        const double tn = bx ? tx : (by ? ty : tz);
        mseg = m;
        ds = tn - t;
        const bool seg = live && ds > 0;
        if (live) t = tn;
        if (UNIFORM)
        {
            if (bx) { tx += rkx; m += dmx; live = --ox > 0; }
            if (by) { ty += rky; m += dmy; live = --oy > 0; }
            if (bz) { tz += rkz; m += dmz; live = --oz > 0; }
        }
        else
        {
            if (bx) { ox += stx; m += dmx; tx = __fma_rn(ldsF64(g.sx + ox), rkx, cx); live = (unsigned)ox <= 8u * g.Nx; }
            if (by) { oy += sty; m += dmy; ty = __fma_rn(ldsF64(g.sy + oy), rky, cy); live = (unsigned)oy <= 8u * g.Ny; }
            if (bz) { oz += stz; m += dmz; tz = __fma_rn(ldsF64(g.sz + oz), rkz, cz); live = (unsigned)oz <= 8u * g.Nz; }
        }
        return seg;
    }
    __device__ __forceinline__ bool step(const CartGrid& g, Counters*, int& mseg, double& ds)
    {
        bool live = true;
        const bool seg = stepLive(g, live, mseg, ds);
        alive = live;
        return seg;
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
    int node = 0, c0 = 0;
    bool haveC0 = false;
    if (g.lookupG > 1)
    {
        // start below the root when the lattice cell's node strictly contains the point: the descent from the root
        // necessarily passes through that node (all comparisons against its ancestors' split planes agree).  The node's
        // child index is read together with its box, and its record (what the Neighbor walker reads next when the node is
        // a leaf) is on its way into L1 meanwhile: one round trip after the lattice read instead of three in a row.
        const int G = g.lookupG;
        const int i = max(0, min(G - 1, (int)((x - g.box[0]) * g.lookupInv[0])));
        const int j = max(0, min(G - 1, (int)((y - g.box[1]) * g.lookupInv[1])));
        const int k = max(0, min(G - 1, (int)((z - g.box[2]) * g.lookupInv[2])));
        const int cand = __ldg(g.lookup + ((size_t)i * G + j) * G + k);
        const double* b = g.box + 6 * (size_t)cand;
        const int c0cand = __ldg(g.child0 + cand);
        if (g.nodeRec) { prefetchL1(g.nodeRec + cand); prefetchL1(reinterpret_cast<const char*>(g.nodeRec + cand) + 64); }
        if (x > __ldg(b) && x < __ldg(b + 3) && y > __ldg(b + 1) && y < __ldg(b + 4) && z > __ldg(b + 2) && z < __ldg(b + 5)) { node = cand; c0 = c0cand; haveC0 = true; }
    }
    if (!haveC0) c0 = __ldg(g.child0 + node);
    while (c0 >= 0)
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
        c0 = __ldg(g.child0 + node);
    }
    return node;
}

// rarely taken parts of a crossing, kept out of line so that the step body the warps loop over stays small
static __device__ __noinline__ int treeWhichNodeCold(const TreeGrid& g, double x, double y, double z) { return treeWhichNode(g, x, y, z); }
__device__ __forceinline__ double nextAfterAlong(double v, double k)
{
    return nextafter(v, (k < 0.0) ? -SKG_DBL_MAX : SKG_DBL_MAX);
}

// TreeDustGrid::path (TreeDustGrid.cpp:390-662) one crossing at a time.  HINT selects how a wall with several neighbours is
// searched (see step()): through the wall-bin table, or in list order starting from the neighbour held with the node.
#ifndef SKG_TREE_HINTS_PATH
#define SKG_TREE_HINTS_PATH true
#endif
#ifndef SKG_TREE_HINTS_MC
#define SKG_TREE_HINTS_MC false
#endif
template<bool HINT> struct TreeWalkerT
{
    static constexpr bool kPredicated = false;
    static constexpr int kStepUnroll = 1;       // large step body: a plain loop (unrolling it costs more in instruction fetch than it gains)
    // A crossing in two halves: step() finds the wall, moves the position and starts the read of the candidate node's record
    // (into L1), resolve() tests the candidate and continues from it.  With kSplitStep the scheduler runs the job's work on
    // the segment between the two, so that the dependent read overlaps it: +9 % on the adaptive mesh (C5), nothing on the
    // trees (C3: peel-off stage -2 %, path records -4 %), where step() therefore resolves at once.
#ifndef SKG_TREE_SPLIT
#define SKG_TREE_SPLIT false
#endif
    static constexpr bool kSplitStep = SKG_TREE_SPLIT;
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;
    double bx[6];               // box of the current node, carried over from the neighbour test that selected it
    int first[6];               // Neighbor search: first neighbour of each wall of the current node (TreeNodeRec)
    int hbase; unsigned hmeta;  // where its wall-bin blocks start, and which walls have one
    int node, cellv;
    int pcand, pinfo;           // between step() and resolve(): the candidate node; bit 0 pending, bit 1 strict test, bits 2-4 the wall
    bool alive;

    // one node record: 96 bytes as three 256-bit reads
    static __device__ __forceinline__ void loadRec(const TreeNodeRec* rec, double (&w)[12])
    {
        const double* rp = reinterpret_cast<const double*>(rec);
#pragma unroll
        for (int u = 0; u < 3; u++)
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(w[4 * u]), "=d"(w[4 * u + 1]), "=d"(w[4 * u + 2]), "=d"(w[4 * u + 3]) : "l"(rp + 4 * u));
    }
    static __device__ __forceinline__ bool recContains(const double (&w)[12], double x, double y, double z)
    { return x >= w[0] && x <= w[3] && y >= w[1] && y <= w[4] && z >= w[2] && z <= w[5]; }
    // continue from the node whose record has been read
    __device__ __forceinline__ void adopt(int id, const double (&w)[12])
    {
        for (int c = 0; c < 6; c++) bx[c] = w[c];
        const long long w6 = __double_as_longlong(w[6]), w8 = __double_as_longlong(w[8]),
                        w9 = __double_as_longlong(w[9]), w10 = __double_as_longlong(w[10]);
        node = id; cellv = (int)(w6 & 0xffffffffll); hbase = (int)(w6 >> 32); hmeta = (unsigned)(__double_as_longlong(w[7]) & 0xffffffffll);
        first[0] = (int)(w8 & 0xffffffffll); first[1] = (int)(w8 >> 32); first[2] = (int)(w9 & 0xffffffffll); first[3] = (int)(w9 >> 32);
        first[4] = (int)(w10 & 0xffffffffll); first[5] = (int)(w10 >> 32);
    }
    // everything a crossing needs from the node tables, fetched as soon as the node is known
    __device__ __forceinline__ void loadNode(const TreeGrid& g, bool withBox)
    {
        if (g.search == 1) { double w[12]; loadRec(g.nodeRec + node, w); adopt(node, w); return; }
        if (withBox) { const double* b = g.box + 6 * (size_t)node; for (int c = 0; c < 6; c++) bx[c] = __ldg(b + c); }
        cellv = __ldg(g.cell + node);
    }

    __device__ __forceinline__ int locator() const { return node; }
    // hint: the leaf the ray starts in, as remembered from an earlier traversal (the shooting stages: a packet keeps its
    // node between its peel-off, absorption and propagation walks), verified against the node's box; -1: locate the point
    __device__ __forceinline__ bool start(const TreeGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en, int hint = -1)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        bool located = false;
        if (hint >= 0)
        {
            node = hint; loadNode(g, true);
            located = cellv >= 0 && x > bx[0] && x < bx[3] && y > bx[1] && y < bx[4] && z > bx[2] && z < bx[5];
        }
        if (!located)
        {
            if (!moveInside(g.box, g.eps, x, y, z, kx, ky, kz, en)) return false;
            node = treeWhichNode(g, x, y, z);
            if (node < 0) return false;
            loadNode(g, true);
        }
        rkx = 1.0 / kx; rky = 1.0 / ky; rkz = 1.0 / kz;
        alive = true; pinfo = 0; pcand = -1;
        return true;
    }

    __device__ __forceinline__ bool step(const TreeGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        const bool nx = kx < 0.0, ny = ky < 0.0, nz = kz < 0.0;
        const double xnext = nx ? bx[0] : bx[3];
        const double ynext = ny ? bx[1] : bx[4];
        const double znext = nz ? bx[2] : bx[5];
        if (g.search == 3)
        {
            // ParticleTreeDustGrid::path (ParticleTreeDustGrid.cpp:262-320): unguarded divisions, the nearest wall among the
            // strictly positive distances, a segment only when there is one, then always a search from the root
            const double px = (xnext - x) / kx, py = (ynext - y) / ky, pz = (znext - z) / kz;
            ds = SKG_DBL_MAX;
            if (px > 0 && px < ds) ds = px;
            if (py > 0 && py < ds) ds = py;
            if (pz > 0 && pz < ds) ds = pz;
            const bool emit = ds < SKG_DBL_MAX;
            if (!emit) ds = 0;
            mseg = cellv;
            const double eps = g.eps;
            x += (ds + eps) * kx; y += (ds + eps) * ky; z += (ds + eps) * kz;
            const int oldnode = node;
            node = treeWhichNode(g, x, y, z);
            if (node == oldnode)
            {
                atomicAdd(&ctr->stuckEscaped, 1ull);
                x = nextAfterAlong(x, kx); y = nextAfterAlong(y, ky); z = nextAfterAlong(z, kz);
                node = treeWhichNodeCold(g, x, y, z);
                if (node == oldnode) { atomicAdd(&ctr->stuckTerminated, 1ull); node = -1; }
            }
            if (node < 0) alive = false; else loadNode(g, true);
            return emit;
        }
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

            pinfo = 1 | (wall << 2); pcand = -1;
            if (g.search == 1)
            {
                // TreeNode::whichnode(wall, r), TreeNode.cpp:84-93: first neighbour whose closed box contains r.
                // The reference sorts every list by decreasing overlap with the wall.  A wall with one neighbour, or whose first
                // neighbour covers at least half of it (a same-size or larger neighbour; the rest of such a list only touches the
                // wall's edges), tests that neighbour: its id travels with the node's record, one dependent read.  A wall shared
                // by several finer neighbours looks up (HINT) which of them covers the G x G wall bin r falls in and tests that
                // one: when r lies strictly inside its box no other leaf can contain r, so it is the list's first match as well.
                // Four finer siblings behind a wall (the usual level transition; tables.h, bit 4) need no table: the sibling
                // covering r's quadrant of the wall follows from the id of the first one, in the shooting stages as well.
                // Anything else (r on a box face, a neighbour finer than the bins resolve) searches the list in order.
                // The shooting stages measured 6 % faster searching such walls in list order from the record's first neighbour
                // (SKG_TREE_HINTS_MC), the path kernels 6-12 % faster with the table.
                const int f0 = wall == 0 ? first[0] : wall == 1 ? first[1] : wall == 2 ? first[2] : wall == 3 ? first[3] : wall == 4 ? first[4] : first[5];
                const unsigned wm = (hmeta >> (5 * wall)) & 31u;
                int cand = f0;
                bool strict = false;
                if ((wm & 9u) == 1u && (HINT || (wm & 16u)))
                {
                    const int a = wall < 2 ? 1 : 0, b = wall < 4 ? 2 : 1;
                    const double pa = a == 1 ? y : x, pb = b == 2 ? z : y;
                    const double la = a == 1 ? bx[1] : bx[0], ha = a == 1 ? bx[4] : bx[3], lb = b == 2 ? bx[2] : bx[1], hb = b == 2 ? bx[5] : bx[4];
                    if (wm & 16u)
                    {
                        // four finer siblings behind the wall: the one covering r's half of each in-plane axis, by arithmetic
                        // (f0 holds the id of the first of them; verified like any other candidate below)
                        cand = f0 + ((pa - la > ha - pa) ? (wall < 2 ? 2 : 1) : 0) + ((pb - lb > hb - pb) ? (wall < 4 ? 4 : 2) : 0);
                    }
                    else
                    {
                        const int G = 2 << ((wm >> 1) & 3u);
                        int off = 0;            // ids of the multi-neighbour walls before this one
                        for (int v = 0; v < 5; v++) { const unsigned vm = (hmeta >> (5 * v)) & 7u; off += (v < wall && (vm & 1u)) ? (4 << (vm & 6u)) : 0; }
                        const int ia = max(0, min(G - 1, (int)((float)G * (float)(pa - la) / (float)(ha - la))));
                        const int ib = max(0, min(G - 1, (int)((float)G * (float)(pb - lb) / (float)(hb - lb))));
                        cand = __ldg(g.nbrHint + 4 * (size_t)hbase + off + G * ia + ib);
                    }
                    strict = true;
                }
                pcand = cand; pinfo |= strict ? 2 : 0;
                if (kSplitStep && cand >= 0) { prefetchL1(g.nodeRec + cand); prefetchL1(reinterpret_cast<const char*>(g.nodeRec + cand) + 64); }
            }
            if (!kSplitStep) resolve(g, ctr);
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

    // second half of a TopDown / Neighbor crossing
    __device__ __forceinline__ void resolve(const TreeGrid& g, Counters* ctr)
    {
        if (!(pinfo & 1)) return;
        const bool strict = (pinfo & 2) != 0; const int wall = pinfo >> 2; const int cand = pcand;
        pinfo = 0;
        const unsigned wm = (hmeta >> (5 * wall)) & 31u;
        {
            const int oldnode = node;
            bool haveBox = false, haveAll = false;
            if (g.search == 1)
            {
                node = -1;
                if (cand >= 0)
                {
                    double w[12]; loadRec(g.nodeRec + cand, w);
                    const bool ok = strict ? (x > w[0] && x < w[3] && y > w[1] && y < w[4] && z > w[2] && z < w[5]) : recContains(w, x, y, z);
                    if (ok) { adopt(cand, w); haveAll = true; }
                    else if (wm & 1u)
                    {
                        // in list order (from the second entry when the first one has just been tested)
                        const int beg = __ldg(g.nbrStart + 6 * (size_t)oldnode + wall), end = __ldg(g.nbrStart + 6 * (size_t)oldnode + wall + 1);
                        for (int q = strict ? beg : beg + 1; q < end; q++)
                        {
                            const int c2 = __ldg(g.nbrIds + q);
                            loadRec(g.nodeRec + c2, w);
                            if (recContains(w, x, y, z)) { adopt(c2, w); haveAll = true; break; }
                        }
                    }
                }
                // (TreeNode::whichnode(Vec) of the root answers "none" for a point outside the root's box: the usual end of a
                // path, decided here without the out-of-line call)
                if (node < 0 && boxContains(g.box, x, y, z)) node = treeWhichNodeCold(g, x, y, z);
            }
            else node = treeWhichNode(g, x, y, z);

            if (node == oldnode)
            {
                atomicAdd(&ctr->stuckEscaped, 1ull);
                x = nextAfterAlong(x, kx); y = nextAfterAlong(y, ky); z = nextAfterAlong(z, kz);
                node = treeWhichNodeCold(g, x, y, z);
                haveBox = false; haveAll = false;
                if (node == oldnode) { atomicAdd(&ctr->stuckTerminated, 1ull); node = -1; }
            }
            if (node < 0) alive = false;
            else if (!haveAll) loadNode(g, !haveBox);
        }
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

// one node record: 96 bytes as three 256-bit reads
struct AMeshRecWords
{
    double w[12];
    __device__ __forceinline__ void load(const AMeshNodeRec* rec)
    {
        const double* rp = reinterpret_cast<const double*>(rec);
#pragma unroll
        for (int u = 0; u < 3; u++)
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(w[4 * u]), "=d"(w[4 * u + 1]), "=d"(w[4 * u + 2]), "=d"(w[4 * u + 3]) : "l"(rp + 4 * u));
    }
    __device__ __forceinline__ bool contains(double x, double y, double z) const { return x >= w[0] && x <= w[3] && y >= w[1] && y <= w[4] && z >= w[2] && z <= w[5]; }
    __device__ __forceinline__ int lo(int i) const { return (int)(__double_as_longlong(w[i]) & 0xffffffffll); }
    __device__ __forceinline__ int hi(int i) const { return (int)(__double_as_longlong(w[i]) >> 32); }
    __device__ __forceinline__ int cell() const { return lo(6); }
    __device__ __forceinline__ int child0() const { return hi(6); }
    __device__ __forceinline__ int nx() const { return lo(7); }
    __device__ __forceinline__ int ny() const { return hi(7); }
    __device__ __forceinline__ int nz() const { return lo(8); }
    __device__ __forceinline__ int parent() const { return hi(8); }
};

// AdaptiveMeshNode::whichnode(Vec) from the root, AdaptiveMeshNode.cpp:132-142 (+ child :109-128).
// Returns -1 when outside, -2 when the reference would throw "Can't locate the appropriate child node".
// On success `rec` holds the record of the returned leaf (one dependent read per level: the child's record is both the
// containment test of this level and the node of the next).
// the descent of AdaptiveMeshNode::whichnode from `node`, whose record is in `rec`
__device__ __forceinline__ int ameshDescend(const AMeshGrid& g, int node, double x, double y, double z, AMeshRecWords& rec)
{
    int c0;
    while ((c0 = rec.child0()) >= 0)
    {
        const int Nx = rec.nx(), Ny = rec.ny(), Nz = rec.nz();
        int i = cellIndex1(x, rec.w[0], rec.w[3], Nx);
        int j = cellIndex1(y, rec.w[1], rec.w[4], Ny);
        int k = cellIndex1(z, rec.w[2], rec.w[5], Nz);
        int child = c0 + (k * Ny + j) * Nx + i;
        rec.load(g.nodeRec + child);
        if (!rec.contains(x, y, z))
        {
            if (x < rec.w[0]) i--; else if (x > rec.w[3]) i++;
            if (y < rec.w[1]) j--; else if (y > rec.w[4]) j++;
            if (z < rec.w[2]) k--; else if (z > rec.w[5]) k++;
            if (i < 0 || i >= Nx || j < 0 || j >= Ny || k < 0 || k >= Nz) return -2;
            child = c0 + (k * Ny + j) * Nx + i;
            rec.load(g.nodeRec + child);
            if (!rec.contains(x, y, z)) return -2;
        }
        node = child;
    }
    return node;
}
__device__ __forceinline__ int ameshWhichNode(const AMeshGrid& g, double x, double y, double z, AMeshRecWords& rec)
{
    if (!boxContains(g.box, x, y, z)) return -1;
    rec.load(g.nodeRec);
    return ameshDescend(g, 0, x, y, z, rec);
}
__device__ __forceinline__ int ameshWhichNode(const AMeshGrid& g, double x, double y, double z)
{ AMeshRecWords rec; return ameshWhichNode(g, x, y, z, rec); }
static __device__ __noinline__ int ameshWhichNodeCold(const AMeshGrid& g, double x, double y, double z) { return ameshWhichNode(g, x, y, z); }

// AdaptiveMesh::path (AdaptiveMesh.cpp:297-367) one crossing at a time
struct AMeshWalker
{
    static constexpr bool kPredicated = false;
    static constexpr int kStepUnroll = 1;
    static constexpr bool kSplitStep = true;    // see TreeWalkerT
    double x, y, z, kx, ky, kz;
    double rkx, rky, rkz;
    double bx[6];               // box of the current node (carried over from the neighbour test that selected it)
    int wn[6];                  // the node beyond each wall of the current leaf
    int node, cellv;
    int pcand; bool pending;    // between step() and resolve(): the node beyond the wall just crossed
    bool alive;

    // continue from the leaf whose record has been read
    __device__ __forceinline__ void adopt(int id, const AMeshRecWords& r)
    {
        for (int c = 0; c < 6; c++) bx[c] = r.w[c];
        node = id; cellv = r.cell();
        wn[0] = r.lo(9); wn[1] = r.hi(9); wn[2] = r.lo(10); wn[3] = r.hi(10); wn[4] = r.lo(11); wn[5] = r.hi(11);
    }

    __device__ __forceinline__ int locator() const { return node; }
    __device__ __forceinline__ bool start(const AMeshGrid& g, Counters* ctr, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en, int hint = -1)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        AMeshRecWords rec;
        bool located = false;
        if (hint >= 0)
        {
            // the leaf remembered from an earlier traversal of the same packet (see TreeWalkerT::start), verified against its box
            rec.load(g.nodeRec + hint);
            located = rec.cell() >= 0 && x > rec.w[0] && x < rec.w[3] && y > rec.w[1] && y < rec.w[4] && z > rec.w[2] && z < rec.w[5];
            if (located) adopt(hint, rec);
        }
        if (!located)
        {
            if (!moveInside(g.box, g.eps, x, y, z, kx, ky, kz, en)) return false;
            const int id = ameshWhichNode(g, x, y, z, rec);
            if (id < 0) { if (id == -2) atomicAdd(&ctr->errors, 1ull); return false; }
            adopt(id, rec);
        }
        rkx = 1.0 / kx; rky = 1.0 / ky; rkz = 1.0 / kz;
        alive = true; pending = false; pcand = -1;
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

        pcand = wall == 0 ? wn[0] : wall == 1 ? wn[1] : wall == 2 ? wn[2] : wall == 3 ? wn[3] : wall == 4 ? wn[4] : wn[5];
        pending = true;
        if (pcand >= 0) { prefetchL1(g.nodeRec + pcand); prefetchL1(reinterpret_cast<const char*>(g.nodeRec + pcand) + 64); }
        return ds > 0;
    }

    // second half of a crossing (see TreeWalkerT): the node beyond the wall, AdaptiveMesh.cpp:338-342
    __device__ __forceinline__ void resolve(const AMeshGrid& g, Counters* ctr)
    {
        if (!pending) return;
        pending = false;
        const int oldnode = node;
        const int cand = pcand;
        AMeshRecWords rec;
        int id = -3;
        if (cand >= 0)
        {
            rec.load(g.nodeRec + cand);
            if (rec.contains(x, y, z)) id = cand;
        }
        if (id == -3)
        {
            // r lies beside the neighbour at the wall's centre -- in one of its siblings, when the wall borders finer cells --
            // or the wall has none.  The reference searches from the root (AdaptiveMesh.cpp:341-342); a node whose box holds r
            // strictly inside lies on that descent (r is in no sibling's closed box at any level above), so the search
            // resumes from the nearest such ancestor of the neighbour: two or three dependent reads instead of one per level.
            int from = -1;
            if (cand >= 0)
            {
                int anc = rec.parent();
                for (int climb = 0; climb < 3 && anc > 0; climb++)
                {
                    rec.load(g.nodeRec + anc);
                    if (x > rec.w[0] && x < rec.w[3] && y > rec.w[1] && y < rec.w[4] && z > rec.w[2] && z < rec.w[5]) { from = anc; break; }
                    anc = rec.parent();
                }
            }
            if (from < 0)
            {
                if (boxContains(g.box, x, y, z)) { rec.load(g.nodeRec); from = 0; }
                else id = -1;
            }
            if (from >= 0) id = ameshDescend(g, from, x, y, z, rec);
        }
        if (id == -2) { atomicAdd(&ctr->errors, 1ull); alive = false; return; }

        if (id == oldnode)
        {
            atomicAdd(&ctr->stuckEscaped, 1ull);
            x = nextAfterAlong(x, kx); y = nextAfterAlong(y, ky); z = nextAfterAlong(z, kz);
            id = ameshWhichNodeCold(g, x, y, z);
            if (id == -2) { atomicAdd(&ctr->errors, 1ull); alive = false; return; }
            if (id == oldnode) { atomicAdd(&ctr->stuckTerminated, 1ull); id = -1; }
            if (id >= 0) rec.load(g.nodeRec + id);
        }
        if (id < 0) { node = id; alive = false; }
        else adopt(id, rec);
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
    int beg = __ldg(g.blkStart + b), end = __ldg(g.blkStart + b + 1);
    // A block with a search tree (more than five cells, VoronoiMesh.cpp:371) answers with the nearest of its cells, which the
    // tree search (Node::nearest) finds through ~25 dependent reads.  The same cell is the minimum over the block's list --
    // independent reads, three round trips in all -- unless two cells are at bit-equal distance: only then does the order
    // of the tree's comparisons decide, and the tree is searched.  (Blocks without tree: the reference's own loop, first minimum.)
    if (tree >= 0 && (end - beg < 1 || end - beg > 64)) return voroKdNearest(g, tree, x, y, z);
    int m = -1;
    double mdist = SKG_DBL_MAX;
    bool tie = false;
#pragma unroll 4
    for (int q = beg; q < end; q++)
    {
        int id = __ldg(g.blkIds + q);
        double idist = voroSD(g, id, x, y, z);
        if (idist < mdist) { m = id; mdist = idist; tie = false; }
        else if (idist == mdist) tie = true;
    }
    if (tree >= 0 && tie) return voroKdNearest(g, tree, x, y, z);
    return m;
}

// VoronoiMesh::path (VoronoiMesh.cpp:749-844) one crossing at a time.
// EXACT: the reference's arithmetic -- every candidate wall distance is the quotient si = n.(p - r) / n.k and the smallest
// positive one wins, first in list order among equals (the deterministic-geometry entry points: bit-exact paths).
// !EXACT (the photon shooting stages, whose results are Monte Carlo estimates), meshes whose records do not stay in L2
// (no plane table): the crossing records with the candidates compared as fractions, num_i * den_best < num_best * den_i
// (all denominators positive), in list order with the same first-wins rule; only the winner is divided.
// !EXACT, meshes whose table stays in L2 (`planes`, built by the engine up to 128 MB): every candidate -- bisector planes and
// domain walls alike -- is a normal n relative to the cell's particle p (tables.h, `planes`; a wall is the bisector towards
// p's mirror image), its distance the fraction num / den = n.(n - 2 (r - p)) / 2 n.k.  Candidates are compared as fractions,
// num_i * den_best < num_best * den_i (all denominators positive), in the same list order with the same first-wins rule;
// only the winner is divided.  The loop body is three subtractions, six multiply-adds, two products and a compare,
// branch-free -- the reference's arithmetic (the EXACT walker) costs twice that per neighbour, with a division and a
// branch: C4 at 200 000 cells +31 %.  At 1e6 cells, where every crossing reads its 544 bytes from DRAM, the same loop
// measured 8 % slower than the crossing records (profiles/r02_q_voronoi_planes.txt), hence the size rule.
// Results agree with the exact walker to rounding (tests: 1e-10 on optical depths); two candidates whose quotients differ
// by less than the rounding of the products may swap, which moves a crossing by an ulp.
template<bool EXACT> struct VoroWalkerT
{
    static constexpr bool kPredicated = false;
    static constexpr int kStepUnroll = 1;
    static constexpr bool kSplitStep = false;
    double x, y, z, kx, ky, kz;
    int mr, rr;                 // current cell and the first slot of its crossing record
    int guard;
    bool alive;

    __device__ __forceinline__ int locator() const { return mr; }
    // hint: the cell the ray starts in, as remembered from an earlier traversal of the same packet (taken on trust: a point
    // that a path placed inside cell m is nearest to particle m up to rounding); -1: VoronoiMesh::cellIndex
    __device__ __forceinline__ bool start(const VoroGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en, int hint = -1)
    {
        alive = false; en.n = 0; guard = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (!finite3(x, y, z) || !finite3(kx, ky, kz)) return false;
        if (hint >= 0 && hint < g.N) mr = hint;
        else
        {
            if (!moveInside(g.ext, g.eps, x, y, z, kx, ky, kz, en)) return false;
            mr = voroCellIndex(g, x, y, z);
        }
        if (mr < 0) return false;
        rr = blockOf(g, mr);
        alive = true;
        return true;
    }
    // first slot of the crossing record (EXACT) / plane record (!EXACT) of a cell: the two tables share their layout
    static __device__ __forceinline__ int blockOf(const VoroGrid& g, int m) { return __ldg(g.nbrStart + m) + m; }

    static __device__ __forceinline__ void loadSlot(const double* p, double (&w)[4])
    { asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(w[0]), "=d"(w[1]), "=d"(w[2]), "=d"(w[3]) : "l"(p)); }

    __device__ __forceinline__ bool stepPlanes(const VoroGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        const double eps = g.eps;
        const double* R = g.planes + 4 * (size_t)rr;
        double h[4], e[4][4];
        loadSlot(R, h);
#pragma unroll
        for (int u = 0; u < 4; u++) loadSlot(R + 4 * (u + 1), e[u]);        // the table ends with 8 spare slots
        const int cnt = (int)(__double_as_longlong(h[3]) & 0xffffffffll);
        // the rest of the block (typically 12 more slots = 3 lines): into L1 while the first group is evaluated
        for (int q0 = 4; q0 < cnt; q0 += 4) prefetchL1(R + 4 * (q0 + 1));
        const double qx2 = 2.0 * (x - h[0]), qy2 = 2.0 * (y - h[1]), qz2 = 2.0 * (z - h[2]);
        double nb = 0.0, db = 1.0;              // best fraction nb / db (db > 0) = twice the distance
        long long btag = 0;
        bool have = false;
        for (int q0 = 0; q0 < cnt; q0 += 4)
        {
            if (q0)
            {
#pragma unroll
                for (int u = 0; u < 4; u++) loadSlot(R + 4 * (q0 + u + 1), e[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                // n.(n - 2 (r - p)) / n.k = twice the distance to the plane along the ray
                const double nx = e[u][0], ny = e[u][1], nz = e[u][2];
                const double num = __fma_rn(nx, nx - qx2, __fma_rn(ny, ny - qy2, nz * (nz - qz2)));
                const double den = __fma_rn(nx, kx, __fma_rn(ny, ky, nz * kz));
                // si > 0 && si < sq with si = num / den, sq = nb / db
                const bool better = q0 + u < cnt && den > 0 && num > 0 && (!have || num * db < nb * den);
                if (better) { nb = num; db = den; btag = __double_as_longlong(e[u][3]); have = true; }
            }
        }
        if (!have)
        {
            // r += bfk*_eps  (Vec operator*(Vec,double))
            x += kx * eps; y += ky * eps; z += kz * eps;
            mr = voroCellIndex(g, x, y, z);
            if (++guard > 1000000) { atomicAdd(&ctr->errors, 1ull); mr = -1; }
            if (mr < 0) alive = false; else rr = blockOf(g, mr);
            return false;
        }
        const double sq = 0.5 * (nb / db);      // > 0 by construction
        mseg = mr; ds = sq;
        x += (sq + eps) * kx; y += (sq + eps) * ky; z += (sq + eps) * kz;
        mr = (int)(btag & 0xffffffffll); rr = (int)(btag >> 32);
        if (mr < -6) { atomicAdd(&ctr->errors, 1ull); }
        if (mr < 0) alive = false;
        return true;
    }

    __device__ __forceinline__ bool step(const VoroGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        // (warp-uniform: a property of the grid -- the engine builds plane records for meshes whose table stays in L2)
        if constexpr (!EXACT) { if (g.planes) return stepPlanes(g, ctr, mseg, ds); }
        const double eps = g.eps;
        double sq = SKG_DBL_MAX;                // EXACT: best quotient
        double nb = 0.0, db = 1.0;              // !EXACT: best fraction nb / db (db > 0); mq == NO_INDEX: none yet
        const int NO_INDEX = -99;
        int mq = NO_INDEX, rq = 0;
        // the neighbour loop of VoronoiMesh.cpp:777-828 over the cell's crossing record (tables.h), four neighbours at a
        // time: the header and the first group travel together, every slot is one 256-bit read, and the neighbours are
        // evaluated in list order so that the first smallest intersection still wins
        const double* R = g.rec + 4 * (size_t)rr;
        double h[4], e[4][4];
        loadSlot(R, h);
#pragma unroll
        for (int u = 0; u < 4; u++) loadSlot(R + 4 * (u + 1), e[u]);        // the table ends with 8 spare slots
        const double prx = h[0], pry = h[1], prz = h[2];
        const int cnt = (int)(__double_as_longlong(h[3]) & 0xffffffffll);
        // the rest of the block (typically 12 more slots = 3 lines): into L1 while the first group is evaluated
        for (int q0 = 4; q0 < cnt; q0 += 4) prefetchL1(R + 4 * (q0 + 1));
        for (int q0 = 0; q0 < cnt; q0 += 4)
        {
            if (q0)
            {
#pragma unroll
                for (int u = 0; u < 4; u++) loadSlot(R + 4 * (q0 + u + 1), e[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                if (q0 + u >= cnt) break;
                const long long tag = __double_as_longlong(e[u][3]);
                const int mi = (int)(tag & 0xffffffffll);
                if (EXACT)
                {
                    double si = 0;
                    if (mi >= 0)
                    {
                        const double pix = e[u][0], piy = e[u][1], piz = e[u][2];
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
                        switch (mi)
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
                    if (si > 0 && si < sq) { sq = si; mq = mi; rq = (int)(tag >> 32); }
                }
                else
                {
                    // the candidate as a fraction num / den, den > 0 (a wall: (border - r_a) / k_a with the sign moved into num)
                    double num, den;
                    if (mi >= 0)
                    {
                        const double pix = e[u][0], piy = e[u][1], piz = e[u][2];
                        const double nxv = pix - prx, nyv = piy - pry, nzv = piz - prz;
                        den = nxv * kx + nyv * ky + nzv * kz;
                        num = nxv * (0.5 * (pix + prx) - x) + nyv * (0.5 * (piy + pry) - y) + nzv * (0.5 * (piz + prz) - z);
                    }
                    else
                    {
                        if (mi < -6) { atomicAdd(&ctr->errors, 1ull); alive = false; return false; }
                        const int a = (-1 - mi) >> 1;                                   // axis of the wall; odd ids are the lower faces
                        const double ka = a == 0 ? kx : (a == 1 ? ky : kz), ra = a == 0 ? x : (a == 1 ? y : z);
                        const double border = g.ext[a + (((-1 - mi) & 1) ? 3 : 0)];
                        num = border - ra; den = ka;
                        if (den < 0) { num = -num; den = -den; }
                    }
                    // si > 0 && si < sq with si = num / den, sq = nb / db
                    const bool better = den > 0 && num > 0 && (mq == NO_INDEX || num * db < nb * den);
                    if (better) { nb = num; db = den; mq = mi; rq = (int)(tag >> 32); }
                }
            }
        }
        if (mq == NO_INDEX)
        {
            // r += bfk*_eps  (Vec operator*(Vec,double))
            x += kx * eps; y += ky * eps; z += kz * eps;
            mr = voroCellIndex(g, x, y, z);
            if (++guard > 1000000) { atomicAdd(&ctr->errors, 1ull); mr = -1; }
            if (mr < 0) alive = false; else rr = blockOf(g, mr);
            return false;
        }
        if (!EXACT) sq = nb / db;
        mseg = mr; ds = sq;                     // sq > 0 by construction
        x += (sq + eps) * kx; y += (sq + eps) * ky; z += (sq + eps) * kz;
        mr = mq; rr = rq;
        if (mr < 0) alive = false;
        return true;
    }
};


// ---------------------------------------------------------------------------------------------------
// Grids with symmetries: Sphere1DDustGrid, Sphere2DDustGrid, Cylinder2DDustGrid
// ---------------------------------------------------------------------------------------------------
// Sphere2DDustGrid.cpp:175-226: smallest positive solution of x^2 + 2 b x + c = 0 (0: none) and friends
__device__ __forceinline__ double symSmallestPositive(double b, double c)
{
    if (b * b > c)
    {
        if (b > 0) { if (c < 0) { double x1 = -b - sqrt(b * b - c); return c / x1; } }
        else
        {
            double x2 = -b + sqrt(b * b - c);
            if (c > 0) { double x1 = c / x2; if (x1 < x2) return x1; }
            return x2;
        }
    }
    return 0;
}
__device__ __forceinline__ double symSmallestPositive(double a, double b, double c)
{
    if (fabs(a) > 1e-9) return symSmallestPositive(b / a, c / a);
    double x = -0.5 * c / b;
    if (x > 0) return x;
    return 0;
}
__device__ __forceinline__ double symFirstSphere(double x, double y, double z, double kx, double ky, double kz, double r)
{ return symSmallestPositive(x * kx + y * ky + z * kz, (x * x + y * y + z * z) - r * r); }            // Vec::dot, Vec::norm2 (Vec.hpp)
__device__ __forceinline__ double symFirstCone(double x, double y, double z, double kx, double ky, double kz, double c)
{
    if (c == 0) return -z / kz;
    const double rk = x * kx + y * ky + z * kz, r2 = x * x + y * y + z * z;
    return symSmallestPositive(c * c - kz * kz, c * c * rk - z * kz, c * c * r2 - z * z);
}
// Position::spherical, Position.cpp:91-106 (r and theta only)
__device__ __forceinline__ void symSpherical(double x, double y, double z, double& r, double& theta)
{
    r = sqrt(x * x + y * y + z * z);
    theta = r == 0 ? 0.0 : acos(z / r);
}
__device__ __forceinline__ int symWhichCell(const SymGrid& g, double x, double y, double z)
{
    if (g.sub == 0) return locateFail(g.v1, sqrt(x * x + y * y + z * z), g.N1 + 1);                    // Sphere1DDustGrid.cpp:81-84
    if (g.sub == 1)                                                                                     // Sphere2DDustGrid.cpp:133-141
    {
        double r, theta; symSpherical(x, y, z, r, theta);
        int i = locateFail(g.v1, r, g.N1 + 1); if (i < 0) return -1;
        return locateClip(g.v2, theta, g.N2 + 1) + g.N2 * i;
    }
    int i = locateFail(g.v1, sqrt(x * x + y * y), g.N1 + 1), k = locateFail(g.v2, z, g.N2 + 1);         // Cylinder2DDustGrid.cpp:95-101
    if (i < 0 || k < 0) return -1;
    return k + g.N2 * i;
}

// The three path() bodies as state machines that yield one segment per step.
//   Sphere1DDustGrid::path  Sphere1DDustGrid.cpp:111-186 (phase 0: inward towards the impact shell, phase 1: outward)
//   Cylinder2DDustGrid::path Cylinder2DDustGrid.cpp:135-374 (the same two phases, upward or downward in z)
//   Sphere2DDustGrid::path  Sphere2DDustGrid.cpp:230-345 (position-based: nearest of up to four cell boundaries)
struct SymWalker
{
    static constexpr bool kPredicated = false;
    static constexpr int kStepUnroll = 1;
    static constexpr bool kSplitStep = false;
    double x, y, z, kx, ky, kz;     // Sphere2D: current position and direction; Cylinder2D: z and kz are the running height / its cosine
    double p, q, qN, zN, kq;        // impact parameter, running and next radial path coordinates (Sphere1D, Cylinder2D); next z border
    int i, k, imin; int phase;      // cell indices; phase 0 inward, 1 outward
    bool alive;

    __device__ __forceinline__ int locator() const { return -1; }

    __device__ __forceinline__ bool start(const SymGrid& g, Counters*, double x0, double y0, double z0, double kx0, double ky0, double kz0, Entry& en, int = -1)
    {
        alive = false; en.n = 0;
        x = x0; y = y0; z = z0; kx = kx0; ky = ky0; kz = kz0;
        if (g.sub == 0)
        {
            const double rmax = g.rmax;
            double r = sqrt(x * x + y * y + z * z);                 // Position::radius() = Vec::norm()
            q = x * kx + y * ky + z * kz;
            p = sqrt((r - q) * (r + q));
            if (r > rmax)
            {
                if (q > 0.0 || p > rmax) return false;
                r = rmax - 1e-8 * (g.v1[g.N1] - g.v1[g.N1 - 1]);
                const double qmax = sqrt((rmax - p) * (rmax + p));
                en.ds[en.n++] = qmax - q;
                q = qmax;
            }
            i = locateClip(g.v1, r, g.N1 + 1);
            phase = 1;
            if (q < 0.0)
            {
                imin = locateClip(g.v1, p, g.N1 + 1);
                const double rN = g.v1[i];
                qN = -sqrt((rN - p) * (rN + p));
                phase = i > imin ? 0 : 1;
            }
            if (phase == 1) { const double rN = g.v1[i + 1]; qN = sqrt((rN - p) * (rN + p)); }
            alive = true;
            return true;
        }
        if (g.sub == 2)
        {
            kq = sqrt(kx * kx + ky * ky);
            if (kz == 0.0) kz = 1e-20;
            if (kq == 0.0) kq = 1e-20;
            double R = sqrt(x * x + y * y);                          // Position::cylradius()
            q = (x * kx + y * ky) / kq;
            const double p2 = (R - q) * (R + q);
            p = sqrt(fmax(0.0, p2));
            const double Rmax = g.rmax, zmin = g.zmin, zmax = g.zmax;
            if (R >= Rmax)
            {
                if (q > 0.0 || p > Rmax) return false;
                R = Rmax - 1e-8 * (g.v1[g.N1] - g.v1[g.N1 - 1]);
                const double qmax = sqrt((Rmax - p) * (Rmax + p));
                const double ds = (qmax - q) / kq;
                en.ds[en.n++] = ds;
                q = qmax;
                z += kz * ds;
            }
            if (z < zmin)
            {
                if (kz <= 0.0) { en.n = 0; return false; }
                const double ds = (zmin - z) / kz;
                en.ds[en.n++] = ds;
                q += kq * ds;
                R = sqrt(p * p + q * q);
                z = zmin + 1e-8 * (g.v2[1] - g.v2[0]);
            }
            else if (z > zmax)
            {
                if (kz >= 0.0) { en.n = 0; return false; }
                const double ds = (zmax - z) / kz;
                en.ds[en.n++] = ds;
                q += kq * ds;
                R = sqrt(p * p + q * q);
                z = zmax - 1e-8 * (g.v2[g.N2] - g.v2[g.N2 - 1]);
            }
            if (isinf(R) || isnan(R) || isinf(z) || isnan(z) || R >= Rmax || z <= zmin || z >= zmax) { en.n = 0; return false; }
            i = locateClip(g.v1, R, g.N1 + 1);
            k = locateClip(g.v2, z, g.N2 + 1);
            const bool up = kz >= 0.0;
            zN = up ? g.v2[k + 1] : g.v2[k];
            phase = 1;
            if (q < 0.0)
            {
                imin = locateClip(g.v1, p, g.N1 + 1);
                const double RN = g.v1[i];
                qN = -sqrt((RN - p) * (RN + p));
                phase = i > imin ? 0 : 1;
            }
            if (phase == 1) { const double RN = g.v1[i + 1]; qN = sqrt((RN - p) * (RN + p)); }
            alive = true;
            return true;
        }
        // Sphere2D
        const double rmax = g.rmax, eps = 1e-11 * rmax;
        const double r2 = x * x + y * y + z * z;
        if (r2 > rmax * rmax)
        {
            const double ds = symFirstSphere(x, y, z, kx, ky, kz, rmax);
            if (!(ds != 0)) return false;
            en.ds[en.n++] = ds;
            x += kx * (ds + eps); y += ky * (ds + eps); z += kz * (ds + eps);
        }
        else if (r2 == 0) { x += kx * eps; y += ky * eps; z += kz * eps; }
        double r, theta; symSpherical(x, y, z, r, theta);
        i = locateFail(g.v1, r, g.N1 + 1);
        k = locateClip(g.v2, theta, g.N2 + 1);
        alive = i < g.N1 && i >= 0;
        return true;
    }

    __device__ __forceinline__ bool step(const SymGrid& g, Counters* ctr, int& mseg, double& ds)
    {
        if (g.sub == 0)
        {
            mseg = i; ds = qN - q;
            if (phase == 0)
            {
                i--; q = qN;
                const double rN = g.v1[i];
                qN = -sqrt((rN - p) * (rN + p));
                if (!(i > imin)) { phase = 1; const double rO = g.v1[i + 1]; qN = sqrt((rO - p) * (rO + p)); }
            }
            else
            {
                i++;
                if (i >= g.N1 - 1) alive = false;           // (the reference never enters the outermost shell on the way out, Sphere1DDustGrid.cpp:178)
                else { q = qN; const double rN = g.v1[i + 1]; qN = sqrt((rN - p) * (rN + p)); }
            }
            return ds > 0;                                  // DustGridPath::addSegment keeps only segments with ds > 0
        }
        if (g.sub == 2)
        {
            const bool up = kz >= 0.0;
            mseg = k + g.N2 * i;
            const double dsq = (qN - q) / kq, dsz = (zN - z) / kz;
            if (dsq < dsz)
            {
                ds = dsq;
                if (phase == 0)
                {
                    i--; q = qN; z += kz * ds;
                    const double RN = g.v1[i];
                    qN = -sqrt((RN - p) * (RN + p));
                    if (!(i > imin)) { phase = 1; const double RO = g.v1[i + 1]; qN = sqrt((RO - p) * (RO + p)); }
                }
                else
                {
                    i++;
                    if (i >= (up ? g.N1 : g.N1 - 1)) alive = false;     // (:258 upward, :353 downward: the reference's own asymmetry)
                    else { q = qN; z += kz * ds; const double RN = g.v1[i + 1]; qN = sqrt((RN - p) * (RN + p)); }
                }
            }
            else
            {
                ds = dsz;
                if (up) { k++; if (k >= g.N2) alive = false; else { q += kq * ds; z = zN; zN = g.v2[k + 1]; } }
                else { k--; if (k < 0) alive = false; else { q += kq * ds; z = zN; zN = g.v2[k]; } }
            }
            return ds > 0;
        }
        // Sphere2D
        const double eps = 1e-11 * g.rmax;
        int inext = i, knext = k;
        ds = SKG_DBL_MAX;
        if (i > 0) { const double s = symFirstSphere(x, y, z, kx, ky, kz, g.v1[i]); if (s > 0 && s < ds) { ds = s; inext = i - 1; knext = k; } }
        { const double s = symFirstSphere(x, y, z, kx, ky, kz, g.v1[i + 1]); if (s > 0 && s < ds) { ds = s; inext = i + 1; knext = k; } }
        if (k > 0) { const double s = symFirstCone(x, y, z, kx, ky, kz, g.cv[k]); if (s > 0 && s < ds) { ds = s; inext = i; knext = k - 1; } }
        if (k < g.N2 - 1) { const double s = symFirstCone(x, y, z, kx, ky, kz, g.cv[k + 1]); if (s > 0 && s < ds) { ds = s; inext = i; knext = k + 1; } }
        bool seg = false;
        if (inext != i || knext != k)
        {
            mseg = k + g.N2 * i; seg = ds > 0;
            x += kx * (ds + eps); y += ky * (ds + eps); z += kz * (ds + eps);
            i = inext; k = knext;
        }
        else
        {
            // "No exit point found from dust grid cell" (:333-342): move a tiny bit along the path and locate again
            atomicAdd(&ctr->stuckEscaped, 1ull);
            x += kx * eps; y += ky * eps; z += kz * eps;
            double r, theta; symSpherical(x, y, z, r, theta);
            i = locateFail(g.v1, r, g.N1 + 1);
            k = locateClip(g.v2, theta, g.N2 + 1);
        }
        alive = i < g.N1 && i >= 0;
        return seg;
    }
};

}   // namespace skg
