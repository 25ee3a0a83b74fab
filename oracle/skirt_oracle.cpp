// ORACLE -- TEST INFRASTRUCTURE ONLY (see skirt_oracle.hpp for the rules and the parity pinning).
//
// CPU restatement of the reference's DustGrid::path() family, DustGridPath and the optical-depth
// integration, working on the flattened tables of include/skirtgpu.h.  Plain scalar C++, one ray at
// a time, written from the reference text cited at every function; the product (skirt_b200/) never
// links or calls this file.
#include "skirt_oracle.hpp"

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <thread>

namespace orc
{

// =====================================================================================================
// DustGridPath
// =====================================================================================================

// DustGridPath::moveInside, DustGridPath.cpp:57-150.  box = xmin,ymin,zmin,xmax,ymax,zmax.
// The three axes are treated one after the other; each may add its own "outside" segment (m = -1).
bool Path::moveInside(const double box[6], double eps, double& x, double& y, double& z)
{
    x = rx; y = ry; z = rz;
    for (int a = 0; a < 3; a++)
    {
        double& c = (a == 0) ? x : (a == 1) ? y : z;
        const double kc = (a == 0) ? kx : (a == 1) ? ky : kz;
        const double lo = box[a], hi = box[3 + a];
        double target, edge;
        if (c <= lo) { if (kc <= 0.0) return false; edge = lo; target = lo + eps; }
        else if (c >= hi) { if (kc >= 0.0) return false; edge = hi; target = hi - eps; }
        else continue;
        double ds = (edge - c) / kc;
        add(-1, ds);
        // the coordinate that crossed is put just inside; the other two advance along the ray
        if (a != 0) x += kx * ds;
        if (a != 1) y += ky * ds;
        if (a != 2) z += kz * ds;
        c = target;
    }
    return true;
}

// DustGridPath::pathlength, DustGridPath.cpp:162-173 with NR::locate<Segment> (NR.hpp:76-95) and
// NR::interpolate_linlin (NR.hpp:296-299); Segment ordering is by tau (DustGridPath.hpp:165)
double Path::pathlength(double tau) const
{
    int N = (int)v.size();
    if (N > 0 && tau > 0)
    {
        int i;
        if (tau < v[0].tau) i = -1;
        else if (v[N - 1].tau < tau) i = N - 1;
        else
        {
            int jl = -1, ju = N;
            while (ju - jl > 1) { int jm = (ju + jl) >> 1; if (tau < v[jm].tau) ju = jm; else jl = jm; }
            i = jl <= 0 ? 0 : (jl >= N - 2 ? N - 2 : jl);
        }
        auto linlin = [](double x, double x1, double x2, double f1, double f2) { return f1 + ((x - x1) / (x2 - x1)) * (f2 - f1); };
        if (i < 0) return linlin(tau, 0, v[0].tau, 0, v[0].s);
        if (i < N - 1) return linlin(tau, v[i].tau, v[i + 1].tau, v[i].s, v[i + 1].s);
        return v[N - 1].s;
    }
    return 0;
}

// DustGridPath::fillOpticalDepth, DustGridPath.hpp:117-129
void fillOpticalDepth(Path& p, const Medium& med, int ell)
{
    double tau = 0;
    for (Seg& sg : p.v)
    {
        sg.dtau = med.kapparho(sg.m, ell) * sg.ds;
        tau += sg.dtau;
        sg.tau = tau;
    }
}

// DustGridPath::opticalDepth(kapparho, distance), DustGridPath.hpp:97-108: the segment that overshoots
// the distance is still counted in full
double opticalDepth(const Path& p, const Medium& med, int ell, double distance)
{
    double tau = 0;
    for (const Seg& sg : p.v)
    {
        tau += med.kapparho(sg.m, ell) * sg.ds;
        if (sg.s > distance) break;
    }
    return tau;
}

namespace
{

// NR::locate_basic_impl, NR.hpp:99-112
int locateBasic(const double* xv, double x, int n)
{
    int jl = -1, ju = n;
    while (ju - jl > 1) { int jm = (ju + jl) >> 1; if (x < xv[jm]) ju = jm; else jl = jm; }
    return jl;
}
int locateClip(const double* xv, int n, double x) { return x < xv[0] ? 0 : locateBasic(xv, x, n - 1); }     // NR.hpp:146-151
int locateFail(const double* xv, int n, double x) { return x > xv[n - 1] ? -1 : locateBasic(xv, x, n - 1); } // NR.hpp:155-160

// Box::contains, closed on every face (Box.hpp:94-95); b = xmin,ymin,zmin,xmax,ymax,zmax
bool inBox(const double* b, double x, double y, double z)
{
    return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5];
}

// one axis of Box::cellindices, Box.hpp:134-139.  static_cast<int> of an out-of-range double is what
// x86-64's cvttsd2si makes of it (INT_MIN), which the clamp then turns into 0.
int boxIndex(double v, double vmin, double vmax, int n)
{
    double q = n * (v - vmin) / (vmax - vmin);
    int i = (q >= 2147483648.0 || q <= -2147483649.0 || q != q) ? INT_MIN : static_cast<int>(q);
    return std::max(0, std::min(n - 1, i));
}

double nextAlong(double v, double k) { return std::nextafter(v, (k < 0.0) ? -DBL_MAX : DBL_MAX); }

double epsOf(const double* b)   // 1e-12 * extent.widths().norm(): TreeDustGrid.cpp:76, VoronoiMesh.cpp:234, AdaptiveMesh.cpp:52
{
    double wx = b[3] - b[0], wy = b[4] - b[1], wz = b[5] - b[2];
    return 1e-12 * std::sqrt(wx * wx + wy * wy + wz * wz);
}

// exit distances towards the walls of a box along k (TreeDustGrid.cpp:416-421, AdaptiveMesh.cpp:317-322)
struct Exit { double dsx, dsy, dsz, xnext, ynext, znext; };
Exit exits(const double* b, double x, double y, double z, double kx, double ky, double kz)
{
    Exit e;
    e.xnext = (kx < 0.0) ? b[0] : b[3];
    e.ynext = (ky < 0.0) ? b[1] : b[4];
    e.znext = (kz < 0.0) ? b[2] : b[5];
    e.dsx = (std::fabs(kx) > 1e-15) ? (e.xnext - x) / kx : DBL_MAX;
    e.dsy = (std::fabs(ky) > 1e-15) ? (e.ynext - y) / ky : DBL_MAX;
    e.dsz = (std::fabs(kz) > 1e-15) ? (e.znext - z) / kz : DBL_MAX;
    return e;
}

// =====================================================================================================
// CartesianDustGrid
// =====================================================================================================
struct Cartesian : Grid
{
    std::vector<double> xv, yv, zv; int Nx, Ny, Nz;
    double xmin, xmax, ymin, ymax, zmin, zmax;
    int numCells() const override { return Nx * Ny * Nz; }
    int index(int i, int j, int k) const { return k + Nz * j + Nz * Ny * i; }       // CartesianDustGrid.cpp:326-329

    // CartesianDustGrid::whichcell, CartesianDustGrid.cpp:109-118
    int whichcell(double x, double y, double z) const override
    {
        int i = locateFail(xv.data(), Nx + 1, x), j = locateFail(yv.data(), Ny + 1, y), k = locateFail(zv.data(), Nz + 1, z);
        if (i < 0 || j < 0 || k < 0) return -1;
        return index(i, j, k);
    }

    // CartesianDustGrid::path, CartesianDustGrid.cpp:136-283
    void path(Path& p) const override
    {
        p.clear();
        double x = p.rx, y = p.ry, z = p.rz; const double kx = p.kx, ky = p.ky, kz = p.kz;
        double ds;
        // entry from outside, :151-222 (strict comparisons, 1e-8 of the outermost bin width as nudge)
        if (x < xmin) { if (kx <= 0.0) return p.clear(); ds = (xmin - x) / kx; p.add(-1, ds); x = xmin + 1e-8 * (xv[1] - xv[0]); y += ky * ds; z += kz * ds; }
        else if (x > xmax) { if (kx >= 0.0) return p.clear(); ds = (xmax - x) / kx; p.add(-1, ds); x = xmax - 1e-8 * (xv[Nx] - xv[Nx - 1]); y += ky * ds; z += kz * ds; }
        if (y < ymin) { if (ky <= 0.0) return p.clear(); ds = (ymin - y) / ky; p.add(-1, ds); x += kx * ds; y = ymin + 1e-8 * (yv[1] - yv[0]); z += kz * ds; }
        else if (y > ymax) { if (ky >= 0.0) return p.clear(); ds = (ymax - y) / ky; p.add(-1, ds); x += kx * ds; y = ymax - 1e-8 * (yv[Ny] - yv[Ny - 1]); z += kz * ds; }
        if (z < zmin) { if (kz <= 0.0) return p.clear(); ds = (zmin - z) / kz; p.add(-1, ds); x += kx * ds; y += ky * ds; z = zmin + 1e-8 * (zv[1] - zv[0]); }
        else if (z > zmax) { if (kz >= 0.0) return p.clear(); ds = (zmax - z) / kz; p.add(-1, ds); x += kx * ds; y += ky * ds; z = zmax - 1e-8 * (zv[Nz] - zv[Nz - 1]); }
        if (x < xmin || x > xmax || y < ymin || y > ymax || z < zmin || z > zmax) return p.clear();      // :224

        int i = locateClip(xv.data(), Nx + 1, x), j = locateClip(yv.data(), Ny + 1, y), k = locateClip(zv.data(), Nz + 1, z);
        if (!std::isfinite(x + y + z + kx + ky + kz)) return p.clear();     // the reference would loop forever on NaN input
        while (true)    // :234-282
        {
            int m = index(i, j, k);
            double xE = (kx < 0.0) ? xv[i] : xv[i + 1];
            double yE = (ky < 0.0) ? yv[j] : yv[j + 1];
            double zE = (kz < 0.0) ? zv[k] : zv[k + 1];
            double dsx = (std::fabs(kx) > 1e-15) ? (xE - x) / kx : DBL_MAX;
            double dsy = (std::fabs(ky) > 1e-15) ? (yE - y) / ky : DBL_MAX;
            double dsz = (std::fabs(kz) > 1e-15) ? (zE - z) / kz : DBL_MAX;
            if (dsx <= dsy && dsx <= dsz)
            {
                p.add(m, dsx);
                i += (kx < 0.0) ? -1 : 1;
                if (i >= Nx || i < 0) return;
                x = xE; y += ky * dsx; z += kz * dsx;
            }
            else if (dsy < dsx && dsy <= dsz)
            {
                p.add(m, dsy);
                j += (ky < 0.0) ? -1 : 1;
                if (j >= Ny || j < 0) return;
                x += kx * dsy; y = yE; z += kz * dsy;
            }
            else if (dsz < dsx && dsz < dsy)
            {
                p.add(m, dsz);
                k += (kz < 0.0) ? -1 : 1;
                if (k >= Nz || k < 0) return;
                x += kx * dsz; y += ky * dsz; z = zE;
            }
            else return;
        }
    }

    // CartesianDustGrid::randomPositionInCell, CartesianDustGrid.cpp:129-132 -> Random::position(Box), Random.cpp:226-234
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
};

// =====================================================================================================
// TreeDustGrid (octree / binary tree), nodes flattened in id order
// =====================================================================================================
struct Tree : Grid
{
    int kind, search, N, ncells = 0;
    std::vector<double> box; std::vector<int> child0, parent, cell, dir, nbrStart, nbrIds;
    double eps;
    int numCells() const override { return ncells; }
    const double* b(int l) const { return box.data() + 6 * (size_t)l; }

    // TreeNode::whichnode(Vec) from the root, TreeNode.cpp:70-80; OctTreeNode::child (OctTreeNode.cpp:184-189)
    // compares with the max corner of child 0, BinTreeNode::child (BinTreeNode.cpp:326-335) with its max along _dir
    int whichnode(double x, double y, double z) const
    {
        if (!inBox(b(0), x, y, z)) return -1;
        int node = 0;
        while (child0[node] >= 0)
        {
            int c0 = child0[node]; const double* cb = b(c0);
            if (kind == 0) node = c0 + ((x < cb[3]) ? 0 : 1) + ((y < cb[4]) ? 0 : 2) + ((z < cb[5]) ? 0 : 4);
            else
            {
                int d = dir[node];
                double v = d == 0 ? x : (d == 1 ? y : z);
                node = (v < cb[3 + d]) ? c0 : c0 + 1;
            }
        }
        return node;
    }
    int whichcell(double x, double y, double z) const override { int n = whichnode(x, y, z); return n >= 0 ? cell[n] : -1; }

    // TreeDustGrid::path, TreeDustGrid.cpp:390-662
    void path(Path& p) const override
    {
        p.clear();
        double x, y, z;
        const double kx = p.kx, ky = p.ky, kz = p.kz;
        if (!std::isfinite(p.rx + p.ry + p.rz + kx + ky + kz)) return;
        if (!p.moveInside(b(0), eps, x, y, z)) return p.clear();
        int node = whichnode(x, y, z);
        if (node < 0) return p.clear();

        if (search == 0 || search == 1)
        {
            // TopDown :412-456 and Neighbor :460-521 differ only in how the next node is looked up
            while (node >= 0)
            {
                Exit e = exits(b(node), x, y, z, kx, ky, kz);
                double ds; int wall;    // walls BACK,FRONT,LEFT,RIGHT,BOTTOM,TOP = 0..5 (TreeNode.hpp:100)
                if (e.dsx <= e.dsy && e.dsx <= e.dsz) { ds = e.dsx; wall = (kx < 0.0) ? 0 : 1; }
                else if (e.dsy <= e.dsx && e.dsy <= e.dsz) { ds = e.dsy; wall = (ky < 0.0) ? 2 : 3; }
                else { ds = e.dsz; wall = (kz < 0.0) ? 4 : 5; }
                p.add(cell[node], ds);
                x += (ds + eps) * kx; y += (ds + eps) * ky; z += (ds + eps) * kz;

                int old = node;
                if (search == 1)
                {
                    // TreeNode::whichnode(wall, r), TreeNode.cpp:84-93: first listed neighbour whose closed box holds r
                    node = -1;
                    for (int q = nbrStart[6 * (size_t)old + wall]; q < nbrStart[6 * (size_t)old + wall + 1]; q++)
                        if (inBox(b(nbrIds[q]), x, y, z)) { node = nbrIds[q]; break; }
                    if (node < 0) node = whichnode(x, y, z);
                }
                else node = whichnode(x, y, z);

                if (node == old)    // :437-454 / :502-519
                {
                    stuck++;
                    x = nextAlong(x, kx); y = nextAlong(y, ky); z = nextAlong(z, kz);
                    node = whichnode(x, y, z);
                    if (node == old) { stuck++; break; }
                }
            }
            return;
        }

        if (search == 3)
        {
            // ParticleTreeDustGrid::path, ParticleTreeDustGrid.cpp:258-325 (search tag 3 in the tables): the nearest wall with a
            // positive distance (plain divisions: a zero direction component gives +-inf or nan, which the comparisons discard),
            // no segment when there is none, the next node always looked up from the root
            while (node >= 0)
            {
                const double* nb = b(node);
                const double dsx = (((kx < 0.0) ? nb[0] : nb[3]) - x) / kx;
                const double dsy = (((ky < 0.0) ? nb[1] : nb[4]) - y) / ky;
                const double dsz = (((kz < 0.0) ? nb[2] : nb[5]) - z) / kz;
                double ds = DBL_MAX;
                if (dsx > 0 && dsx < ds) ds = dsx;
                if (dsy > 0 && dsy < ds) ds = dsy;
                if (dsz > 0 && dsz < ds) ds = dsz;
                if (ds < DBL_MAX) p.add(cell[node], ds); else ds = 0;
                x += (ds + eps) * kx; y += (ds + eps) * ky; z += (ds + eps) * kz;
                const int old = node;
                node = whichnode(x, y, z);
                if (node == old)
                {
                    stuck++;
                    x = nextAlong(x, kx); y = nextAlong(y, ky); z = nextAlong(z, kz);
                    node = whichnode(x, y, z);
                    if (node == old) { stuck++; break; }
                }
            }
            return;
        }

        // Bookkeeping (octree only), :527-659: children of a node have ids 8q+1 .. 8q+8 relative order, so the
        // octant of node l within its father is ((l-1)%8)
        int l = node;
        while (true)
        {
            Exit e = exits(b(l), x, y, z, kx, ky, kz);
            if (e.dsx <= e.dsy && e.dsx <= e.dsz)
            {
                p.add(cell[l], e.dsx);
                x = e.xnext; y += ky * e.dsx; z += kz * e.dsx;
                while (true)
                {
                    int oct = ((l - 1) % 8) + 1;
                    bool place = (kx < 0.0) ? (oct % 2 == 1) : (oct % 2 == 0);
                    if (!place) break;
                    l = parent[l];
                    if (l == 0) return;
                }
                l += (kx < 0.0) ? -1 : 1;
                while (cell[l] == -1)
                {
                    int c0 = child0[l]; double yM = b(c0)[4], zM = b(c0)[5];
                    int base = (kx < 0.0) ? 1 : 0;
                    l = c0 + base + ((y <= yM) ? 0 : 2) + ((z <= zM) ? 0 : 4);
                }
            }
            else if (e.dsy < e.dsx && e.dsy <= e.dsz)
            {
                p.add(cell[l], e.dsy);
                x += kx * e.dsy; y = e.ynext; z += kz * e.dsy;
                while (true)
                {
                    bool place = (ky < 0.0) ? ((l - 1) % 4 < 2) : ((l - 1) % 4 > 1);
                    if (!place) break;
                    l = parent[l];
                    if (l == 0) return;
                }
                l += (ky < 0.0) ? -2 : 2;
                while (cell[l] == -1)
                {
                    int c0 = child0[l]; double xM = b(c0)[3], zM = b(c0)[5];
                    int base = (ky < 0.0) ? 2 : 0;
                    l = c0 + base + ((x <= xM) ? 0 : 1) + ((z <= zM) ? 0 : 4);
                }
            }
            else if (e.dsz < e.dsx && e.dsz < e.dsy)
            {
                p.add(cell[l], e.dsz);
                x += kx * e.dsz; y += ky * e.dsz; z = e.znext;
                while (true)
                {
                    int oct = ((l - 1) % 8) + 1;
                    bool place = (kz < 0.0) ? (oct < 5) : (oct > 4);
                    if (!place) break;
                    l = parent[l];
                    if (l == 0) return;
                }
                l += (kz < 0.0) ? -4 : 4;
                while (cell[l] == -1)
                {
                    int c0 = child0[l]; double xM = b(c0)[3], yM = b(c0)[4];
                    int base = (kz < 0.0) ? 4 : 0;
                    l = c0 + base + ((x <= xM) ? 0 : 1) + ((y <= yM) ? 0 : 2);
                }
            }
            else return;
        }
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
    std::vector<int> leafOfCell;
};

// =====================================================================================================
// AdaptiveMesh
// =====================================================================================================
struct AMesh : Grid
{
    int N, ncells = 0;
    std::vector<double> box; std::vector<int> nxyz, child0, cell, wallNbr;
    double eps;
    mutable long errors = 0;
    int numCells() const override { return ncells; }
    const double* b(int l) const { return box.data() + 6 * (size_t)l; }

    // AdaptiveMeshNode::whichnode(Vec) :132-142 with child() :109-128; -2 stands for the FATALERROR
    int whichnode(double x, double y, double z) const
    {
        if (!inBox(b(0), x, y, z)) return -1;
        int node = 0;
        while (child0[node] >= 0)
        {
            const double* nb = b(node);
            int Nx = nxyz[3 * node], Ny = nxyz[3 * node + 1], Nz = nxyz[3 * node + 2];
            int i = boxIndex(x, nb[0], nb[3], Nx), j = boxIndex(y, nb[1], nb[4], Ny), k = boxIndex(z, nb[2], nb[5], Nz);
            int c = child0[node] + (k * Ny + j) * Nx + i;
            if (!inBox(b(c), x, y, z))
            {
                const double* cb = b(c);
                if (x < cb[0]) i--; else if (x > cb[3]) i++;
                if (y < cb[1]) j--; else if (y > cb[4]) j++;
                if (z < cb[2]) k--; else if (z > cb[5]) k++;
                if (i < 0 || i >= Nx || j < 0 || j >= Ny || k < 0 || k >= Nz) return -2;
                c = child0[node] + (k * Ny + j) * Nx + i;
                if (!inBox(b(c), x, y, z)) return -2;
            }
            node = c;
        }
        return node;
    }
    int whichcell(double x, double y, double z) const override { int n = whichnode(x, y, z); return n >= 0 ? cell[n] : -1; }

    // AdaptiveMesh::path, AdaptiveMesh.cpp:297-367
    void path(Path& p) const override
    {
        p.clear();
        double x, y, z;
        const double kx = p.kx, ky = p.ky, kz = p.kz;
        if (!std::isfinite(p.rx + p.ry + p.rz + kx + ky + kz)) return;
        if (!p.moveInside(b(0), eps, x, y, z)) return p.clear();
        int node = whichnode(x, y, z);
        if (node < 0) { if (node == -2) errors++; return p.clear(); }
        while (node >= 0)
        {
            Exit e = exits(b(node), x, y, z, kx, ky, kz);
            double ds; int wall;
            if (e.dsx <= e.dsy && e.dsx <= e.dsz) { ds = e.dsx; wall = (kx < 0.0) ? 0 : 1; }
            else if (e.dsy <= e.dsx && e.dsy <= e.dsz) { ds = e.dsy; wall = (ky < 0.0) ? 2 : 3; }
            else { ds = e.dsz; wall = (kz < 0.0) ? 4 : 5; }
            p.add(cell[node], ds);
            x += (ds + eps) * kx; y += (ds + eps) * ky; z += (ds + eps) * kz;      // r += (ds+_eps)*k

            // AdaptiveMeshNode::whichnode(wall, r) :146-151: the single stored neighbour if its box holds r
            int old = node;
            int cand = wallNbr[6 * (size_t)old + wall];
            node = (cand >= 0 && inBox(b(cand), x, y, z)) ? cand : whichnode(x, y, z);
            if (node == -2) { errors++; return; }
            if (node == old)
            {
                stuck++;
                x = nextAlong(x, kx); y = nextAlong(y, ky); z = nextAlong(z, kz);
                node = whichnode(x, y, z);
                if (node == -2) { errors++; return; }
                if (node == old) { stuck++; break; }
            }
        }
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
    std::vector<int> leafOfCell;
};

// =====================================================================================================
// VoronoiMesh
// =====================================================================================================
struct Voronoi : Grid
{
    int N, nb;
    std::vector<double> part, cellBox; std::vector<int> nbrStart, nbrIds, blkStart, blkIds, blkTree, kdM, kdAxis, kdUp, kdLeft, kdRight;
    double ext[6];      // xmin,ymin,zmin,xmax,ymax,zmax
    double eps;
    mutable long errors = 0;
    int numCells() const override { return N; }
    const double* pt(int m) const { return part.data() + 3 * (size_t)m; }
    double sd(int m, double x, double y, double z) const    // VoronoiCell::squaredDistanceTo, VoronoiMesh.cpp:63-64
    { const double* q = pt(m); double dx = x - q[0], dy = y - q[1], dz = z - q[2]; return dx * dx + dy * dy + dz * dz; }

    // lessthan(p1,p2,axis), VoronoiMesh.cpp:77-105: lexicographic with the split axis first, cyclic order after it
    static bool lessthan(const double* p1, const double* p2, int axis)
    {
        for (int t = 0; t < 3; t++)
        {
            int a = (axis + t) % 3;
            if (p1[a] < p2[a]) return true;
            if (t < 2 && p1[a] > p2[a]) return false;
        }
        return false;
    }

    // Node::nearest, VoronoiMesh.cpp:180-225 (recursive like the reference)
    int nearest(int top, const double* r) const
    {
        int current = top;
        while (true)
        {
            int child = lessthan(r, pt(kdM[current]), kdAxis[current]) ? kdLeft[current] : kdRight[current];
            if (child < 0) break;
            current = child;
        }
        int best = current;
        double bestSD = sd(kdM[best], r[0], r[1], r[2]);
        while (true)
        {
            double currentSD = sd(kdM[current], r[0], r[1], r[2]);
            if (currentSD < bestSD) { best = current; bestSD = currentSD; }
            double d = pt(kdM[current])[kdAxis[current]] - r[kdAxis[current]];
            double splitSD = d * d;
            if (splitSD < bestSD)
            {
                int other = lessthan(r, pt(kdM[current]), kdAxis[current]) ? kdRight[current] : kdLeft[current];
                if (other >= 0)
                {
                    int otherBest = nearest(other, r);
                    double otherSD = sd(kdM[otherBest], r[0], r[1], r[2]);
                    if (otherSD < bestSD) { best = otherBest; bestSD = otherSD; }
                }
            }
            if (current == top) break;
            current = kdUp[current];
        }
        return best;
    }

    // VoronoiMesh::cellIndex, VoronoiMesh.cpp:512-541
    int whichcell(double x, double y, double z) const override
    {
        if (!inBox(ext, x, y, z)) return -1;
        int i = boxIndex(x, ext[0], ext[3], nb), j = boxIndex(y, ext[1], ext[4], nb), k = boxIndex(z, ext[2], ext[5], nb);
        size_t blk = ((size_t)i * nb + j) * nb + k;
        if (blkTree[blk] >= 0) { double r[3] = {x, y, z}; return kdM[nearest(blkTree[blk], r)]; }
        int m = -1; double mdist = DBL_MAX;
        for (int q = blkStart[blk]; q < blkStart[blk + 1]; q++)
        {
            double idist = sd(blkIds[q], x, y, z);
            if (idist < mdist) { m = blkIds[q]; mdist = idist; }
        }
        return m;
    }

    // VoronoiMesh::path, VoronoiMesh.cpp:749-844
    void path(Path& p) const override
    {
        p.clear();
        double x, y, z;
        const double kx = p.kx, ky = p.ky, kz = p.kz;
        if (!std::isfinite(p.rx + p.ry + p.rz + kx + ky + kz)) return;
        if (!p.moveInside(ext, eps, x, y, z)) return p.clear();
        int mr = whichcell(x, y, z);
        if (mr < 0) return p.clear();
        long guard = 0;
        while (mr >= 0)
        {
            const double* pr = pt(mr);
            double sq = DBL_MAX; const int NO_INDEX = -99; int mq = NO_INDEX;
            for (int q = nbrStart[mr]; q < nbrStart[mr + 1]; q++)
            {
                int mi = nbrIds[q];
                double si = 0;
                if (mi >= 0)
                {
                    const double* pi = pt(mi);
                    double nx = pi[0] - pr[0], ny = pi[1] - pr[1], nz = pi[2] - pr[2];
                    double ndotk = nx * kx + ny * ky + nz * kz;
                    if (ndotk > 0)
                    {
                        double px = 0.5 * (pi[0] + pr[0]), py = 0.5 * (pi[1] + pr[1]), pz = 0.5 * (pi[2] + pr[2]);
                        si = (nx * (px - x) + ny * (py - y) + nz * (pz - z)) / ndotk;
                    }
                }
                else if (mi >= -6)
                {
                    // walls -1..-6 = xmin,xmax,ymin,ymax,zmin,zmax (:810-817)
                    int a = (-mi - 1) / 2; bool upper = ((-mi - 1) % 2) == 1;
                    double c = a == 0 ? x : (a == 1 ? y : z), kc = a == 0 ? kx : (a == 1 ? ky : kz);
                    si = (ext[a + (upper ? 3 : 0)] - c) / kc;
                }
                else { errors++; return; }
                if (si > 0 && si < sq) { sq = si; mq = mi; }
            }
            if (mq == NO_INDEX)
            {
                x += kx * eps; y += ky * eps; z += kz * eps;
                mr = whichcell(x, y, z);
                if (++guard > 1000000) { errors++; return; }
            }
            else
            {
                p.add(mr, sq);
                x += (sq + eps) * kx; y += (sq + eps) * ky; z += (sq + eps) * kz;
                mr = mq;
            }
        }
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
};

// =====================================================================================================
// Sphere1DDustGrid: concentric shells, borders rv[0..Nr] (Sphere1DDustGrid.cpp:24-33)
// =====================================================================================================
struct Sphere1D : Grid
{
    std::vector<double> rv; int Nr;
    int numCells() const override { return Nr; }
    // Sphere1DDustGrid::whichcell, Sphere1DDustGrid.cpp:85-88
    int whichcell(double x, double y, double z) const override { return locateFail(rv.data(), Nr + 1, std::sqrt(x * x + y * y + z * z)); }

    // Sphere1DDustGrid::path, Sphere1DDustGrid.cpp:111-185.  The ray is described by its impact parameter p and the signed
    // distance q to the point of closest approach; shell border rN is crossed at q = -/+ sqrt((rN-p)(rN+p)).  Kept as the
    // reference has them: a ray from outside continues from q = +qmax (the far side), and the outward walk ends when it
    // reaches shell Nr-1.
    void path(Path& P) const override
    {
        P.clear();
        const double x = P.rx, y = P.ry, z = P.rz, rmax = rv[Nr];
        double r = std::sqrt(x * x + y * y + z * z);
        double q = x * P.kx + y * P.ky + z * P.kz;
        const double p = std::sqrt((r - q) * (r + q));
        auto cross = [&](double rN) { return std::sqrt((rN - p) * (rN + p)); };
        if (r > rmax)
        {
            if (q > 0.0 || p > rmax) { P.clear(); return; }
            r = rmax - 1e-8 * (rv[Nr] - rv[Nr - 1]);
            const double qmax = cross(rmax);
            P.add(-1, qmax - q);
            q = qmax;
        }
        int i = locateClip(rv.data(), Nr + 1, r);
        if (q < 0.0)        // inward, down to the shell that holds the point of closest approach
        {
            const int imin = locateClip(rv.data(), Nr + 1, p);
            while (i > imin) { const double qN = -cross(rv[i]); P.add(i, qN - q); q = qN; i--; }
        }
        for (;;)            // outward
        {
            const double qN = cross(rv[i + 1]);
            P.add(i, qN - q);
            if (++i >= Nr - 1) return;
            q = qN;
        }
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
};

// =====================================================================================================
// Sphere2DDustGrid: shells x polar bins, borders rv[0..Nr], thetav[0..Nt], cv = cos(thetav) with exact 1, 0, -1 at the
// poles and in the xy-plane (Sphere2DDustGrid.cpp:27-75); m = k + Nt*i (:387-390)
// =====================================================================================================
// smallest positive root of x^2 + 2 b x + c = 0, 0 if none (Sphere2DDustGrid.cpp:188-215): the root of smaller magnitude is
// taken as c / (the other root) so that it does not suffer cancellation
double firstRoot(double b, double c)
{
    if (!(b * b > c)) return 0;
    if (b > 0)
    {
        if (c < 0) { const double x1 = -b - std::sqrt(b * b - c); return c / x1; }
        return 0;
    }
    const double x2 = -b + std::sqrt(b * b - c);
    if (c > 0) { const double x1 = c / x2; if (x1 < x2) return x1; }
    return x2;
}
// the same for a x^2 + 2 b x + c = 0 (Sphere2DDustGrid.cpp:218-224)
double firstRoot(double a, double b, double c)
{
    if (std::fabs(a) > 1e-9) return firstRoot(b / a, c / a);
    const double x = -0.5 * c / b;
    return x > 0 ? x : 0;
}

struct Sphere2D : Grid
{
    std::vector<double> rv, thetav, cv; int Nr, Nt;
    int numCells() const override { return Nr * Nt; }
    // Position::spherical, Position.cpp:95-108 (phi is not needed)
    static void spherical(double x, double y, double z, double& r, double& theta)
    {
        r = std::sqrt(x * x + y * y + z * z);
        theta = r == 0 ? 0.0 : std::acos(z / r);
    }
    // Sphere2DDustGrid::whichcell, Sphere2DDustGrid.cpp:148-156
    int whichcell(double x, double y, double z) const override
    {
        double r, theta; spherical(x, y, z, r, theta);
        const int i = locateFail(rv.data(), Nr + 1, r);
        if (i < 0) return -1;
        return locateClip(thetav.data(), Nt + 1, theta) + Nt * i;
    }

    // Sphere2DDustGrid::path, Sphere2DDustGrid.cpp:249-358: from the current point the nearest crossing with the two spheres and
    // the two cones of the cell (Sphere2DDustGrid.cpp:228-243) decides the next cell; the point is pushed eps beyond each wall
    void path(Path& P) const override
    {
        const double rmax = rv[Nr], eps = 1e-11 * rmax;
        P.clear();
        double x = P.rx, y = P.ry, z = P.rz; const double kx = P.kx, ky = P.ky, kz = P.kz;
        auto advance = [&](double d) { x += kx * d; y += ky * d; z += kz * d; };
        auto sphere = [&](double rad) { return firstRoot(x * kx + y * ky + z * kz, (x * x + y * y + z * z) - rad * rad); };
        auto cone = [&](double c)
        {
            return c ? firstRoot(c * c - kz * kz, c * c * (x * kx + y * ky + z * kz) - z * kz, c * c * (x * x + y * y + z * z) - z * z)
                     : -z / kz;
        };
        const double r2 = x * x + y * y + z * z;
        if (r2 > rmax * rmax)
        {
            const double ds = sphere(rmax);
            if (!ds) { P.clear(); return; }
            P.add(-1, ds);
            advance(ds + eps);
        }
        else if (r2 == 0) advance(eps);

        double r, theta; spherical(x, y, z, r, theta);
        int i = locateFail(rv.data(), Nr + 1, r);
        int k = locateClip(thetav.data(), Nt + 1, theta);
        int inext = i, knext = k;
        while (i < Nr && i >= 0)
        {
            double ds = DBL_MAX;
            auto candidate = [&](double s, int in, int kn) { if (s > 0 && s < ds) { ds = s; inext = in; knext = kn; } };
            if (i > 0) candidate(sphere(rv[i]), i - 1, k);
            candidate(sphere(rv[i + 1]), i + 1, k);
            if (k > 0) candidate(cone(cv[k]), i, k - 1);
            if (k < Nt - 1) candidate(cone(cv[k + 1]), i, k + 1);
            if (inext != i || knext != k)
            {
                P.add(k + Nt * i, ds);
                advance(ds + eps);
                i = inext; k = knext;
            }
            else        // "No exit point found from dust grid cell": a tiny step, the cell looked up again
            {
                stuck++;
                advance(eps);
                spherical(x, y, z, r, theta);
                i = locateFail(rv.data(), Nr + 1, r);
                k = locateClip(thetav.data(), Nt + 1, theta);
            }
        }
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
};

// =====================================================================================================
// Cylinder2DDustGrid: rings R x z, borders Rv[0..NR], zv[0..Nz]; m = k + Nz*i (Cylinder2DDustGrid.cpp:391-394)
// =====================================================================================================
struct Cylinder2D : Grid
{
    std::vector<double> Rv, zv; int NR, Nz;
    int numCells() const override { return NR * Nz; }
    // Cylinder2DDustGrid::whichcell, Cylinder2DDustGrid.cpp:101-107
    int whichcell(double x, double y, double z) const override
    {
        const int i = locateFail(Rv.data(), NR + 1, std::sqrt(x * x + y * y));
        const int k = locateFail(zv.data(), Nz + 1, z);
        return (i < 0 || k < 0) ? -1 : k + Nz * i;
    }

    // Cylinder2DDustGrid::path, Cylinder2DDustGrid.cpp:135-385.  In the projection on the equatorial plane the ray is a line with
    // impact parameter p and abscissa q (speed kq); the reference's four loops (up / down x inward / outward) differ only in
    // which border is next on each axis, so they are one loop here with the two directions as parameters.
    void path(Path& P) const override
    {
        P.clear();
        const double kx = P.kx, ky = P.ky; double kz = P.kz;
        double kq = std::sqrt(kx * kx + ky * ky);
        if (kz == 0.0) kz = 1e-20;
        if (kq == 0.0) kq = 1e-20;
        const double x = P.rx, y = P.ry; double z = P.rz;
        double R = std::sqrt(x * x + y * y);
        double q = (x * kx + y * ky) / kq;
        const double p2 = (R - q) * (R + q);
        const double p = std::sqrt(std::max(0.0, p2));
        const double Rmax = Rv[NR], zmin = zv[0], zmax = zv[Nz];
        auto cross = [&](double RN) { return std::sqrt((RN - p) * (RN + p)); };

        if (R >= Rmax)
        {
            if (q > 0.0 || p > Rmax) { P.clear(); return; }
            R = Rmax - 1e-8 * (Rv[NR] - Rv[NR - 1]);
            const double qmax = cross(Rmax);
            const double ds = (qmax - q) / kq;
            P.add(-1, ds);
            q = qmax;
            z += kz * ds;
        }
        if (z < zmin)
        {
            if (kz <= 0.0) { P.clear(); return; }
            const double ds = (zmin - z) / kz;
            P.add(-1, ds);
            q += kq * ds;
            R = std::sqrt(p * p + q * q);
            z = zmin + 1e-8 * (zv[1] - zv[0]);
        }
        else if (z > zmax)
        {
            if (kz >= 0.0) { P.clear(); return; }
            const double ds = (zmax - z) / kz;
            P.add(-1, ds);
            q += kq * ds;
            R = std::sqrt(p * p + q * q);
            z = zmax - 1e-8 * (zv[Nz] - zv[Nz - 1]);
        }
        if (std::isinf(R) || std::isnan(R) || std::isinf(z) || std::isnan(z) || R >= Rmax || z <= zmin || z >= zmax) { P.clear(); return; }

        int i = locateClip(Rv.data(), NR + 1, R);
        int k = locateClip(zv.data(), Nz + 1, z);
        const bool up = kz >= 0.0;
        const int dk = up ? 1 : -1, zside = up ? 1 : 0;
        // the reference ends a downward path one ring early (i >= NR-1, Cylinder2DDustGrid.cpp:338; upward: i >= NR, :262)
        const int iend = up ? NR : NR - 1;
        // one step of either phase: returns false when the path has left the grid through the top, the bottom or the outer wall
        auto walk = [&](bool inward, double qN) -> bool
        {
            const double zN = zv[k + zside];
            const double dsq = (qN - q) / kq, dsz = (zN - z) / kz;
            const int m = k + Nz * i;
            if (dsq < dsz)
            {
                P.add(m, dsq);
                if (inward) i--; else if (++i >= iend) return false;
                q = qN;
                z += kz * dsq;
            }
            else
            {
                P.add(m, dsz);
                k += dk;
                if (k >= Nz || k < 0) return false;
                q += kq * dsz;
                z = zN;
            }
            return true;
        };
        if (q < 0.0)
        {
            const int imin = locateClip(Rv.data(), NR + 1, p);
            while (i > imin) if (!walk(true, -cross(Rv[i]))) return;
        }
        while (walk(false, cross(Rv[i + 1]))) {}
    }
    void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const override;
};

}   // anonymous namespace

// =====================================================================================================
// MT19937 stream, Random.cpp:41-126 (the reference keeps one per thread; seed + thread index)
// =====================================================================================================
struct Rng
{
    unsigned long mt[624]; int mti;
    explicit Rng(unsigned long seed)
    {
        // Random::initialize, Random.cpp:60-85: Knuth's LCG 69069 seeding of the state vector
        mt[0] = seed & 0xffffffffUL;
        for (mti = 1; mti < 624; mti++) mt[mti] = (69069 * mt[mti - 1]) & 0xffffffffUL;
    }
    double uniform()    // Random::uniform, Random.cpp:89-126: (0,1) open interval
    {
        const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL, A = 0x9908b0dfUL;
        double ans = 0.0;
        do
        {
            unsigned long y;
            if (mti >= 624)
            {
                int kk;
                for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1) ? A : 0); }
                for (; kk < 623; kk++) { y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1) ? A : 0); }
                y = (mt[623] & UPPER) | (mt[0] & LOWER); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1) ? A : 0);
                mti = 0;
            }
            y = mt[mti++];
            y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680UL; y ^= (y << 15) & 0xefc60000UL; y ^= (y >> 18);
            ans = ((double)y) / ((unsigned long)0xffffffff);
        }
        while (ans <= 0.0 || ans >= 1.0);
        return ans;
    }
};

namespace
{
// Random::position(Box), Random.cpp:226-234: x, y, z drawn in this order through Box::fracpos (Box.hpp:125-126)
void randomInBox(const double* b, Rng& rng, double& x, double& y, double& z)
{
    double fx = rng.uniform(), fy = rng.uniform(), fz = rng.uniform();
    x = b[0] + fx * (b[3] - b[0]); y = b[1] + fy * (b[4] - b[1]); z = b[2] + fz * (b[5] - b[2]);
}
void Cartesian::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const
{
    int i = m / (Nz * Ny), j = (m / Nz) % Ny, k = m % Nz;       // CartesianDustGrid::invertindex, :333-343
    double bb[6] = {xv[i], yv[j], zv[k], xv[i + 1], yv[j + 1], zv[k + 1]};
    randomInBox(bb, rng, x, y, z);
}
void Tree::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const { randomInBox(b(leafOfCell[m]), rng, x, y, z); }     // TreeDustGrid.cpp:383-386
void AMesh::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const { randomInBox(b(leafOfCell[m]), rng, x, y, z); }    // AdaptiveMesh.cpp:163-167
void Voronoi::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const
{
    // VoronoiMesh::randomPosition, VoronoiMesh.cpp:591-618: rejection in the cell's enclosing box
    for (int i = 0; i < 10000; i++)
    {
        randomInBox(cellBox.data() + 6 * (size_t)m, rng, x, y, z);
        double target = sd(m, x, y, z); bool closest = true;
        for (int q = nbrStart[m]; q < nbrStart[m + 1] && closest; q++)
            if (nbrIds[q] >= 0 && sd(nbrIds[q], x, y, z) < target) closest = false;
        if (closest) return;
    }
    errors++;
}
// Random::direction(), Random.cpp:179-184 with Direction(theta, phi), Direction.cpp:12-40
void randomDirection(Rng& rng, double& kx, double& ky, double& kz)
{
    const double theta = std::acos(2.0 * rng.uniform() - 1.0);
    const double phi = 2.0 * M_PI * rng.uniform();
    const double eps = 1e-8;
    if (theta <= eps) { kx = ky = 0.0; kz = 1.0; }
    else if (theta >= M_PI - eps) { kx = ky = 0.0; kz = -1.0; }
    else { const double sintheta = std::sin(theta); kx = sintheta * std::cos(phi); ky = sintheta * std::sin(phi); kz = std::cos(theta); }
}
// Sphere1DDustGrid::randomPositionInCell, Sphere1DDustGrid.cpp:100-106: the direction first, then the radius
void Sphere1D::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const
{
    double kx, ky, kz; randomDirection(rng, kx, ky, kz);
    const double r = rv[m] + (rv[m + 1] - rv[m]) * rng.uniform();
    x = r * kx; y = r * ky; z = r * kz;
}
// Sphere2DDustGrid::randomPositionInCell, Sphere2DDustGrid.cpp:172-183 (r, theta, phi in this order; Position.cpp:32-45)
void Sphere2D::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const
{
    const int i = m / Nt, k = m % Nt;
    const double ris = rv[i] * rv[i], ri1s = rv[i + 1] * rv[i + 1];
    const double r = std::sqrt(ris + (ri1s - ris) * rng.uniform());
    const double theta = thetav[k] + (thetav[k + 1] - thetav[k]) * rng.uniform();
    const double phi = 2.0 * M_PI * rng.uniform();
    const double costheta = std::cos(theta), sintheta = std::sin(theta), cosphi = std::cos(phi), sinphi = std::sin(phi);
    x = r * sintheta * cosphi; y = r * sintheta * sinphi; z = r * costheta;
}
// Cylinder2DDustGrid::randomPositionInCell, Cylinder2DDustGrid.cpp:123-131 (R, phi, z in this order; Position.cpp:23-30)
void Cylinder2D::randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const
{
    const int i = m / Nz, k = m % Nz;
    const double R = Rv[i] + (Rv[i + 1] - Rv[i]) * rng.uniform();
    const double phi = 2.0 * M_PI * rng.uniform();
    z = zv[k] + (zv[k + 1] - zv[k]) * rng.uniform();
    x = R * std::cos(phi); y = R * std::sin(phi);
}
}   // anonymous namespace

// =====================================================================================================
// factories
// =====================================================================================================
Grid* makeCartesian(const double* xv, int Nx, const double* yv, int Ny, const double* zv, int Nz)
{
    Cartesian* g = new Cartesian();
    g->xv.assign(xv, xv + Nx + 1); g->yv.assign(yv, yv + Ny + 1); g->zv.assign(zv, zv + Nz + 1);
    g->Nx = Nx; g->Ny = Ny; g->Nz = Nz;
    // for every Mesh of the reference mesh[0]=0 and mesh[N]=1, so the outer borders are the extent (CartesianDustGrid.cpp:34-36)
    g->xmin = xv[0]; g->xmax = xv[Nx]; g->ymin = yv[0]; g->ymax = yv[Ny]; g->zmin = zv[0]; g->zmax = zv[Nz];
    return g;
}

Grid* makeTree(int kind, int search, int N, const double* box, const int* child0, const int* parent, const int* cell,
               const int* dir, const int* nbrStart, const int* nbrIds)
{
    Tree* g = new Tree();
    g->kind = kind; g->search = search; g->N = N;
    g->box.assign(box, box + 6 * (size_t)N); g->child0.assign(child0, child0 + N); g->parent.assign(parent, parent + N);
    g->cell.assign(cell, cell + N);
    if (dir) g->dir.assign(dir, dir + N); else g->dir.assign(N, 0);
    if (nbrStart) { g->nbrStart.assign(nbrStart, nbrStart + 6 * (size_t)N + 1); g->nbrIds.assign(nbrIds, nbrIds + nbrStart[6 * (size_t)N]); }
    for (int l = 0; l < N; l++) if (cell[l] >= 0) g->ncells++;
    g->leafOfCell.assign(g->ncells, -1);
    for (int l = 0; l < N; l++) if (cell[l] >= 0) g->leafOfCell[cell[l]] = l;
    g->eps = epsOf(box);
    return g;
}

Grid* makeAdaptiveMesh(int N, const double* box, const int* nxyz, const int* child0, const int* cell, const int* wallNbr)
{
    AMesh* g = new AMesh();
    g->N = N;
    g->box.assign(box, box + 6 * (size_t)N); g->nxyz.assign(nxyz, nxyz + 3 * (size_t)N); g->child0.assign(child0, child0 + N);
    g->cell.assign(cell, cell + N); g->wallNbr.assign(wallNbr, wallNbr + 6 * (size_t)N);
    for (int l = 0; l < N; l++) if (cell[l] >= 0) g->ncells++;
    g->leafOfCell.assign(g->ncells, -1);
    for (int l = 0; l < N; l++) if (cell[l] >= 0) g->leafOfCell[cell[l]] = l;
    g->eps = epsOf(box);
    return g;
}

Grid* makeVoronoi(int N, const double* particles, const int* nbrStart, const int* nbrIds, const double extent[6], int nb,
                  const int* blkStart, const int* blkIds, const int* blkTree, int Nkd, const int* kdM, const int* kdAxis,
                  const int* kdUp, const int* kdLeft, const int* kdRight, const double* cellBox)
{
    Voronoi* g = new Voronoi();
    g->N = N; g->nb = nb;
    size_t nb3 = (size_t)nb * nb * nb;
    g->part.assign(particles, particles + 3 * (size_t)N);
    g->nbrStart.assign(nbrStart, nbrStart + N + 1); g->nbrIds.assign(nbrIds, nbrIds + nbrStart[N]);
    g->blkStart.assign(blkStart, blkStart + nb3 + 1); g->blkIds.assign(blkIds, blkIds + blkStart[nb3]);
    g->blkTree.assign(blkTree, blkTree + nb3);
    if (Nkd > 0)
    {
        g->kdM.assign(kdM, kdM + Nkd); g->kdAxis.assign(kdAxis, kdAxis + Nkd); g->kdUp.assign(kdUp, kdUp + Nkd);
        g->kdLeft.assign(kdLeft, kdLeft + Nkd); g->kdRight.assign(kdRight, kdRight + Nkd);
        for (int& a : g->kdAxis) a %= 3;
    }
    if (cellBox) g->cellBox.assign(cellBox, cellBox + 6 * (size_t)N);
    // extent arrives as xmin,xmax,ymin,ymax,zmin,zmax (the order of the ski properties)
    g->ext[0] = extent[0]; g->ext[1] = extent[2]; g->ext[2] = extent[4]; g->ext[3] = extent[1]; g->ext[4] = extent[3]; g->ext[5] = extent[5];
    g->eps = epsOf(g->ext);
    return g;
}

Grid* makeSphere1D(int Nr, const double* rv) { Sphere1D* g = new Sphere1D(); g->Nr = Nr; g->rv.assign(rv, rv + Nr + 1); return g; }
Grid* makeSphere2D(int Nr, const double* rv, int Nt, const double* thetav, const double* cv)
{
    Sphere2D* g = new Sphere2D(); g->Nr = Nr; g->Nt = Nt;
    g->rv.assign(rv, rv + Nr + 1); g->thetav.assign(thetav, thetav + Nt + 1); g->cv.assign(cv, cv + Nt + 1);
    return g;
}
Grid* makeCylinder2D(int NR, const double* Rv, int Nz, const double* zv)
{ Cylinder2D* g = new Cylinder2D(); g->NR = NR; g->Nz = Nz; g->Rv.assign(Rv, Rv + NR + 1); g->zv.assign(zv, zv + Nz + 1); return g; }

}   // namespace orc

// =====================================================================================================
// C interface for ctypes (oracle/oracle_py.py)
// =====================================================================================================
using namespace orc;

extern "C"
{
void* orc_grid_cartesian(const double* xv, int Nx, const double* yv, int Ny, const double* zv, int Nz) { return makeCartesian(xv, Nx, yv, Ny, zv, Nz); }
void* orc_grid_tree(int kind, int search, int N, const double* box, const int* child0, const int* parent, const int* cell,
                    const int* dir, const int* nbrStart, const int* nbrIds)
{ return makeTree(kind, search, N, box, child0, parent, cell, dir, nbrStart, nbrIds); }
void* orc_grid_amesh(int N, const double* box, const int* nxyz, const int* child0, const int* cell, const int* wallNbr)
{ return makeAdaptiveMesh(N, box, nxyz, child0, cell, wallNbr); }
void* orc_grid_voronoi(int N, const double* particles, const int* nbrStart, const int* nbrIds, const double* extent, int nb,
                       const int* blkStart, const int* blkIds, const int* blkTree, int Nkd, const int* kdM, const int* kdAxis,
                       const int* kdUp, const int* kdLeft, const int* kdRight, const double* cellBox)
{ return makeVoronoi(N, particles, nbrStart, nbrIds, extent, nb, blkStart, blkIds, blkTree, Nkd, kdM, kdAxis, kdUp, kdLeft, kdRight, cellBox); }
void* orc_grid_sphere1d(int Nr, const double* rv) { return makeSphere1D(Nr, rv); }
void* orc_grid_sphere2d(int Nr, const double* rv, int Nt, const double* thetav, const double* cv) { return makeSphere2D(Nr, rv, Nt, thetav, cv); }
void* orc_grid_cylinder2d(int NR, const double* Rv, int Nz, const double* zv) { return makeCylinder2D(NR, Rv, Nz, zv); }
void orc_grid_destroy(void* g) { delete (Grid*)g; }
int orc_num_cells(void* g) { return ((Grid*)g)->numCells(); }
long orc_stuck(void* g) { return ((Grid*)g)->stuck; }

void* orc_medium(int Ncells, int Ncomp, int Nlambda, const double* rho, const double* kext, const double* ksca, const double* g)
{
    Medium* m = new Medium();
    m->Ncells = Ncells; m->Ncomp = Ncomp; m->Nlambda = Nlambda;
    m->rho.assign(rho, rho + (size_t)Ncells * Ncomp); m->kext.assign(kext, kext + (size_t)Ncomp * Nlambda);
    if (ksca) m->ksca.assign(ksca, ksca + (size_t)Ncomp * Nlambda); else m->ksca.assign((size_t)Ncomp * Nlambda, 0.0);
    if (g) m->g.assign(g, g + (size_t)Ncomp * Nlambda); else m->g.assign((size_t)Ncomp * Nlambda, 0.0);
    return m;
}
void orc_medium_destroy(void* m) { delete (Medium*)m; }

// batched path(): pass cap = 0 to obtain the counts/offsets, then again with arrays of offsets[n] entries.
// ell: NULL (geometry only), one value (ellStride 0) or one per ray (ellStride 1)
long orc_path_batch(void* gh, void* mh, const double* r, const double* k, long n, const int* ell, int ellStride, long cap,
                    long* offsets, int* m, double* ds, double* s, double* dtau, double* tau, int nthreads)
{
    const Grid* g = (const Grid*)gh; const Medium* med = (const Medium*)mh;
    std::vector<int> counts(n);
    auto work = [&](int pass, int tid, int nt)
    {
        Path p;
        for (long i = tid; i < n; i += nt)
        {
            p.rx = r[3 * i]; p.ry = r[3 * i + 1]; p.rz = r[3 * i + 2]; p.kx = k[3 * i]; p.ky = k[3 * i + 1]; p.kz = k[3 * i + 2];
            g->path(p);
            if (pass == 0) { counts[i] = (int)p.v.size(); continue; }
            if (ell && med) fillOpticalDepth(p, *med, ell[i * ellStride]);
            long o = offsets[i];
            if (o + (long)p.v.size() > cap) continue;
            for (size_t j = 0; j < p.v.size(); j++)
            { m[o + j] = p.v[j].m; ds[o + j] = p.v[j].ds; s[o + j] = p.v[j].s; dtau[o + j] = p.v[j].dtau; tau[o + j] = p.v[j].tau; }
        }
    };
    auto run = [&](int pass)
    {
        int nt = std::max(1, nthreads);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, pass, t, nt);
        work(pass, 0, nt);
        for (auto& t : th) t.join();
    };
    if (cap <= 0)
    {
        run(0);
        long total = 0;
        for (long i = 0; i < n; i++) { offsets[i] = total; total += counts[i]; }
        offsets[n] = total;
        return total;
    }
    run(1);
    return offsets[n];
}

void orc_whichcell(void* gh, const double* r, long n, int* m)
{ const Grid* g = (const Grid*)gh; for (long i = 0; i < n; i++) m[i] = g->whichcell(r[3 * i], r[3 * i + 1], r[3 * i + 2]); }

// DustSystem::opticaldepth(pp, distance), DustSystem.cpp:984-1000
void orc_opticaldepth(void* gh, void* mh, const double* r, const double* k, long n, const int* ell, int ellStride,
                      const double* distance, double* tau)
{
    const Grid* g = (const Grid*)gh; const Medium* med = (const Medium*)mh;
    Path p;
    for (long i = 0; i < n; i++)
    {
        p.rx = r[3 * i]; p.ry = r[3 * i + 1]; p.rz = r[3 * i + 2]; p.kx = k[3 * i]; p.ky = k[3 * i + 1]; p.kz = k[3 * i + 2];
        g->path(p);
        tau[i] = opticalDepth(p, *med, ell[i * ellStride], distance ? distance[i] : DBL_MAX);
    }
}

// DustGridPath::pathlength on the path of one ray (for the propagation check)
double orc_pathlength(void* gh, void* mh, const double* r, const double* k, int ell, double tau)
{
    const Grid* g = (const Grid*)gh; const Medium* med = (const Medium*)mh;
    Path p; p.rx = r[0]; p.ry = r[1]; p.rz = r[2]; p.kx = k[0]; p.ky = k[1]; p.kz = k[2];
    g->path(p); fillOpticalDepth(p, *med, ell);
    return p.pathlength(tau);
}

void orc_random_positions(void* gh, int m, unsigned long seed, long n, double* xyz)
{
    const Grid* g = (const Grid*)gh; Rng rng(seed);
    for (long i = 0; i < n; i++) g->randomPositionInCell(m, rng, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
void orc_uniforms(unsigned long seed, long n, double* u) { Rng rng(seed); for (long i = 0; i < n; i++) u[i] = rng.uniform(); }
}
