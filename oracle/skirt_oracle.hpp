// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("port") of the SKIRT v7.3 photon-packet propagation hot path, working on the
// same flattened POD tables that the CUDA engine consumes (see include/skirtgpu.h).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; the product path (skirt_b200/) never links, imports or calls it.
//
// Parity pinning: every routine below is checked in this repository's CPU test-suite against
//   (1) oracle/_ref/libskirtref.so -- the reference's OWN translation units compiled in place
//       from /root/reference (recipe: oracle/Makefile), when that library is present, and
//   (2) the golden vectors under tests/golden/ which were generated from that library by
//       oracle/make_golden.py (committed with the vectors).
// The reference ships no tests / golden vectors of its own.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off (no -march=native): the reference is built with
// plain -O3 on x86-64 (SKIRTcore/SKIRTcore.pro), i.e. without FMA contraction.
#ifndef SKIRT_ORACLE_HPP
#define SKIRT_ORACLE_HPP

#include <cstdint>
#include <vector>

namespace orc
{

// ---------------------------------------------------------------------------------------------
// path record; follows DustGridPath (SKIRTcore/DustGridPath.hpp:157-167, DustGridPath.cpp:38-53)
// ---------------------------------------------------------------------------------------------
struct Seg { int m; double ds, s, dtau, tau; };

struct Path
{
    double rx = 0, ry = 0, rz = 0;   // start position
    double kx = 0, ky = 0, kz = 1;   // direction
    double s = 0;                    // running path length
    std::vector<Seg> v;

    void clear() { s = 0; v.clear(); }
    void add(int m, double ds)      // DustGridPath::addSegment, DustGridPath.cpp:46-53
    {
        if (ds > 0) { s += ds; v.push_back(Seg{m, ds, s, 0., 0.}); }
    }
    // DustGridPath::moveInside, DustGridPath.cpp:57-150; returns false for the OUTSIDE position
    bool moveInside(const double box[6], double eps, double& x, double& y, double& z);
    double tau() const { return v.empty() ? 0. : v.back().tau; }
    double pathlength(double tau) const;   // DustGridPath.cpp:162-173
};

struct Rng;   // MT19937 stream, Random.cpp:41-126

// ---------------------------------------------------------------------------------------------
// grids
// ---------------------------------------------------------------------------------------------
struct Grid
{
    virtual ~Grid() {}
    virtual int numCells() const = 0;
    virtual void path(Path& p) const = 0;
    virtual int whichcell(double x, double y, double z) const = 0;
    virtual void randomPositionInCell(int m, Rng& rng, double& x, double& y, double& z) const = 0;
    mutable long stuck = 0;   // number of "stuck packet" warnings (host side effect in the reference)
};

Grid* makeCartesian(const double* xv, int Nx, const double* yv, int Ny, const double* zv, int Nz);

// kind: 0 octree, 1 binary (k-d) tree;  search: 0 TopDown, 1 Neighbor, 2 Bookkeeping
Grid* makeTree(int kind, int search, int Nnodes, const double* box, const int* firstChild,
               const int* parent, const int* cell, const int* splitDir,
               const int* nbrStart, const int* nbrIds);

Grid* makeAdaptiveMesh(int Nnodes, const double* box, const int* nxyz, const int* firstChild,
                       const int* cell, const int* wallNbr);

Grid* makeVoronoi(int Ncells, const double* particles, const int* nbrStart, const int* nbrIds,
                  const double extent[6], int nb, const int* blkStart, const int* blkIds,
                  const int* blkTree, int Nkd, const int* kdM, const int* kdAxis, const int* kdUp,
                  const int* kdLeft, const int* kdRight, const double* cellBox);

// grids with symmetries (Sphere1DDustGrid.cpp, Sphere2DDustGrid.cpp, Cylinder2DDustGrid.cpp): the border arrays are the whole state
Grid* makeSphere1D(int Nr, const double* rv);
Grid* makeSphere2D(int Nr, const double* rv, int Ntheta, const double* thetav, const double* cv);
Grid* makeCylinder2D(int NR, const double* Rv, int Nz, const double* zv);

// ---------------------------------------------------------------------------------------------
// medium (DustSystem density table + DustMix per-wavelength scalars)
// ---------------------------------------------------------------------------------------------
struct Medium
{
    int Ncells = 0, Ncomp = 0, Nlambda = 0;
    std::vector<double> rho;    // [Ncells*Ncomp]  DustSystem.hpp:434 (_rhovv(m,h))
    std::vector<double> kext;   // [Ncomp*Nlambda]
    std::vector<double> ksca;   // [Ncomp*Nlambda]
    std::vector<double> g;      // [Ncomp*Nlambda]
    // KappaRho functor, DustSystem.cpp:465-491 (+ density(m,h) :918-921)
    double kapparho(int m, int ell) const
    {
        double result = 0;
        for (int h = 0; h < Ncomp; h++)
            result += kext[h*Nlambda + ell] * (m >= 0 ? rho[(size_t)m*Ncomp + h] : 0.);
        return result;
    }
};

void fillOpticalDepth(Path& p, const Medium& med, int ell);                 // DustGridPath.hpp:117-129
double opticalDepth(const Path& p, const Medium& med, int ell, double distance);  // :97-108

}   // namespace orc

#endif
