// Device-resident table views consumed by the traversal / Monte Carlo kernels.
// Layouts are those documented in include/skirtgpu.h (flattened state of the reference's grids).
#pragma once
#include <cstdint>

namespace skg
{

enum GridKind { GRID_NONE = -1, GRID_CART = 0, GRID_TREE = 1, GRID_AMESH = 2, GRID_VORO = 3 };

struct CartGrid
{
    const double* xv; const double* yv; const double* zv;   // borders, Nx+1 / Ny+1 / Nz+1 values
    int Nx, Ny, Nz;
    double ext[6];      // xmin,xmax,ymin,ymax,zmin,zmax of the BoxDustGrid extent
    unsigned sx, sy, sz;        // shared-window byte addresses of the staged borders (kernel-side view only)
    int staged;                 // non-zero when xv/yv/zv have been staged in shared memory (stageCart)
};

// doubles of shared memory taken by the staged borders (three arrays with one pad element on either side)
#define SKG_CART_SMEM_DOUBLES(c) ((size_t)((c).Nx + (c).Ny + (c).Nz + 9))

// one entry of a tree node's neighbour list, self-contained: everything the walker needs to continue from that
// neighbour (its box, cell number and the offsets of ITS six neighbour lists), so that a crossing costs a single
// dependent memory round trip instead of id -> box -> offsets
struct __align__(32) TreeNbrRec
{
    double box[6];
    int id, cell;
    int nb[7];
    int pad[3];
};

struct TreeGrid
{
    const double* box;              // [6N] xmin,ymin,zmin,xmax,ymax,zmax
    const int* child0; const int* parent; const int* cell; const int* dir;
    const int* nbrStart; const int* nbrIds;
    const int* cellNode;            // leaf node of every cell (TreeDustGrid::getnode, for randomPositionInCell)
    const TreeNbrRec* nbrRec;       // expanded neighbour lists (same order as nbrIds), or null
    const int* lookup;              // [G^3] deepest node whose box contains the whole lookup cell (entry point of root descents)
    int lookupG; double lookupInv[3];
    int N, kind, search;
    double eps;
};

struct AMeshGrid
{
    const double* box; const int* nxyz; const int* child0; const int* cell; const int* wallNbr;
    const int* cellNode;            // leaf node of every cell
    int N;
    double eps;
};

struct VoroGrid
{
    const double* particles;        // [3N]
    const int* nbrStart; const int* nbrIds;
    const int* blkStart; const int* blkIds; const int* blkTree;
    const int* kdM; const int* kdAxis; const int* kdUp; const int* kdLeft; const int* kdRight;
    const double* cellBox;          // [6N] xmin,ymin,zmin,xmax,ymax,zmax
    double ext[6];                  // xmin,ymin,zmin,xmax,ymax,zmax
    double eps;
    int N, nb;
};

struct Medium
{
    const double* rho;              // [Ncells*Ncomp]
    const double* kext; const double* ksca; const double* g;    // [Ncomp*Nlambda]
    int Ncells, Ncomp, Nlambda;
};

// counters updated by the kernels (device memory, one instance per engine)
struct Counters
{
    unsigned long long stuckEscaped;        // "seems stuck -- escaping" (TreeDustGrid.cpp:437-446)
    unsigned long long stuckTerminated;     // "is stuck -- terminating this path" (:449-454)
    unsigned long long errors;              // conditions on which the reference throws FATALERROR
    unsigned long long segments;            // packet-steps (addSegment with ds>0)
    unsigned long long paths;               // traversals
    unsigned long long scatterings;
    unsigned long long packets;
    unsigned long long absorbSegments;      // segments that updated the absorption table (one fp64 atomic each)
    unsigned long long detections;          // detector updates (one fp64 atomic each)
    unsigned long long pad;
};

}   // namespace skg
