// Device-resident table views consumed by the traversal / Monte Carlo kernels.
// Layouts are those documented in include/skirtgpu.h (flattened state of the reference's grids).
#pragma once
#include <cstdint>

namespace skg
{

enum GridKind { GRID_NONE = -1, GRID_CART = 0, GRID_TREE = 1, GRID_AMESH = 2, GRID_VORO = 3, GRID_SYM = 4 };

struct CartGrid
{
    const double* xv; const double* yv; const double* zv;   // borders, Nx+1 / Ny+1 / Nz+1 values
    int Nx, Ny, Nz;
    double ext[6];      // xmin,xmax,ymin,ymax,zmin,zmax of the BoxDustGrid extent
    unsigned sx, sy, sz;        // shared-window byte addresses of the staged borders (kernel-side view only)
    int staged;                 // non-zero when xv/yv/zv have been staged in shared memory (stageCart)
    int uniform;                // every axis is a LinMesh: borders = min + i * width (to rounding); then wx, wy, wz are the bin widths
    double wx, wy, wz;
    const double* rhoAhead;     // kernel-side view only: density table whose row of the NEXT cell a crossing pulls into L1 (or null)
    int rhoAheadStride;         // doubles per cell in that table
};

// doubles of shared memory taken by the staged borders (three arrays with one pad element on either side)
#define SKG_CART_SMEM_DOUBLES(c) ((size_t)((c).Nx + (c).Ny + (c).Nz + 9))

// everything the walker needs to continue from a tree node, in one 96-byte record per NODE (three 256-bit reads): its box,
// cell number, the first neighbour of each wall (the common case is a single neighbour per wall) and where the rest of its
// neighbour lists start.  A node table stays resident in L2 (96 B x N; the lists expanded per neighbour did not), and a
// crossing into a single-neighbour wall costs one dependent L2 round trip.
struct __align__(32) TreeNodeRec
{
    double box[6];
    int cell;
    int hbase;                  // nbrHint offset / 4 of the node's first wall-bin block
    unsigned hmeta; int pad0;   // per wall w, bits 5w..5w+4: bit 0 = the wall has several neighbours, bits 1-2 = lg with G = 2 << lg,
                                // bit 3 = the first neighbour covers at least half of the wall, bit 4 = the wall's four neighbours
                                // are siblings whose ids are first[w] + sa*ia + sb*ib for the 2 x 2 wall bin (ia, ib)
    int first[6];               // first neighbour of each wall (bit 4: the neighbour of bin (0, 0)), -1 when the list is empty
    int pad1[2];
};

struct TreeGrid
{
    const double* box;              // [6N] xmin,ymin,zmin,xmax,ymax,zmax
    const int* child0; const int* parent; const int* cell; const int* dir;
    const int* nbrStart; const int* nbrIds;
    const int* cellNode;            // leaf node of every cell (TreeDustGrid::getnode, for randomPositionInCell)
    const TreeNodeRec* nodeRec;     // per-node records for the Neighbor search, or null
    const int* nbrHint;             // per multi-neighbour wall, G x G ids: the neighbour covering the centre of each wall bin (a hint: the walker verifies it);
                                    // the blocks of a node follow each other in wall order from 4 hbase
    const int* lookup;              // [G^3] deepest node whose box contains the whole lookup cell (entry point of root descents)
    int lookupG; double lookupInv[3];
    int N, kind, search;
    double eps;
};

// one adaptive mesh node in 96 bytes (three 256-bit reads): what a crossing needs from the neighbour beyond a wall (box,
// cell, its own wall neighbours) and what a descent needs from an internal node (box, child grid, first child)
struct __align__(32) AMeshNodeRec
{
    double box[6];
    int cell, child0;
    int nx, ny, nz; int parent;     // (the root is its own parent's 0)
    int wallNbr[6];
};

struct AMeshGrid
{
    const double* box; const int* nxyz; const int* child0; const int* cell; const int* wallNbr;
    const AMeshNodeRec* nodeRec;
    const int* cellNode;            // leaf node of every cell
    int N;
    double eps;
};

// The grids with symmetries: Sphere1DDustGrid (sub 0: radial borders rv), Sphere2DDustGrid (sub 1: rv and polar borders thetav with
// their cosines cv, the xy-plane among them) and Cylinder2DDustGrid (sub 2: radial borders Rv in v1, vertical borders zv in v2)
struct SymGrid
{
    int sub; int N1, N2;            // bins along the first (r / R) and second (theta / z; 0 for sub 0) coordinate
    const double* v1; const double* v2; const double* cv;
    double rmax, zmin, zmax;
};

struct VoroGrid
{
    const double* particles;        // [3N]
    const int* nbrStart; const int* nbrIds;
    const int* blkStart; const int* blkIds; const int* blkTree;
    const int* kdM; const int* kdAxis; const int* kdUp; const int* kdLeft; const int* kdRight;
    const double* cellBox;          // [6N] xmin,ymin,zmin,xmax,ymax,zmax
    // crossing records, 32-byte slots: cell m owns slots [nbrStart[m] + m, nbrStart[m+1] + m + 1): a header {particle of m,
    // (neighbour count, 0)} followed by one slot per neighbour in list order {particle of the neighbour, (id, first slot of
    // the neighbour's own block)} -- a crossing reads one contiguous block instead of ids -> particle positions
    const double* rec;
    // the same for the shooting stages (whose results are Monte Carlo estimates), same slots: the header {particle p of m,
    // (neighbour count, 0)}, then per neighbour {n = p_i - p, (id, first slot of the neighbour's block)} -- the bisector plane
    // is n.(r - p) = |n|^2 / 2, the distance along a ray (|n|^2 / 2 - n.(r - p)) / n.k.  A wall of the domain at distance D
    // from p is the bisector towards p's mirror image: n = 2 D along the outward normal, so that one formula serves all
    const double* planes;
    double ext[6];                  // xmin,ymin,zmin,xmax,ymax,zmax
    double eps;
    int N, nb;
};

struct Medium
{
    const double* rho;              // [Ncells*Ncomp]
    const double* kext; const double* ksca; const double* g;    // [Ncomp*Nlambda]
    int Ncells, Ncomp, Nlambda;
    // polarisation (DustMix::addpolarization, DustMix.cpp:325-361): Mueller matrix coefficients [Ncomp*Nlambda*Ntheta] on
    // theta_t = t*pi/(Ntheta-1), the cumulative distribution of theta per wavelength and the phase function normalisation
    // (DustMix.cpp:96-123); Ntheta == 0: no polarisation (Henyey-Greenstein scattering)
    const double* S11; const double* S12; const double* S33; const double* S34;
    const double* thetaX;           // [Ncomp*Nlambda*Ntheta]
    const double* pfnorm;           // [Ncomp*Nlambda]
    int Ntheta;
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
    unsigned long long peelSegments;        // packet-steps of the peel-off stage
    unsigned long long propSegments;        // packet-steps of the propagation stage
    unsigned long long pad;
};

}   // namespace skg
