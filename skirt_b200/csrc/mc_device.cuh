// Device-side pieces of the photon-packet life cycle shared by the stage kernels (mc_kernels.cu):
// packet pool layout, launch samplers, life-cycle sinks and warp-aggregated accumulation.
//   launch                StellarSystem::launch (StellarSystem.cpp:116-158) + geometry samplers
//   escape + absorption   MonteCarloSimulation.cpp:438-515 (DustSystem::absorb -> atomicAdd, replaces LockFree::add)
//   forced propagation    :519-537 (+ DustGridPath::pathlength, DustGridPath.cpp:162-173)
//   scattering            :541-549 (DustMix.cpp:607-614, Random::direction Random.cpp:188-222)
#pragma once
#include <cmath>
#include "engine.h"
#include "geom.cuh"
#include "sinks.cuh"
#include "philox.cuh"

namespace skg
{

struct GridSetMC { CartGrid cart; TreeGrid tree; AMeshGrid amesh; VoroGrid voro; SymGrid sym; };

// Packet pool: one 96-byte record per in-flight photon packet (PhotonPackage, PhotonPackage.hpp:28,134-138;
// unpolarised: position, direction, luminosity, wavelength index, number of scatterings) plus the engine's own
// bookkeeping (Philox stream position, interaction optical depth sampled by the absorb stage).  The packets in flight
// are kept COMPACT: the absorb stage writes its survivors to consecutive records of a second pool and new packets are
// appended behind them, so every stage streams through records [0, n) -- consecutive lanes read consecutive 96-byte
// records (three 32-byte sectors each, moved as a whole), no index indirection and no random DRAM access.
struct __align__(32) Packet
{
    double x, y, z, kx, ky, kz;
    double L, target;
    unsigned long long id;      // Philox stream id of the packet
    int ell, nscatt; unsigned rngCtr; int fresh;
    int hint;                   // the walker's locator of the packet's position (tree / adaptive-mesh leaf node, Voronoi cell) when a
                                // traversal has established it, else -1: the next traversal from this position skips the point location
    int comp;                   // PhotonPackage::stellarCompIndex(): the stellar component that emitted the packet, -1 for dust emission
};
typedef Packet* PacketPool;

// whole-record access: three 256-bit transactions (LDG/STG.E.ENL2.256 on sm_100a)
__device__ __forceinline__ Packet loadPacket(const Packet* p)
{
#ifdef SKG_PLAIN_PACKET_LOAD
    return *p;
#endif
    Packet r; double* w = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < 3; i++)
        asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(w[4 * i]), "=d"(w[4 * i + 1]), "=d"(w[4 * i + 2]), "=d"(w[4 * i + 3]) : "l"(reinterpret_cast<const double*>(p) + 4 * i));
    return r;
}
__device__ __forceinline__ void storePacket(Packet* p, const Packet& r)
{
    const double* w = reinterpret_cast<const double*>(&r);
#pragma unroll
    for (int i = 0; i < 3; i++)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(reinterpret_cast<double*>(p) + 4 * i), "d"(w[4 * i]), "d"(w[4 * i + 1]), "d"(w[4 * i + 2]), "d"(w[4 * i + 3]) : "memory");
}

// Stokes state of a packet in a simulation with polarisation (StokesVector.hpp:88-90: Q, U, V relative to I, the normal
// to the last scattering plane, and whether a scattering plane exists yet); one 64-byte record per pool slot, in a pool
// of its own that exists only when the medium is polarised (skg_medium_polarization)
struct __align__(32) PolState
{
    double Q, U, V, nx, ny, nz;
    int polarized, pad; double pad2;
};

// instruments that look along the same direction share one peel-off traversal
struct ObsGroup { double kx, ky, kz; int first, count; };
// PerspectiveInstrument: eye position, world -> pixel transform (HomogeneousTransform::M, row vector convention), pixel size, data cube
struct PerspDev { double Ex, Ey, Ez; double M[4][4]; int Nx, Ny; double s; double* frame; };

struct McDev
{
    Medium med;
    const SourceDev* sources; int Nsources;
    const double* L;        // [Nsources*Nlambda]
    const double* Ltot;     // [Nlambda]
    const double* Lcdf;     // [Nlambda*(Nsources+1)]
    double emissionBias;
    const InstrDev* instr; int Ninstr;      // sorted by observer group
    const PerspDev* persp; int Npersp;      // PerspectiveInstruments: one peel-off ray per packet each, towards the eye
    const ObsGroup* groups; int Ngroups; int maxGroupCount;     // largest number of instruments sharing one line of sight
    double* labs;           // [Nlambda*Ncells] (wavelength-major on the device) or null
    double Lscale;          // total packets per wavelength over all engines
    double minWeightReduction, minfs, xi;
    int contScatt;          // MonteCarloSimulation::continuousScattering: peel-off from every path segment instead of the interaction points
    uint64_t seed, streamOffset;
    const int* ellList;     // wavelength indices with nonzero luminosity, in shooting order
    unsigned long long NppInt;
    PacketPool pool;        // packets of this iteration, compact: [0, nAlive)
    PacketPool poolNext;    // survivors of the absorb stage are written here, compact again
    PolState* pol; PolState* polNext;       // the packets' Stokes states, same slots (null: no polarisation)
    // dust emission phases (PanMonteCarloSimulation.cpp:187-342)
    int phase;              // SKG_PHASE_*
    unsigned rngKind;       // Philox stream kind, so that the phases of one simulation never share deviates
    const double* dustLv;   // [Nlambda*Ncells] luminosity per cell, wavelength-major
    const double* dustCdf;  // [Nlambda*(Ncells+1)] normalised cumulative distributions
    double dustBias;        // PanDustSystem::emissionBias
    int refill;             // lanes of a warp that must be parked before they finish and draw new work together
    int propRefill;         // the same for the propagate stage
    int peelRefill;         // the same for the peel-off stage (one item per packet and observer direction: cheap to begin)
};

// expm1(x) for x <= 0: the per-segment absorbed fraction -expm1(-dtau) (MonteCarloSimulation.cpp:452).  Two branch-free
// forms, chosen per WARP (the library function is a long, divergent slow path that segments with a large optical depth
// would drag the whole warp through): when every lane's argument is small, the Taylor polynomial up to x^8/8! (remainder
// < 3e-18 relative for |x| <= 2^-5, 8 FMAs); otherwise x = k ln2 + r with |r| <= ln2/2, expm1(r) by its Taylor
// polynomial up to r^13/13! (remainder < 5e-16 relative) and expm1(x) = 2^k expm1(r) + (2^k - 1) -- exact cancellation-
// free for k = 0, and correct to ~1 ulp of the result elsewhere.
__device__ __forceinline__ double expm1Poly8(double x)
{
    double p = 1.0 / 40320.0;
    p = __fma_rn(p, x, 1.0 / 5040.0);
    p = __fma_rn(p, x, 1.0 / 720.0);
    p = __fma_rn(p, x, 1.0 / 120.0);
    p = __fma_rn(p, x, 1.0 / 24.0);
    p = __fma_rn(p, x, 1.0 / 6.0);
    p = __fma_rn(p, x, 0.5);
    p = __fma_rn(p, x, 1.0);
    return p * x;
}
__device__ __forceinline__ double expm1Reduced(double x)
{
    x = fmax(x, -700.0);                                            // exp(-700) ~ 1e-304: expm1 = -1 to the last bit
    const double kf = rint(x * 1.4426950408889634074);
    double r = __fma_rn(kf, -6.93147180369123816490e-01, x);        // Cody-Waite: ln2 = hi + lo
    r = __fma_rn(kf, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = __fma_rn(p, r, 1.0 / 479001600.0);
    p = __fma_rn(p, r, 1.0 / 39916800.0);
    p = __fma_rn(p, r, 1.0 / 3628800.0);
    p = __fma_rn(p, r, 1.0 / 362880.0);
    p = __fma_rn(p, r, 1.0 / 40320.0);
    p = __fma_rn(p, r, 1.0 / 5040.0);
    p = __fma_rn(p, r, 1.0 / 720.0);
    p = __fma_rn(p, r, 1.0 / 120.0);
    p = __fma_rn(p, r, 1.0 / 24.0);
    p = __fma_rn(p, r, 1.0 / 6.0);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    const double e = p * r;                                         // expm1(r)
    const double s = __longlong_as_double((long long)((int)kf + 1023) << 52);      // 2^k, k in [-1010, 0]
    return __fma_rn(s, e, s - 1.0);
}
__device__ __forceinline__ double expm1Small(double x)
{
    if (__any_sync(__activemask(), x < -0.03125)) return expm1Reduced(x);
    return expm1Poly8(x);
}

// One-component media: the density gather of the shooting stages.  A random 8-byte gather costs the SM's load/store unit
// one wavefront per lane whatever its width, and that -- not arithmetic -- is what bounds the stage kernels.  So a lane
// reads the whole aligned 32-byte SECTOR that holds its cell (four consecutive cell numbers: on a Cartesian grid four
// neighbours along z, m = k + Nz*j + Nz*Ny*i) with one 256-bit load and keeps it: the next crossings along z, the most
// frequent ones in flattened grids, find their density in registers.  touch() starts the load of the sector of cell m
// unless it is the one held; get() picks a cell of the held sector.  Jobs consume a segment one crossing after it was
// parked (same summation order as without the cache), so the load has a crossing's worth of time to land.
// Measured on B200 (profiles/r02_f_sweep.txt): SLOWER than the plain 8-byte gather -- peel 68 -> 85 ms, absorb 130 -> 146 ms
// per C2 phase (a random 256-bit load occupies the load/store data path four times as long, and the eight extra
// registers cost occupancy) -- so it is compiled out; kept as a switch for the record.
#ifndef SKG_RHO_SECTOR
#define SKG_RHO_SECTOR 0
#endif
#if !SKG_RHO_SECTOR
struct RhoSector        // experiment switch: plain 8-byte gather per crossing
{
    double c0; int group;
    __device__ __forceinline__ void reset() { group = -1; c0 = 0.0; }
#ifdef SKG_EXP_NOGATHER
    __device__ __forceinline__ void touch(const double* rho, int m) { group = m; c0 = 1e-24 * (m & 7); }      // experiment: no density gather
#else
    __device__ __forceinline__ void touch(const double* rho, int m) { group = m; c0 = __ldg(rho + m); }
#endif
    __device__ __forceinline__ double get(int) const { return c0; }
};
#else
struct RhoSector
{
    double c0, c1, c2, c3; int group;
    __device__ __forceinline__ void reset() { group = -1; c0 = c1 = c2 = c3 = 0.0; }
    __device__ __forceinline__ void touch(const double* rho, int m)
    {
        const int g = m >> 2;
        if (g != group)
        {
            group = g;
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(rho + 4 * (size_t)g));
        }
    }
    __device__ __forceinline__ double get(int m) const { const int j = m & 3; return j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3)); }
};
#endif

// ---- accumulation ----------------------------------------------------------------------------------------
// LockFree::add (LockFree.hpp:25-37) on the device: lanes of a warp that target the same address are summed
// first and issue ONE fp64 atomicAdd (SED bins are a single address per wavelength, edge-on frames concentrate
// flux in few pixels).
__device__ __forceinline__ void warpAggregatedAdd(double* addr, double v)
{
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, (unsigned long long)addr);
    const int lane = threadIdx.x & 31;
    if (peers == 0xffffffffu)
    {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(addr, v);
        return;
    }
    const int leader = __ffs(peers) - 1;
    double sum = 0;
    for (unsigned rem = peers; rem; rem &= rem - 1) sum += __shfl_sync(peers, v, __ffs(rem) - 1);
    if (lane == leader) atomicAdd(addr, sum);
}

// The same for a value that (almost) all lanes of a CONVERGED warp add to one address -- SED bins: one address per
// wavelength, and the pool holds few adjacent wavelengths at a time.  Called by all 32 lanes; lanes without a
// contribution pass take = false.  When the contributing lanes agree on the address, the values are summed by a
// butterfly (5 shuffle steps) and lane 0 issues the one atomic; otherwise every group of equal addresses is served as in
// warpAggregatedAdd.
__device__ __forceinline__ void warpConvergedAdd(bool take, double* addr, double v)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned takers = __ballot_sync(FULL, take);
    if (!takers) return;
    const int first = __ffs(takers) - 1;
    double* ref = reinterpret_cast<double*>(__shfl_sync(FULL, (unsigned long long)addr, first));
    if (__all_sync(FULL, !take || addr == ref))
    {
        double sum = take ? v : 0.0;
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(ref, sum);
        return;
    }
    if (take) warpAggregatedAdd(addr, v);
}

// stream compaction: the lanes with `take` set obtain consecutive positions in an output array, one atomicAdd per warp
__device__ __forceinline__ int warpAppendPosition(bool take, int* count)
{
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, take);
    if (!mask) return -1;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(mask) - 1) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(active, base, __ffs(mask) - 1);
    return take ? base + __popc(mask & ((1u << lane) - 1)) : -1;
}

// ---- polarisation: StokesVector (StokesVector.cpp) and the Mueller-matrix parts of DustMix (DustMix.cpp:540-731) ------------
__device__ __forceinline__ void stokesUnpolarized(PolState& s) { s.Q = s.U = s.V = 0; s.nx = s.ny = s.nz = 0; s.polarized = 0; s.pad = 0; s.pad2 = 0; }
__device__ __forceinline__ double stokesLinearDegree(const PolState& s) { return sqrt(s.Q * s.Q + s.U * s.U); }                      // StokesVector.cpp:38-41
__device__ __forceinline__ double stokesAngle(const PolState& s) { return (s.U == 0 && s.Q == 0) ? 0.0 : 0.5 * atan2(s.U, s.Q); }     // :45-51

// StokesVector::rotateStokes, StokesVector.cpp:55-92
__device__ __forceinline__ void stokesRotate(PolState& s, double phi, double kx, double ky, double kz)
{
    if (!s.polarized)
    {
        // first scattering: generate the normal to the scattering plane (the Bianchi formula with phi = 0, theta = 90 deg)
        if (fabs(kz) > 0.99999) { s.nx = 1; s.ny = 0; s.nz = 0; }
        else { const double nz = sqrt((1.0 - kz) * (1.0 + kz)); s.nx = -kx * kz / nz; s.ny = -ky * kz / nz; s.nz = nz; }
        s.polarized = 1;
    }
    else
    {
        const double c2 = cos(2.0 * phi), s2 = sin(2.0 * phi);
        const double Q = c2 * s.Q + s2 * s.U, U = -s2 * s.Q + c2 * s.U;
        s.Q = Q; s.U = U;
    }
    // Rodrigues' rotation of the stored scattering plane about k, then renormalised
    const double c = cos(phi), sn = sin(phi);
    const double cx = ky * s.nz - kz * s.ny, cy = kz * s.nx - kx * s.nz, cz = kx * s.ny - ky * s.nx;       // Vec::cross(k, normal)
    double nx = s.nx * c + cx * sn, ny = s.ny * c + cy * sn, nz = s.nz * c + cz * sn;
    const double norm = sqrt(nx * nx + ny * ny + nz * nz);                                                 // Direction(Vec) stores as is; /= norm()
    s.nx = nx / norm; s.ny = ny / norm; s.nz = nz / norm;
}

// StokesVector::applyMueller (:96-105) followed by setPolarized (:13-27)
__device__ __forceinline__ void stokesMueller(PolState& s, double S11, double S12, double S33, double S34)
{
    const double I = S11 * 1. + S12 * s.Q, Q = S12 * 1. + S11 * s.Q, U = S33 * s.U + S34 * s.V, V = -S34 * s.U + S33 * s.V;
    if (I != 0.0) { s.Q = Q / I; s.U = U / I; s.V = V / I; s.polarized = 1; }
    else stokesUnpolarized(s);
}

// indexForTheta, DustMix.cpp:543-551
__device__ __forceinline__ int indexForTheta(double theta, int Ntheta)
{
    const double dt = M_PI / (Ntheta - 1);
    int t = static_cast<int>(theta / dt + 0.5);
    return t < 0 ? 0 : (t >= Ntheta ? Ntheta - 1 : t);
}

// angleBetweenScatteringPlanes(np, kc, kn), DustMix.cpp:557-567
__device__ __forceinline__ double anglePlanes(double npx, double npy, double npz, double kcx, double kcy, double kcz, double knx, double kny, double knz)
{
    double ncx = kcy * knz - kcz * kny, ncy = kcz * knx - kcx * knz, ncz = kcx * kny - kcy * knx;          // cross(kc, kn)
    const double norm = sqrt(ncx * ncx + ncy * ncy + ncz * ncz);
    ncx /= norm; ncy /= norm; ncz /= norm;
    const double cosphi = npx * ncx + npy * ncy + npz * ncz;
    const double ax = npy * ncz - npz * ncy, ay = npz * ncx - npx * ncz, az = npx * ncy - npy * ncx;       // cross(np, nc)
    const double sinphi = ax * kcx + ay * kcy + az * kcz;
    const double phi = atan2(sinphi, cosphi);
    return isfinite(phi) ? phi : 0.0;
}

// angleBetweenScatteringAndInstrumentReference(n, knew, ky), DustMix.cpp:573-579
__device__ __forceinline__ double angleInstrument(double nx, double ny, double nz, double kx, double ky, double kz, double yx, double yy, double yz)
{
    const double cosalpha = nx * yx + ny * yy + nz * yz;
    const double ax = ny * yz - nz * yy, ay = nz * yx - nx * yz, az = nx * yy - ny * yx;                   // cross(n, ky)
    return atan2(ax * kx + ay * ky + az * kz, cosalpha);
}

// Random::cdf(xv, Xv) (Random.cpp:131-136) for the tabulated theta distribution of (component, wavelength): DustMix::sampleTheta
__device__ __forceinline__ double sampleTheta(const Medium& med, int h, int ell, double X)
{
    const int Nt = med.Ntheta;
    const double* Xv = med.thetaX + ((size_t)h * med.Nlambda + ell) * Nt;
    const int i = locateClip(Xv, X, Nt);
    const double dt = M_PI / (Nt - 1);
    const double x1 = Xv[i], x2 = Xv[i + 1];
    return i * dt + ((X - x1) / (x2 - x1)) * ((i + 1) * dt - i * dt);              // NR::interpolate_linlin
}

// DustMix::samplePhi (DustMix.cpp:723-731): the azimuth distribution 1 + 2 PF' cos 2(phi - polAngle) as the cumulative
// table phiX[f] = phi/2pi + PF cos(2 polAngle) sin(2 phi) + PF sin(2 polAngle) (1 - cos 2 phi) on Nphi = 361 points,
// evaluated where the bisection of NR::locate_clip asks for it instead of being stored
__device__ __forceinline__ double samplePhi(const Medium& med, int h, int ell, double theta, double polDegree, double polAngle, double X)
{
    const int Nphi = 361;
    const int t = indexForTheta(theta, med.Ntheta);
    const size_t o = ((size_t)h * med.Nlambda + ell) * med.Ntheta + t;
    const double PF = polDegree * med.S12[o] / med.S11[o] / (4 * M_PI);
    const double A = cos(2 * polAngle) * PF, B = sin(2 * polAngle) * PF;
    const double df = 2 * M_PI / (Nphi - 1);
    auto phiX = [&](int f) { const double phi = f * df; return phi / (2 * M_PI) + A * sin(2 * phi) + B * (1 - cos(2 * phi)); };
    int i;
    if (X < phiX(0)) i = 0;                                                         // NR::locate_clip, NR.hpp:146-151
    else
    {
        int jl = -1, ju = Nphi - 1;                                                 // locate_basic over the first n-1 entries
        while (ju - jl > 1) { const int jm = (ju + jl) >> 1; if (X < phiX(jm)) ju = jm; else jl = jm; }
        i = jl;
    }
    const double x1 = phiX(i), x2 = phiX(i + 1);
    return i * df + ((X - x1) / (x2 - x1)) * ((i + 1) * df - i * df);
}

// ---- samplers ------------------------------------------------------------------------------------------

// SpecialFunctions::LambertW1, SpecialFunctions.cpp:579-627 (branch W_-1 for -1/e <= z < 0)
static __device__ double lambertW1(double z)
{
    const double eps = 1.0e-12;
    const double em1 = 0.3678794411714423215955237701614608;
    if (z == 0.0) return -SKG_DBL_MAX;
    double q = z + em1;
    if (q < 0) q = 0;
    double r = -sqrt(q);
    double t8 = -8.401032217523977370984161688514 + r * (12.250753501314460424 + r * (-18.100697012472442755 + r * 27.029044799010561650));
    double t5 = 3.066858901050631912893148922704 + r * (-4.175335600258177138854984177460 + r * (5.858023729874774148815053846119 + r * t8));
    double t1 = 2.331643981597124203363536062168 + r * (-1.812187885639363490240191647568 + r * (1.936631114492359755363277457668
              + r * (-2.353551201881614516821543561516 + r * t5)));
    double w0 = -1.0 + r * t1;
    if (q < 3.0e-3) return w0;
    double w;
    if (z < -1e-6) w = w0;
    else { double l1 = log(-z); double l2 = log(-l1); w = l1 - l2 + l2 / l1; }
    for (int i = 0; i < 10; i++)
    {
        double e = exp(w);
        double t = w * e - z;
        double p = w + 1.0;
        t /= e * p - 0.5 * (p + 1.0) * t / p;
        w -= t;
        if (fabs(t) < eps * (1.0 + fabs(w))) return w;
    }
    return w;
}

// Direction(theta, phi), Direction.cpp:12-38
__device__ __forceinline__ void directionFromAngles(double theta, double phi, double& kx, double& ky, double& kz)
{
    const double eps = 1e-8;
    if (theta <= eps) { kx = 0; ky = 0; kz = 1; }
    else if (theta >= M_PI - eps) { kx = 0; ky = 0; kz = -1; }
    else { double st = sin(theta); kx = st * cos(phi); ky = st * sin(phi); kz = cos(theta); }
}

// Random::direction(), Random.cpp:179-184
__device__ __forceinline__ void randomDirection(Philox& rng, double& kx, double& ky, double& kz)
{
    double theta = acos(2.0 * rng.uniform() - 1.0);
    double phi = 2.0 * M_PI * rng.uniform();
    directionFromAngles(theta, phi, kx, ky, kz);
}

// Random::direction(bfk, costheta), Random.cpp:188-222
__device__ __forceinline__ void scatterDirection(Philox& rng, double costheta, double& kx, double& ky, double& kz)
{
    double phi = 2.0 * M_PI * rng.uniform();
    double cosphi = cos(phi), sinphi = sin(phi);
    double sintheta = sqrt(fabs((1.0 - costheta) * (1.0 + costheta)));
    double kxn, kyn, kzn;
    if (kz > 0.99999) { kxn = cosphi * sintheta; kyn = sinphi * sintheta; kzn = costheta; }
    else if (kz < -0.99999) { kxn = cosphi * sintheta; kyn = sinphi * sintheta; kzn = -costheta; }
    else
    {
        double root = sqrt((1.0 - kz) * (1.0 + kz));
        kxn = sintheta / root * (-kx * kz * cosphi + ky * sinphi) + kx * costheta;
        kyn = -sintheta / root * (ky * kz * cosphi + kx * sinphi) + ky * costheta;
        kzn = root * sintheta * cosphi + kz * costheta;
    }
    kx = kxn; ky = kyn; kz = kzn;
}

// Random::exponcutoff, Random.cpp:162-175
__device__ __forceinline__ double exponCutoff(Philox& rng, double xmax)
{
    if (xmax == 0.0) return 0.0;
    else if (xmax < 1e-10) return rng.uniform() * xmax;
    double x = -log(1.0 - rng.uniform() * (1.0 - exp(-xmax)));
    while (x > xmax) x = -log(1.0 - rng.uniform() * (1.0 - exp(-xmax)));
    return x;
}

// NR::interpolate_loglog, NR.hpp:321-345
__device__ __forceinline__ double interpLogLog(double x, double x1, double x2, double f1, double f2)
{
    x = log10(x); x1 = log10(x1); x2 = log10(x2);
    bool logf = f1 > 0 && f2 > 0;
    if (logf) { f1 = log10(f1); f2 = log10(f2); }
    double fx = f1 + ((x - x1) / (x2 - x1)) * (f2 - f1);
    if (logf) fx = pow(10.0, fx);
    return fx;
}

// Geometry::generatePosition for the supported geometries
static __device__ void generatePosition(const SourceDev& s, Philox& rng, double& x, double& y, double& z)
{
    if (s.geometry == SKG_GEOM_EXPDISK)
    {
        // SepAxGeometry::generatePosition (SepAxGeometry.cpp:21-30) + ExpDiskGeometry::randomR/randomz (:134-161)
        const double hR = s.p[0], hz = s.p[1], Rmax = s.p[2], zmax = s.p[3], Rmin = s.p[4];
        double R, zz;
        do
        {
            double X = rng.uniform();
            R = hR * (-1.0 - lambertW1((X - 1.0) / M_E));
        }
        while ((Rmax > 0.0 && R >= Rmax) || R <= Rmin);
        double phi = 2.0 * M_PI * rng.uniform();
        do
        {
            double X = rng.uniform();
            zz = (X <= 0.5) ? hz * log(2.0 * X) : -hz * log(2.0 * (1.0 - X));
        }
        while (zmax > 0.0 && fabs(zz) >= zmax);
        x = R * cos(phi); y = R * sin(phi); z = zz;     // Position(R,phi,z,CYLINDRICAL), Position.cpp:23-31
    }
    else
    {
        // SpheGeometry::generatePosition (SpheGeometry.cpp:36-44) with SersicGeometry::randomradius (:85-91),
        // SersicFunction::inversemass (SersicFunction.cpp:112-124); SpheroidalGeometryDecorator (:78-85)
        const double reff = s.p[0], q = s.p[1];
        double X = rng.uniform();
        int Ns = s.ntab; double sval;
        if (X <= s.Xv[0]) sval = s.rv[0];
        else if (X >= s.Xv[Ns - 1]) sval = s.rv[Ns - 1];
        else
        {
            int i = locateClip(s.Xv, X, Ns);
            sval = interpLogLog(X, s.Xv[i], s.Xv[i + 1], s.rv[i], s.rv[i + 1]);
        }
        double r = reff * sval;
        double kx, ky, kz; randomDirection(rng, kx, ky, kz);
        x = r * kx; y = r * ky; z = r * kz;             // Position(r,bfk), Position.cpp:51-54
        z = q * z;
    }
    if (s.spiral_arms > 0)
    {
        // SpiralStructureGeometryDecorator::generatePosition, SpiralStructureGeometryDecorator.cpp:177-192
        double R = sqrt(x * x + y * y);
        double c = s.spiral_c;
        double phi, t;
        do
        {
            phi = 2.0 * M_PI * rng.uniform();
            // perturbation(R,phi), :224-229
            double gamma = log(R / s.spiral_radius) / s.spiral_tanp + s.spiral_phase + 0.5 * M_PI / s.spiral_arms;
            double pert = (1.0 - s.spiral_weight) + s.spiral_weight * s.spiral_cn * pow(sin(0.5 * s.spiral_arms * (gamma - phi)), 2 * s.spiral_index);
            t = rng.uniform() * c / pert;
        }
        while (t > 1);
        x = R * cos(phi); y = R * sin(phi);
    }
}

template<int KIND>
__device__ __forceinline__ int whichCellMC(const GridSetMC& G, const CartGrid& cart, double x, double y, double z)
{
    if (KIND == GRID_CART) return cartWhichCell(cart, x, y, z);
    else if (KIND == GRID_TREE) { int node = treeWhichNode(G.tree, x, y, z); return node >= 0 ? G.tree.cell[node] : -1; }
    else if (KIND == GRID_AMESH) { int node = ameshWhichNode(G.amesh, x, y, z); return node >= 0 ? G.amesh.cell[node] : -1; }
    else if (KIND == GRID_SYM) return symWhichCell(G.sym, x, y, z);
    else return voroCellIndex(G.voro, x, y, z);
}

}   // namespace skg
