// Photon shooting as a wavefront of converged stage kernels over a pool of in-flight packets:
//
//   launch     fills free pool slots with new packets            StellarSystem::launch, StellarSystem.cpp:116-158
//   peel       one traversal per (packet, observer direction)    peeloffemission / peeloffscattering,
//                                                                MonteCarloSimulation.cpp:305-363 + Instrument::detect
//   absorb     scatter (old packets), walk + absorb, terminate   simulatescattering :541-549, fillOpticalDepth +
//              or sample the interaction optical depth           simulateescapeandabsorption :438-515, :289, :519-533
//   propagate  re-walk to the sampled optical depth and move     DustGridPath::pathlength, PhotonPackage::propagate
//
// which is MonteCarloSimulation::dostellaremissionchunk (MonteCarloSimulation.cpp:265-301) turned inside out: the
// reference runs one packet through its whole life on one thread; here all packets of the pool advance one stage
// at a time, so that the 32 lanes of a warp always execute the same walker loop.  Survivors and free slots are
// compacted into index lists by warp-aggregated appends; packets are launched in wavelength order so that the
// pool holds few wavelengths at a time (the absorption table is wavelength-major on the device).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <cub/device/device_scan.cuh>
#include "mc_device.cuh"
#include "wavefront.cuh"

// resident CTAs per SM the stage kernels are compiled for (register budget 65536 / (128 x blocks)): the crossing loop
// is a chain of dependent fp64 operations, so warps in flight are what hides its latency -- measured on B200:
// absorb 205 / 176 / 163 / 187 ms per C2 phase at 3 / 4 / 5 / 6 CTAs, peel 109 / 90 / 102 (spills at 5).  The walkers of the
// hierarchical / unstructured grids carry more state: 4 CTAs (no spills)
#ifndef SKG_MC_RHO_AHEAD
#define SKG_MC_RHO_AHEAD false       // the guarded next-cell prefetch of the Cartesian walker compiled into the stage kernels
#endif
#ifndef SKG_PEEL_BATCHES
#define SKG_PEEL_BATCHES 2   // the same for the peel-off stage (no absorption atomics between the votes: measured 3 % faster with 2)
#endif
#ifndef SKG_MC_BATCHES
#define SKG_MC_BATCHES 1     // batches of SKG_PERIOD crossings between two warp votes in the stage kernels
#endif
#ifndef SKG_MC_FAST
#define SKG_MC_FAST true     // the shooting stages walk Cartesian grids with CartFastWalker (geom.cuh); false: the bit-exact walker
#endif
#ifndef SKG_PEEL_DEPTH
#define SKG_PEEL_DEPTH 4     // crossings between a density gather of the peel-off stage and its use (registers: 4 per level);
#endif                       // measured per C2 phase: 67.4 / 55.4 / 50.5 ms at depth 1 / 2 / 4 (profiles/r02_h_sweep.txt)
// (the escape + absorption stage consumes a gather at the NEXT crossing: a deeper FIFO of parked segments measured slower,
// 132.7 / 139.0 / 181.2 ms per C2 phase at depth 1 / 2 / 4, profiles/r02_i_absorb_depth_sweep.txt -- that stage is bound
// by its fp64 atomics, 53 of its 130 ms with L2 at 74 %, not by the latency of the gather)
#ifndef SKG_OTHER_MINBLOCKS
#define SKG_OTHER_MINBLOCKS 4     // resident CTAs per SM the stage kernels are compiled for on the tree / adaptive mesh / Voronoi grids
#endif
#ifndef SKG_PEEL_MINBLOCKS
#define SKG_PEEL_MINBLOCKS 4
#endif
#ifndef SKG_ABSORB_MINBLOCKS
#define SKG_ABSORB_MINBLOCKS 5
#endif
#ifndef SKG_PROP_MINBLOCKS
#define SKG_PROP_MINBLOCKS 5
#endif

namespace skg
{

// per-kernel statistics: warp-reduced, one atomic per warp and counter
__device__ __forceinline__ void flushStats(Counters* ctr, unsigned long long nSeg, unsigned long long nPaths, unsigned long long nScatt,
                                           unsigned long long nPackets, unsigned long long nAbs, unsigned long long nDet)
{
    for (int o = 16; o > 0; o >>= 1)
    {
        nSeg += __shfl_down_sync(0xffffffffu, nSeg, o); nPaths += __shfl_down_sync(0xffffffffu, nPaths, o);
        nScatt += __shfl_down_sync(0xffffffffu, nScatt, o); nPackets += __shfl_down_sync(0xffffffffu, nPackets, o);
        nAbs += __shfl_down_sync(0xffffffffu, nAbs, o); nDet += __shfl_down_sync(0xffffffffu, nDet, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (nSeg) atomicAdd(&ctr->segments, nSeg);
        if (nPaths) atomicAdd(&ctr->paths, nPaths);
        if (nScatt) atomicAdd(&ctr->scatterings, nScatt);
        if (nPackets) atomicAdd(&ctr->packets, nPackets);
        if (nAbs) atomicAdd(&ctr->absorbSegments, nAbs);
        if (nDet) atomicAdd(&ctr->detections, nDet);
    }
}

__device__ __forceinline__ void flushStageSegments(unsigned long long* counter, unsigned long long n)
{
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

// ---- launch ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) launchStage(const __grid_constant__ McDev P, Counters* ctr, int nLaunch, unsigned long long firstPacket,
                                                   int aliveBase)
{
    unsigned long long nPackets = 0;
    const int Nlambda = P.med.Nlambda;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nLaunch; j += gridDim.x * blockDim.x)
    {
        const int slot = aliveBase + j;        // appended behind the survivors
        const unsigned long long gidx = firstPacket + j;
        const int ell = P.ellList[gidx / P.NppInt];
        const unsigned long long ipkt = gidx % P.NppInt;
        // MonteCarloSimulation.cpp:267-268
        double L = __ldg(P.Ltot + ell) / P.Lscale;
        const unsigned long long id = (P.streamOffset + ipkt) * (unsigned long long)Nlambda + ell;
        Philox rng; rng.init(P.seed, id, P.rngKind);
        nPackets++;

        // ---- StellarSystem::launch, StellarSystem.cpp:116-158 ----
        int h = 0;
        if (P.Nsources > 1)
        {
            int N = P.Nsources;
            double X = rng.uniform();
            if (X < P.emissionBias) h = max(0, min(N - 1, (int)(N * X / P.emissionBias)));
            else h = locateClip(P.Lcdf + (size_t)ell * (N + 1), (X - P.emissionBias) / (1.0 - P.emissionBias), N + 1);
            double Lh = __ldg(P.L + (size_t)h * Nlambda + ell);
            if (Lh > 0)
            {
                double Lmean = __ldg(P.Ltot + ell) / N;
                double weight = 1.0 / (1.0 - P.emissionBias + P.emissionBias * Lmean / Lh);
                L = L * weight;
            }
            else L = 0;
        }
        double x = 0, y = 0, z = 0, kx = 0, ky = 0, kz = 1;
        if (L > 0)
        {
            generatePosition(P.sources[h], rng, x, y, z);        // GeometricStellarComp::launch, GeometricStellarComp.cpp:75-81
            randomDirection(rng, kx, ky, kz);                    // Geometry::generateDirection, Geometry.cpp:33
        }
        Packet* q = P.pool;
        Packet pk; pk.x = x; pk.y = y; pk.z = z; pk.kx = kx; pk.ky = ky; pk.kz = kz;
        pk.L = L; pk.target = 0; pk.id = id; pk.ell = ell; pk.nscatt = 0; pk.rngCtr = rng.c2; pk.fresh = 1; pk.hint = -1; pk.comp = h;
        storePacket(q + slot, pk);
        if (P.pol) { PolState ps; stokesUnpolarized(ps); P.pol[slot] = ps; }       // PhotonPackage::launch: setUnpolarized()
    }
    flushStats(ctr, 0, 0, 0, nPackets, 0, 0);
}

// DustGrid::randomPositionInCell: Cartesian CartesianDustGrid.cpp:129-132 (+ box(m) :333-343), tree TreeDustGrid.cpp:383-386,
// adaptive mesh AdaptiveMesh.cpp:163-167 -- Random::position(Box) draws x, y, z in this order (Random.cpp:226-234,
// Box::fracpos Box.hpp:125-126); Voronoi: rejection in the cell's enclosing box (VoronoiMesh.cpp:591-618)
template<int KIND>
__device__ __forceinline__ bool randomPositionInCell(const GridSetMC& G, int m, Philox& rng, double& x, double& y, double& z)
{
    double b[6];
    if (KIND == GRID_SYM)
    {
        const SymGrid& g = G.sym;
        if (g.sub == 0)
        {
            // Sphere1DDustGrid::randomPositionInCell, Sphere1DDustGrid.cpp:99-105: a random direction, then the radius
            double dx, dy, dz; randomDirection(rng, dx, dy, dz);
            const double r = g.v1[m] + (g.v1[m + 1] - g.v1[m]) * rng.uniform();
            x = r * dx; y = r * dy; z = r * dz;
            return true;
        }
        const int i = m / g.N2, k = m % g.N2;
        if (g.sub == 1)
        {
            // Sphere2DDustGrid::randomPositionInCell, Sphere2DDustGrid.cpp:157-167
            const double ris = g.v1[i] * g.v1[i], ri1s = g.v1[i + 1] * g.v1[i + 1];
            const double r = sqrt(ris + (ri1s - ris) * rng.uniform());
            const double theta = g.v2[k] + (g.v2[k + 1] - g.v2[k]) * rng.uniform();
            const double phi = 2.0 * M_PI * rng.uniform();
            x = r * sin(theta) * cos(phi); y = r * sin(theta) * sin(phi); z = r * cos(theta);
            return true;
        }
        // Cylinder2DDustGrid::randomPositionInCell, Cylinder2DDustGrid.cpp:120-128
        const double Rr = g.v1[i] + (g.v1[i + 1] - g.v1[i]) * rng.uniform();
        const double phi = 2.0 * M_PI * rng.uniform();
        z = g.v2[k] + (g.v2[k + 1] - g.v2[k]) * rng.uniform();
        x = Rr * cos(phi); y = Rr * sin(phi);
        return true;
    }
    if (KIND == GRID_CART)
    {
        const CartGrid& g = G.cart;
        int i = m / (g.Nz * g.Ny), j = (m / g.Nz) % g.Ny, k = m % g.Nz;
        b[0] = g.xv[i]; b[1] = g.yv[j]; b[2] = g.zv[k]; b[3] = g.xv[i + 1]; b[4] = g.yv[j + 1]; b[5] = g.zv[k + 1];
    }
    else if (KIND == GRID_TREE) { const double* nb = G.tree.box + 6 * (size_t)__ldg(G.tree.cellNode + m); for (int c = 0; c < 6; c++) b[c] = nb[c]; }
    else if (KIND == GRID_AMESH) { const double* nb = G.amesh.box + 6 * (size_t)__ldg(G.amesh.cellNode + m); for (int c = 0; c < 6; c++) b[c] = nb[c]; }
    else { const double* nb = G.voro.cellBox + 6 * (size_t)m; for (int c = 0; c < 6; c++) b[c] = nb[c]; }
    for (int attempt = 0; attempt < 10000; attempt++)
    {
        double fx = rng.uniform(), fy = rng.uniform(), fz = rng.uniform();
        x = b[0] + fx * (b[3] - b[0]); y = b[1] + fy * (b[4] - b[1]); z = b[2] + fz * (b[5] - b[2]);
        if (KIND != GRID_VORO) return true;
        // VoronoiMesh::isPointClosestTo, VoronoiMesh.cpp:610-618
        const VoroGrid& g = G.voro;
        double target = voroSD(g, m, x, y, z); bool closest = true;
        for (int q = __ldg(g.nbrStart + m); q < __ldg(g.nbrStart + m + 1) && closest; q++)
        { int id = __ldg(g.nbrIds + q); if (id >= 0 && voroSD(g, id, x, y, z) < target) closest = false; }
        if (closest) return true;
    }
    return false;       // the reference throws "Can't find random position in cell"
}

// launch of dust emission: dodustselfabsorptionchunk :208-221 / dodustemissionchunk :296-316
template<int KIND>
__global__ void __launch_bounds__(128) launchDustStage(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P, Counters* ctr, int nLaunch,
                                                       unsigned long long firstPacket, int aliveBase)
{
    unsigned long long nPackets = 0;
    const int Nlambda = P.med.Nlambda, Ncells = P.med.Ncells;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nLaunch; j += gridDim.x * blockDim.x)
    {
        const int slot = aliveBase + j;        // appended behind the survivors
        const unsigned long long gidx = firstPacket + j;
        const int ell = P.ellList[gidx / P.NppInt];
        const unsigned long long ipkt = gidx % P.NppInt;
        const double Ltot = __ldg(P.Ltot + ell);
        double L = Ltot / P.Lscale;
        const unsigned long long id = (P.streamOffset + ipkt) * (unsigned long long)Nlambda + ell;
        Philox rng; rng.init(P.seed, id, P.rngKind);
        nPackets++;
        const double* cdf = P.dustCdf + (size_t)ell * (Ncells + 1);
        int m;
        double X = rng.uniform();
        if (P.phase == SKG_PHASE_DUST_SELFABS) m = locateClip(cdf, X, Ncells + 1);
        else
        {
            const double xi = P.dustBias;
            if (X < xi) m = max(0, min(Ncells - 1, (int)(Ncells * X / xi)));
            else m = locateClip(cdf, (X - xi) / (1 - xi), Ncells + 1);
            double Lmean = Ltot / Ncells;
            double weight = 1.0 / (1 - xi + xi * Lmean / __ldg(P.dustLv + (size_t)ell * Ncells + m));
            L = L * weight;
        }
        double x = 0, y = 0, z = 0, kx = 0, ky = 0, kz = 1;
        if (!randomPositionInCell<KIND>(G, m, rng, x, y, z)) { atomicAdd(&ctr->errors, 1ull); L = 0; }
        randomDirection(rng, kx, ky, kz);
        Packet* q = P.pool;
        Packet pk; pk.x = x; pk.y = y; pk.z = z; pk.kx = kx; pk.ky = ky; pk.kz = kz;
        pk.L = L; pk.target = 0; pk.id = id; pk.ell = ell; pk.nscatt = 0; pk.rngCtr = rng.c2; pk.fresh = 1; pk.hint = -1; pk.comp = -1;
        storePacket(q + slot, pk);
        if (P.pol) { PolState ps; stokesUnpolarized(ps); P.pol[slot] = ps; }       // PhotonPackage::launch: setUnpolarized()
    }
    flushStats(ctr, 0, 0, 0, nPackets, 0, 0);
}

// NR::cdf (NR.hpp:388-394): Pv[0] = 0, Pv /= Pv[n]; the running sums Pv[1..n] come from the scan
__global__ void captureTotal(double* cdf, double* Ltot, int Ncells) { *Ltot = cdf[Ncells]; cdf[0] = 0.0; }
__global__ void normalizeCdf(double* cdf, const double* Ltot, int Ncells)
{
    const double total = *Ltot;
    for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i <= Ncells; i += gridDim.x * blockDim.x) cdf[i] = cdf[i] / total;
}

__global__ void sumOverWavelengths(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int Ncells, int Nlambda)
{
    // PanDustSystem::Labs(m), PanDustSystem.cpp:337-348: stellar table first, then the dust table, each in wavelength order
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Ncells; m += gridDim.x * blockDim.x)
    {
        double sum = 0;
        if (a) for (int ell = 0; ell < Nlambda; ell++) sum += a[(size_t)ell * Ncells + m];
        if (b) for (int ell = 0; ell < Nlambda; ell++) sum += b[(size_t)ell * Ncells + m];
        out[m] = sum;
    }
}

// ---- peel-off + detection ------------------------------------------------------------------------------------
// SingleFrameInstrument::pixelondetector, SingleFrameInstrument.cpp:130-147; -1 when the position maps off the frame
__device__ __forceinline__ int pixelOnDetector(const InstrDev& I, double x, double y, double z)
{
    double xpp = -I.sinphi * x + I.cosphi * y;
    double ypp = -I.cosphi * I.costheta * x - I.sinphi * I.costheta * y + I.sintheta * z;
    double xp = I.cospa * xpp - I.sinpa * ypp;
    double yp = I.sinpa * xpp + I.cospa * ypp;
    int i = (int)floor((xp - I.xpmin) / I.xpsiz);
    int j = (int)floor((yp - I.ypmin) / I.ypsiz);
    if (i < 0 || i >= I.Nxp || j < 0 || j >= I.Nyp) return -1;
    return i + I.Nxp * j;
}

// InstrumentFrame::pixelondetector (InstrumentFrame.cpp:153-171): the frame of the packet's wavelength, the instrument's angles
__device__ __forceinline__ long long pixelOnMultiFrame(const InstrDev& I, int ell, double x, double y, double z)
{
    const FrameDev F = I.frames[ell];
    double xpp = -I.sinphi * x + I.cosphi * y;
    double ypp = -I.cosphi * I.costheta * x - I.sinphi * I.costheta * y + I.sintheta * z;
    double xp = I.cospa * xpp - I.sinpa * ypp;
    double yp = I.sinpa * xpp + I.cospa * ypp;
    int i = (int)floor((xp - F.xpmin) / F.xpsiz);
    int j = (int)floor((yp - F.ypmin) / F.ypsiz);
    if (i < 0 || i >= F.Nxp || j < 0 || j >= F.Nyp) return -1;
    return F.offset + i + F.Nxp * j;
}
// does this instrument look at a packet at (x, y, z) before any optical depth is computed?  Frame-type detectors drop packets
// that map outside their frame first (FrameInstrument.cpp:36, InstrumentFrame.cpp:177); every other kind needs the optical depth
__device__ __forceinline__ bool instrumentRecords(const InstrDev& I, int ell, double x, double y, double z)
{
    if (I.kind == SKG_INSTR_FRAME) return pixelOnDetector(I, x, y, z) >= 0;
    if (I.kind == SKG_INSTR_MULTIFRAME) return pixelOnMultiFrame(I, ell, x, y, z) >= 0;
    return true;
}
// InstrumentFrame::detect (InstrumentFrame.cpp:175-187); returns the number of detector updates
__device__ __forceinline__ int detectMultiFrame(const InstrDev& I, int ell, double x, double y, double z, double Lextf, int comp)
{
    const long long l = pixelOnMultiFrame(I, ell, x, y, z);
    if (l < 0) return 0;
    int n = 0;
    if (I.mfTotal >= 0) { atomicAdd(I.frame + (size_t)I.mfTotal * I.mfPixels + l, Lextf); n++; }
    if (I.mfComp0 >= 0 && comp >= 0 && comp < I.mfNcomp) { atomicAdd(I.frame + (size_t)(I.mfComp0 + comp) * I.mfPixels + l, Lextf); n++; }
    return n;
}

// FullInstrument::detect (FullInstrument.cpp:107-172, unpolarised part): the flux goes to the channel of its origin --
// stellar / dust emission, direct / scattered -- plus the transparent channel (unextincted direct stellar light) and
// the channel of its scattering level.  Kept out of line: it runs once per peel-off ray, outside the crossing loop.
static __device__ __noinline__ int detectFull(const InstrDev& I, int Nlambda, double x, double y, double z, int ell,
                                              double L, double Lextf, int nscatt, bool stellar, double sQ = 0, double sU = 0, double sV = 0)
{
    const int l = pixelOnDetector(I, x, y, z);
    const size_t Nf = (size_t)I.Nxp * I.Nyp;
    int n = 0;
    auto add = [&](int c, double v)
    {
        atomicAdd(I.chanSed + (size_t)c * Nlambda + ell, v); n++;
        if (l >= 0) { atomicAdd(I.chanFrame + ((size_t)c * Nlambda + ell) * Nf + l, v); n++; }
    };
    if (stellar)
    {
        if (nscatt == 0) { add(SKG_CHAN_TRANSPARENT, L); add(SKG_CHAN_STELLAR_DIRECT, Lextf); }
        else { add(SKG_CHAN_STELLAR_SCATTERED, Lextf); if (nscatt <= I.Nscatt) add(SKG_CHAN_SCATTERING_LEVEL1 + nscatt - 1, Lextf); }
    }
    else add(nscatt == 0 ? SKG_CHAN_DUST_DIRECT : SKG_CHAN_DUST_SCATTERED, Lextf);
    // Stokes Q, U, V of the total flux (FullInstrument.cpp:136-141,165-170); packets without a scattering carry none
    if (I.pol && (sQ != 0 || sU != 0 || sV != 0))
    { const int c0 = SKG_CHAN_SCATTERING_LEVEL1 + I.Nscatt; add(c0, Lextf * sQ); add(c0 + 1, Lextf * sU); add(c0 + 2, Lextf * sV); }
    return n;
}

// One peel-off ray per (packet, observer direction): peeloffemission / peeloffscattering + Instrument::detect
template<int KIND, bool SINGLE, bool POL, bool PERSP = false> struct PeelJob
{
    static constexpr bool kCartFast = SKG_MC_FAST; static constexpr bool kCartRegBorders = SKG_CART_REGBORDERS, kCartTinySelect = true; static constexpr bool kTreeHints = SKG_TREE_HINTS_MC, kCartRhoAhead = SKG_MC_RHO_AHEAD; static constexpr int kBatches = SKG_PEEL_BATCHES;
    const GridSetMC& G; const CartGrid& cart; const McDev& P;
    double rx, ry, rz, dx, dy, dz;          // the ray (runJobs interface): only read by the walker's start(), not kept over the walk
    // state carried over the walk, kept small (registers are what limits the warps in flight): everything else the
    // detection needs (position, wavelength, scattering count) is re-read from the packet record in finish()
    double Lw, tau; int item, ell, hint;
    // one-component media: the density gather of a crossing is consumed one crossing later, so that its latency
    // overlaps the next step's arithmetic (same summation order: tau += (kext*rho[m])*ds per segment)
    // one-component media: the density gather of a crossing is consumed kDepth crossings later (the optical depth is a
    // plain sum: the order of its terms is free), so that the L2 latency of the gather overlaps that many crossings
    static constexpr int kDepth = SKG_PEEL_DEPTH;
    double kext0, pendRho[kDepth], pendDs[kDepth];
    static constexpr bool single = SINGLE;  // one dust component: compile-time, no branch in the crossing loop
    double sedAdd;                          // what finish() leaves for collective(): the extincted luminosity for the SED bins
    // PERSP (one item per packet and PerspectiveInstrument): the walk ends with the first segment beyond the distance to the viewport
    // plane (DustGridPath::opticalDepth(kapparho, distance), DustGridPath.hpp:97-108); the pixel found in begin()
    double sacc, smax; int pix;
    unsigned nSeg = 0, nPaths = 0, nDet = 0;
    __device__ PeelJob(const GridSetMC& G_, const CartGrid& c_, const McDev& P_) : G(G_), cart(c_), P(P_) {}

    __device__ __forceinline__ int begin(int it)
    {
        item = it;
        const int ngroups = PERSP ? P.Npersp : P.Ngroups;
        const int slot = it / ngroups;
        ObsGroup gl; gl.first = 0; gl.count = 0;
        const ObsGroup& g = PERSP ? gl : P.groups[it % ngroups];
        const Packet pk = loadPacket(P.pool + slot);
        double L = pk.L;
        if (!(L > 0)) return 0;                                 // MonteCarloSimulation.cpp:281
        if (P.contScatt && !pk.fresh) return 0;                 // continuous scattering: no peel-off at the interaction points (:291)
        rx = pk.x; ry = pk.y; rz = pk.z;
        ell = pk.ell; hint = KIND == GRID_CART ? -1 : pk.hint;
        if constexpr (PERSP)
        {
            const PerspDev& V = P.persp[it % ngroups];
            // PerspectiveInstrument::detect (PerspectiveInstrument.cpp:322-334): world -> pixel coordinates; packets that arrive outside
            // the viewport or come from behind / very close to it are ignored before any optical depth is computed
            const double xp = rx * V.M[0][0] + ry * V.M[1][0] + rz * V.M[2][0] + V.M[3][0], yp = rx * V.M[0][1] + ry * V.M[1][1] + rz * V.M[2][1] + V.M[3][1];
            const double zp = rx * V.M[0][2] + ry * V.M[1][2] + rz * V.M[2][2] + V.M[3][2], wp = rx * V.M[0][3] + ry * V.M[1][3] + rz * V.M[2][3] + V.M[3][3];
            const int i = (int)(xp / wp), j = (int)(yp / wp);
            if (!(i >= 0 && i < V.Nx && j >= 0 && j < V.Ny && zp > V.s / 10.)) return 0;
            pix = i + V.Nx * j; smax = zp; sacc = 0;
            // bfkobs(bfr) (:290-304): from the packet towards the eye
            const double ex = V.Ex - rx, ey = V.Ey - ry, ez = V.Ez - rz, D = sqrt(ex * ex + ey * ey + ez * ez);
            if (D < 1e-20) { gl.kx = 0; gl.ky = 0; gl.kz = 1; } else { gl.kx = ex / D; gl.ky = ey / D; gl.kz = ez / D; }
        }
        // which instruments of this direction record the packet?  FrameInstrument ignores packets that map outside
        // its frame before any optical depth is computed (FrameInstrument.cpp:36); SED/Simple always need tau
        bool need = PERSP;
        for (int c = 0; c < g.count; c++)
        {
            if (instrumentRecords(P.instr[g.first + c], ell, rx, ry, rz)) { need = true; break; }
        }
        if (!need) return 0;
        const int Ncomp = P.med.Ncomp, Nlambda = P.med.Nlambda;
        if (!pk.fresh)
        {
            // ---- peeloffscattering, MonteCarloSimulation.cpp:319-363: weight by the phase function towards the observer ----
            const double kx = pk.kx, ky = pk.ky, kz = pk.kz;
            const double cosalpha = kx * g.kx + ky * g.ky + kz * g.kz;          // Direction::dot
            double w = 0;
            if constexpr (POL)
            {
                // DustMix::phaseFunctionValue with polarisation, DustMix.cpp:650-662, summed over the components weighted by
                // kappasca*rho (MonteCarloSimulation.cpp:325-336); the Stokes vector for each instrument follows in finish()
                const PolState ps = P.pol[slot];
                const double phi = anglePlanes(ps.nx, ps.ny, ps.nz, kx, ky, kz, g.kx, g.ky, g.kz);
                const double theta = acos(cosalpha);
                const int t = indexForTheta(theta, P.med.Ntheta);
                const double polDegree = stokesLinearDegree(ps), polAngle = stokesAngle(ps);
                double wv[8]; wv[0] = 1.0;
                if (!SINGLE)
                {
                    int mcell = whichCellMC<KIND>(G, cart, rx, ry, rz);
                    if (mcell == -1) return 0;
                    double sum = 0;
                    for (int c = 0; c < Ncomp && c < 8; c++)
                    { wv[c] = __ldg(P.med.ksca + (size_t)c * Nlambda + ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c); sum += wv[c]; }
                    if (sum <= 0) return 0;
                    for (int c = 0; c < Ncomp && c < 8; c++) wv[c] /= sum;
                }
                for (int c = 0; c < Ncomp && c < 8; c++)
                {
                    const size_t o = ((size_t)c * Nlambda + ell) * P.med.Ntheta + t;
                    w += wv[c] * (P.med.pfnorm[(size_t)c * Nlambda + ell] * (P.med.S11[o] + polDegree * P.med.S12[o] * cos(2. * (phi - polAngle))));
                }
            }
            else if (SINGLE)
            {
                // one component: weight 1 (MonteCarloSimulation.cpp:325-326); DustMix::phaseFunctionValue (HG), DustMix.cpp:665-668
                const double gg = __ldg(P.med.g + ell);
                const double tt = 1.0 + gg * gg - 2 * gg * cosalpha;
                w = (1.0 - gg) * (1.0 + gg) / sqrt(tt * tt * tt);
            }
            else
            {
                double wv[8];
                int mcell = whichCellMC<KIND>(G, cart, rx, ry, rz);
                if (mcell == -1) return 0;
                double sum = 0;
                for (int c = 0; c < Ncomp && c < 8; c++)
                { wv[c] = __ldg(P.med.ksca + (size_t)c * Nlambda + ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c); sum += wv[c]; }
                if (sum <= 0) return 0;
                for (int c = 0; c < Ncomp && c < 8; c++) wv[c] /= sum;
                for (int c = 0; c < Ncomp && c < 8; c++)
                {
                    double gg = __ldg(P.med.g + (size_t)c * Nlambda + ell);
                    double tt = 1.0 + gg * gg - 2 * gg * cosalpha;
                    w += wv[c] * ((1.0 - gg) * (1.0 + gg) / sqrt(tt * tt * tt));
                }
            }
            L = L * w;                                              // launchScatteringPeelOff, PhotonPackage.cpp:51-62
        }
        Lw = L; tau = 0;
        dx = g.kx; dy = g.ky; dz = g.kz;
        kext0 = single ? __ldg(P.med.kext + ell) : 0.0;
#pragma unroll
        for (int q = 0; q < kDepth; q++) { pendRho[q] = 0; pendDs[q] = 0; }
        if (!P.med.rho) return 2;                                   // Instrument::opticalDepth: 0 without dust
        nPaths++;
        return 1;
    }
    // The Stokes vector of a scattering peel-off packet as instrument I records it: per component
    // DustMix::scatteringPeelOffPolarization (DustMix.cpp:619-644) weighted by w_h = wv[h] * phaseFunctionValue, summed and
    // normalised by I = sum w_h (MonteCarloSimulation.cpp:337-352, PhotonPackage::setPolarized).  Out of line: once per detection.
    __device__ __noinline__ void peelStokes(const InstrDev& I, const ObsGroup& g, const Packet& pk, const PolState& pp, double& sQ, double& sU, double& sV) const
    {
        const int Ncomp = P.med.Ncomp, Nlambda = P.med.Nlambda, Nt = P.med.Ntheta;
        const double kx = pk.kx, ky = pk.ky, kz = pk.kz;
        const double cosalpha = kx * g.kx + ky * g.ky + kz * g.kz;
        const double theta = acos(cosalpha); const int t = indexForTheta(theta, Nt);
        const double phi0 = anglePlanes(pp.nx, pp.ny, pp.nz, kx, ky, kz, g.kx, g.ky, g.kz);
        const double polDegree = stokesLinearDegree(pp), polAngle = stokesAngle(pp);
        double wv[8]; wv[0] = 1.0;
        if (Ncomp > 1)
        {
            const int mcell = whichCellMC<KIND>(G, cart, pk.x, pk.y, pk.z);
            double sum = 0;
            for (int c = 0; c < Ncomp && c < 8; c++)
            { wv[c] = mcell >= 0 ? __ldg(P.med.ksca + (size_t)c * Nlambda + pk.ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c) : 0.0; sum += wv[c]; }
            for (int c = 0; c < Ncomp && c < 8; c++) wv[c] = sum > 0 ? wv[c] / sum : 0.0;
        }
        double Isum = 0, Q = 0, U = 0, V = 0;
        for (int c = 0; c < Ncomp && c < 8; c++)
        {
            const size_t o = ((size_t)c * Nlambda + pk.ell) * Nt + t;
            const double w = wv[c] * (P.med.pfnorm[(size_t)c * Nlambda + pk.ell] * (P.med.S11[o] + polDegree * P.med.S12[o] * cos(2. * (phi0 - polAngle))));
            PolState sv = pp;
            stokesRotate(sv, 0.0, kx, ky, kz);                                        // generates the normal at a first scattering
            const double phi = anglePlanes(sv.nx, sv.ny, sv.nz, kx, ky, kz, g.kx, g.ky, g.kz);
            stokesRotate(sv, phi, kx, ky, kz);                                        // into the peel-off scattering plane
            stokesMueller(sv, P.med.S11[o], P.med.S12[o], P.med.S33[o], P.med.S34[o]);
            const double alpha = angleInstrument(sv.nx, sv.ny, sv.nz, g.kx, g.ky, g.kz, I.kyx, I.kyy, I.kyz);
            stokesRotate(sv, alpha, kx, ky, kz);                                      // into the instrument's frame
            Isum += w; Q += w * sv.Q; U += w * sv.U; V += w * sv.V;
        }
        if (Isum != 0.0) { sQ = Q / Isum; sU = U / Isum; sV = V / Isum; }
    }
    __device__ __forceinline__ int cellHint() const { return hint; }
    // the first traversal from a position establishes where it lies: remembered in the packet record for the next ones
    __device__ __forceinline__ void noteStart(int loc) { if (KIND != GRID_CART) P.pool[item / (PERSP ? P.Npersp : P.Ngroups)].hint = loc; }
    __device__ __forceinline__ bool outside(double d) { nSeg++; if constexpr (PERSP) { sacc += d; return !(sacc > smax); } return true; }
    template<int U> __device__ __forceinline__ bool segmentU(int m, double ds)
    {
        nSeg++;
        constexpr int q = U % kDepth;
        if (single) { tau += (kext0 * pendRho[q]) * pendDs[q]; pendRho[q] = __ldg(P.med.rho + m); pendDs[q] = ds; }
        else tau += KappaRho{P.med.rho, P.med.kext + ell, P.med.Ncomp, P.med.Nlambda}(m) * ds;
        if constexpr (PERSP) { sacc += ds; return !(sacc > smax); }
        return true;
    }
    __device__ __forceinline__ bool segment(int m, double ds) { return segmentU<0>(m, ds); }
    template<int U> __device__ __forceinline__ void idleU() {}
    __device__ __forceinline__ void finish()
    {
        if (single)
        {
#pragma unroll
            for (int q = 0; q < kDepth; q++) tau += (kext0 * pendRho[q]) * pendDs[q];
        }
        const double Lextf = Lw * exp(-tau);
        sedAdd = Lextf;
        if constexpr (PERSP)
        {
            // PerspectiveInstrument.cpp:336-350: the luminosity adjusted for the distance from the launch position to the viewport
            const PerspDev& V = P.persp[item % P.Npersp];
            const double r = V.s / (2. * smax), rar = r / atan(r);
            atomicAdd(V.frame + (size_t)pix + (size_t)V.Nx * V.Ny * ell, Lextf * (rar * rar)); nDet++;
            return;
        }
        const ObsGroup& g = P.groups[item % P.Ngroups];
        // position and scattering count of the peel-off packet (launchEmissionPeelOff / launchScatteringPeelOff,
        // PhotonPackage.cpp:34-62: 0 for emission, else previous scatterings + 1) from the packet record
        const Packet* q = P.pool + item / P.Ngroups;
        const double px = q->x, py = q->y, pz = q->z;
        for (int c = 0; c < g.count; c++)
        {
            const InstrDev& I = P.instr[g.first + c];
            if (I.kind == SKG_INSTR_FULL)
            {
                const int ns = q->fresh ? 0 : q->nscatt + 1;
                double sQ = 0, sU = 0, sV = 0;
                if constexpr (POL) { if (I.pol && !q->fresh) peelStokes(I, g, *q, P.pol[item / P.Ngroups], sQ, sU, sV); }
                nDet += detectFull(I, P.med.Nlambda, px, py, pz, ell, Lw, Lextf, ns, P.phase == SKG_PHASE_STELLAR, sQ, sU, sV); continue;
            }
            if (I.kind == SKG_INSTR_MULTIFRAME) { nDet += detectMultiFrame(I, ell, px, py, pz, Lextf, P.phase == SKG_PHASE_STELLAR ? q->comp : -1); continue; }
            // SEDInstrument::detect SEDInstrument.cpp:32-42, FrameInstrument::detect FrameInstrument.cpp:32-47, SimpleInstrument.cpp:33-49
            // (the SED bins are served by collective(): one address per wavelength, summed over the converged warp)
            if (I.kind != SKG_INSTR_FRAME) nDet++;
            if (I.kind != SKG_INSTR_SED)
            {
                int l = pixelOnDetector(I, px, py, pz);
                if (l >= 0) { atomicAdd(I.frame + (size_t)l + (size_t)ell * I.Nxp * I.Nyp, Lextf); nDet++; }
            }
        }
    }
    // SED bins (SEDInstrument::detect SEDInstrument.cpp:32-42, SimpleInstrument.cpp:33-49), warp-uniformly: up to
    // maxCount instruments per observer direction, every lane taking part in every round
    __device__ __forceinline__ void collective(bool fin)
    {
        if constexpr (PERSP) return;
        if (!__any_sync(0xffffffffu, fin)) return;
        const ObsGroup& g = P.groups[fin ? item % P.Ngroups : 0];
        const int rounds = P.maxGroupCount;
        for (int c = 0; c < rounds; c++)
        {
            const bool has = fin && c < g.count;
            const InstrDev& I = P.instr[has ? g.first + c : 0];
            const bool take = has && I.kind != SKG_INSTR_FRAME && I.kind != SKG_INSTR_FULL && I.kind != SKG_INSTR_MULTIFRAME;
            warpConvergedAdd(take, take ? I.sed + ell : nullptr, sedAdd);
        }
    }
    __device__ __forceinline__ void periodic() {}
};

template<int KIND, bool SINGLE, bool POL, bool PERSP = false>
__global__ void __launch_bounds__(128, KIND == GRID_CART ? SKG_PEEL_MINBLOCKS : SKG_OTHER_MINBLOCKS) peelStage(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P, Counters* ctr, bool cartSmem,
                                                 int nAlive, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    PeelJob<KIND, SINGLE, POL, PERSP> job(G, cart, P);
    runJobs<KIND>(G, cart, ctr, job, nAlive * (PERSP ? P.Npersp : P.Ngroups), work, P.peelRefill);
    flushStats(ctr, job.nSeg, job.nPaths, 0, 0, 0, job.nDet);
    flushStageSegments(&ctr->peelSegments, job.nSeg);
}

// MonteCarloSimulation::continuouspeeloffscattering (MonteCarloSimulation.cpp:367-434; continuousScattering = true): instead of
// one peel-off at every interaction point, EVERY segment of the packet's path sends a peel-off packet from a random
// position inside it towards every instrument, weighted by the fraction of the luminosity scattered in that segment,
// albedo * exp(-tau_start) * (-expm1(-dtau)), and by the phase function.  One item per (packet, observer direction): the
// packet's own path is walked by the scheduler, and each of its segments walks the peel-off ray with a second, nested
// walker.  Far more traversals per packet than the standard mode -- as in the reference.
template<int KIND> struct WalkerOf;
template<> struct WalkerOf<GRID_CART> { typedef CartFastWalkerT<false> type; };
template<> struct WalkerOf<GRID_TREE> { typedef TreeWalkerT<SKG_TREE_HINTS_MC> type; };
template<> struct WalkerOf<GRID_AMESH> { typedef AMeshWalker type; };
template<> struct WalkerOf<GRID_VORO> { typedef VoroWalkerT<false> type; };
template<> struct WalkerOf<GRID_SYM> { typedef SymWalker type; };

template<int KIND, bool SINGLE, bool POL> struct ContPeelJob
{
    static constexpr bool kCartFast = SKG_MC_FAST; static constexpr bool kCartRegBorders = SKG_CART_REGBORDERS, kCartTinySelect = false; static constexpr bool kTreeHints = SKG_TREE_HINTS_MC, kCartRhoAhead = false; static constexpr int kBatches = 1;
    const GridSetMC& G; const CartGrid& cart; const McDev& P; Counters* ctr;
    double rx, ry, rz, dx, dy, dz;
    double L, tau, sacc, w0; int item, ell, hint; Philox rng;
    unsigned nSeg = 0, nPaths = 0, nDet = 0;
    __device__ ContPeelJob(const GridSetMC& G_, const CartGrid& c_, const McDev& P_, Counters* ctr_) : G(G_), cart(c_), P(P_), ctr(ctr_) {}

    // the phase function towards the observer for dust component c (HG: DustMix.cpp:665-668; polarised: :650-662)
    __device__ __forceinline__ double phase(int c, const Packet& pk, const ObsGroup& g) const
    {
        const int Nlambda = P.med.Nlambda;
        const double cosalpha = pk.kx * g.kx + pk.ky * g.ky + pk.kz * g.kz;
        if constexpr (POL)
        {
            const PolState ps = P.pol[item / P.Ngroups];
            const double phi = anglePlanes(ps.nx, ps.ny, ps.nz, pk.kx, pk.ky, pk.kz, g.kx, g.ky, g.kz);
            const size_t o = ((size_t)c * Nlambda + ell) * P.med.Ntheta + indexForTheta(acos(cosalpha), P.med.Ntheta);
            return P.med.pfnorm[(size_t)c * Nlambda + ell] * (P.med.S11[o] + stokesLinearDegree(ps) * P.med.S12[o] * cos(2. * (phi - stokesAngle(ps))));
        }
        const double gg = __ldg(P.med.g + (size_t)c * Nlambda + ell);
        const double tt = 1.0 + gg * gg - 2 * gg * cosalpha;
        return (1.0 - gg) * (1.0 + gg) / sqrt(tt * tt * tt);
    }
    __device__ __forceinline__ int begin(int it)
    {
        item = it;
        const Packet pk = loadPacket(P.pool + it / P.Ngroups);
        L = pk.L;
        if (!(L > 0) || !P.med.rho) return 0;
        rx = pk.x; ry = pk.y; rz = pk.z; dx = pk.kx; dy = pk.ky; dz = pk.kz;
        ell = pk.ell; hint = KIND == GRID_CART ? -1 : pk.hint;
        tau = 0; sacc = 0;
        w0 = SINGLE ? phase(0, pk, P.groups[it % P.Ngroups]) : 0.0;
        // the positions inside the segments come from a Philox stream of their own: (packet, scattering order, direction)
        rng.init(P.seed, pk.id, P.rngKind + 8u); rng.c2 = ((unsigned)pk.nscatt * (unsigned)P.Ngroups + (unsigned)(it % P.Ngroups)) << 10;
        nPaths++;
        return 1;
    }
    __device__ __forceinline__ int cellHint() const { return hint; }
    __device__ __forceinline__ void noteStart(int) {}
    __device__ __forceinline__ bool outside(double ds) { nSeg++; sacc += ds; return true; }
    template<int U> __device__ __forceinline__ bool segmentU(int m, double ds) { return segment(m, ds); }
    template<int U> __device__ __forceinline__ void idleU() {}
    __device__ __noinline__ bool segment(int m, double ds)
    {
        nSeg++;
        const int Ncomp = P.med.Ncomp, Nlambda = P.med.Nlambda;
        const ObsGroup& g = P.groups[item % P.Ngroups];
        const Packet* q = P.pool + item / P.Ngroups;
        // scattering and extinction opacity of the cell, and the phase function averaged over the components (:392-405, :418-428)
        double ksca = 0.0, kext = 0.0, w = 0.0;
        if (SINGLE) { const double rho = __ldg(P.med.rho + m); ksca = rho * __ldg(P.med.ksca + ell); kext = rho * __ldg(P.med.kext + ell); w = w0; }
        else
        {
            const Packet pk = loadPacket(q);
            for (int h = 0; h < Ncomp; h++)
            {
                const double rho = __ldg(P.med.rho + (size_t)m * Ncomp + h), ks = rho * __ldg(P.med.ksca + (size_t)h * Nlambda + ell);
                ksca += ks; kext += rho * __ldg(P.med.kext + (size_t)h * Nlambda + ell);
                if (ks > 0) w += ks * phase(h, pk, g);
            }
            if (ksca > 0) w /= ksca;
        }
        const double dtau = kext * ds;
        if (ksca > 0.0)
        {
            const double albedo = ksca / kext;
            const double factorm = albedo * exp(-tau) * (-expm1(-dtau));
            const double s = sacc + rng.uniform() * ds;
            const double px = rx + s * dx, py = ry + s * dy, pz = rz + s * dz;             // bfrnew
            // which instruments of this direction record a packet from there?  (FrameInstrument.cpp:36)
            bool need = false;
            for (int c = 0; c < g.count; c++)
            if (instrumentRecords(P.instr[g.first + c], ell, px, py, pz)) { need = true; break; }
            if (need)
            {
                // Instrument::opticalDepth of the peel-off packet: a traversal of its own, with a second walker
                typedef typename WalkerOf<KIND>::type W2;
                W2 w2; Entry en; double tau2 = 0;
                const int loc = KIND == GRID_TREE ? __ldg(G.tree.cellNode + m) : (KIND == GRID_AMESH ? __ldg(G.amesh.cellNode + m) : (KIND == GRID_VORO ? m : -1));
                bool ok;
                if constexpr (KIND == GRID_CART) ok = w2.start(cart, ctr, px, py, pz, g.kx, g.ky, g.kz, en, -1);
                else if constexpr (KIND == GRID_TREE) ok = w2.start(G.tree, ctr, px, py, pz, g.kx, g.ky, g.kz, en, loc);
                else if constexpr (KIND == GRID_AMESH) ok = w2.start(G.amesh, ctr, px, py, pz, g.kx, g.ky, g.kz, en, loc);
                else if constexpr (KIND == GRID_SYM) ok = w2.start(G.sym, ctr, px, py, pz, g.kx, g.ky, g.kz, en, -1);
                else ok = w2.start(G.voro, ctr, px, py, pz, g.kx, g.ky, g.kz, en, loc);
                if (ok)
                {
                    nPaths++;
                    while (w2.alive)
                    {
                        int m2; double ds2; bool seg;
                        if constexpr (KIND == GRID_CART) seg = w2.step(cart, ctr, m2, ds2);
                        else if constexpr (KIND == GRID_TREE) seg = w2.step(G.tree, ctr, m2, ds2);
                        else if constexpr (KIND == GRID_AMESH) seg = w2.step(G.amesh, ctr, m2, ds2);
                        else if constexpr (KIND == GRID_SYM) seg = w2.step(G.sym, ctr, m2, ds2);
                        else seg = w2.step(G.voro, ctr, m2, ds2);
                        if (seg) { nSeg++; tau2 += KappaRho{P.med.rho, P.med.kext + ell, Ncomp, Nlambda}(m2) * ds2; }
                        // (walkers whose crossing comes in two halves: geom.cuh, kSplitStep)
                        if constexpr (KIND == GRID_TREE) w2.resolve(G.tree, ctr);
                        else if constexpr (KIND == GRID_AMESH) w2.resolve(G.amesh, ctr);
                    }
                }
                const double Lw = L * (factorm * w);                                       // launchScatteringPeelOff(pp, bfrnew, bfkobs, factorm*I)
                const double Lextf = Lw * exp(-tau2);
                for (int c = 0; c < g.count; c++)
                {
                    const InstrDev& I = P.instr[g.first + c];
                    if (I.kind == SKG_INSTR_FULL)
                    {
                        double sQ = 0, sU = 0, sV = 0;
                        if constexpr (POL) { if (I.pol) { PeelJob<KIND, SINGLE, POL> pj(G, cart, P); pj.peelStokes(I, g, loadPacket(q), P.pol[item / P.Ngroups], sQ, sU, sV); } }
                        nDet += detectFull(I, Nlambda, px, py, pz, ell, Lw, Lextf, q->nscatt + 1, P.phase == SKG_PHASE_STELLAR, sQ, sU, sV); continue;
                    }
                    if (I.kind == SKG_INSTR_MULTIFRAME) { nDet += detectMultiFrame(I, ell, px, py, pz, Lextf, P.phase == SKG_PHASE_STELLAR ? q->comp : -1); continue; }
                    if (I.kind != SKG_INSTR_FRAME) { atomicAdd(I.sed + ell, Lextf); nDet++; }
                    if (I.kind != SKG_INSTR_SED)
                    {
                        const int l = pixelOnDetector(I, px, py, pz);
                        if (l >= 0) { atomicAdd(I.frame + (size_t)l + (size_t)ell * I.Nxp * I.Nyp, Lextf); nDet++; }
                    }
                }
            }
        }
        tau += dtau; sacc += ds;
        return true;
    }
    __device__ __forceinline__ void finish() {}
    __device__ __forceinline__ void collective(bool) {}
    __device__ __forceinline__ void periodic() {}
};

template<int KIND, bool SINGLE, bool POL>
__global__ void __launch_bounds__(128, 3) contPeelStage(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P, Counters* ctr, bool cartSmem, int nAlive, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    ContPeelJob<KIND, SINGLE, POL> job(G, cart, P, ctr);
    runJobs<KIND>(G, cart, ctr, job, nAlive * P.Ngroups, work, 8);
    flushStats(ctr, job.nSeg, job.nPaths, 0, 0, 0, job.nDet);
    flushStageSegments(&ctr->peelSegments, job.nSeg);
}

// scatter (packets that come from an interaction) + escape/absorption + termination + interaction sampling
template<int KIND, bool SINGLE, bool STORE, bool POL> struct AbsorbJob
{
    static constexpr bool kCartFast = SKG_MC_FAST; static constexpr bool kCartRegBorders = SKG_CART_REGBORDERS, kCartTinySelect = false; static constexpr bool kTreeHints = SKG_TREE_HINTS_MC, kCartRhoAhead = SKG_MC_RHO_AHEAD; static constexpr int kBatches = SKG_MC_BATCHES;
    const GridSetMC& G; const CartGrid& cart; const McDev& P;
    int* counts;
    double rx, ry, rz, dx, dy, dz;          // the ray (runJobs interface): only read by the walker's start(), not kept over the walk
    // escape + absorption state carried over the walk, kept small (registers are what limits the warps in flight); the
    // packet's other fields are re-read from its record when the path has ended.  The reference evaluates
    // L*exp(-tau_start)*(-expm1(-dtau)) per segment (MonteCarloSimulation.cpp:452); here the attenuation
    // E = exp(-tau_start) is carried along multiplicatively, E += E*expm1(-dtau): one transcendental per segment instead
    // of two (the forms agree to a few ulp over a path)
    double Labs0;                           // (1 - albedo) * L of the packet: what a fully absorbing segment would take (one component)
    double tau, E, Lsca; double* labs;
    int slot, ell, hint; bool walked, survive;
    // one-component media: a crossing parks (m, ds) and starts the gather of the cell's density; the entry is consumed at
    // the next crossing, in path order (the attenuation E must be that at the START of each segment)
    double kext0, pendDs, pendRho; int pendM;
    // what finish() hands to collective() (alive only during a refill, not over the walk)
    double Lout, target; unsigned rngOut;
    unsigned nSeg = 0, nPaths = 0, nScatt = 0, nAbs = 0;
    __device__ AbsorbJob(const GridSetMC& G_, const CartGrid& c_, const McDev& P_, int* c2_) : G(G_), cart(c_), P(P_), counts(c2_) {}

    __device__ __forceinline__ int begin(int item)
    {
        Packet* q = P.pool + item;
        slot = item;
        survive = false; walked = false;
        Packet pk = loadPacket(q);
        const double L = pk.L;
        if (!(L > 0) || !P.med.rho) return 2;       // nothing to propagate: the slot is simply not copied to the next pool
        const int Ncomp = P.med.Ncomp, Nlambda = P.med.Nlambda;
        ell = pk.ell; hint = KIND == GRID_CART ? -1 : pk.hint;
        rx = pk.x; ry = pk.y; rz = pk.z;
        dx = pk.kx; dy = pk.ky; dz = pk.kz;
        if (!pk.fresh)
        {
            // ---- simulatescattering, MonteCarloSimulation.cpp:541-549 ----
            Philox rng; rng.init(P.seed, pk.id, P.rngKind); rng.c2 = pk.rngCtr;
            int hmix = 0;
            if (Ncomp > 1)
            {
                // DustSystem::randomMixForPosition, DustSystem.cpp:879-893
                int mcell = whichCellMC<KIND>(G, cart, rx, ry, rz);
                if (mcell >= 0)
                {
                    double Xv[9]; Xv[0] = 0;
                    for (int c = 0; c < Ncomp && c < 8; c++)
                        Xv[c + 1] = Xv[c] + __ldg(P.med.ksca + (size_t)c * Nlambda + ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c);
                    double tot = Xv[Ncomp];
                    for (int c = 0; c <= Ncomp; c++) Xv[c] /= tot;
                    hmix = locateClip(Xv, rng.uniform(), Ncomp + 1);
                }
            }
            if constexpr (POL)
            {
                // DustMix::scatteringDirectionAndPolarization with polarisation, DustMix.cpp:586-605
                PolState ps = P.pol[slot];
                const double theta = sampleTheta(P.med, hmix, ell, rng.uniform());
                const double phi = samplePhi(P.med, hmix, ell, theta, stokesLinearDegree(ps), stokesAngle(ps), rng.uniform());
                stokesRotate(ps, phi, dx, dy, dz);                                    // the Stokes vector and the scattering plane
                const size_t o = ((size_t)hmix * Nlambda + ell) * P.med.Ntheta + indexForTheta(theta, P.med.Ntheta);
                stokesMueller(ps, P.med.S11[o], P.med.S12[o], P.med.S33[o], P.med.S34[o]);
                // the propagation direction turns by theta in the scattering plane: k cos(theta) + (normal x k) sin(theta)
                const double ct = cos(theta), st = sin(theta);
                const double cx = ps.ny * dz - ps.nz * dy, cy = ps.nz * dx - ps.nx * dz, cz = ps.nx * dy - ps.ny * dx;
                double nx = dx * ct + cx * st, ny = dy * ct + cy * st, nz = dz * ct + cz * st;
                const double norm = sqrt(nx * nx + ny * ny + nz * nz);
                dx = nx / norm; dy = ny / norm; dz = nz / norm;
                P.pol[slot] = ps;
            }
            else
            {
                // DustMix::scatteringDirectionAndPolarization (HG branch), DustMix.cpp:607-614
                double g = __ldg(P.med.g + (size_t)hmix * Nlambda + ell);
                if (fabs(g) < 1e-6) randomDirection(rng, dx, dy, dz);
                else
                {
                    double f = ((1.0 - g) * (1.0 + g)) / (1.0 - g + 2.0 * g * rng.uniform());
                    double costheta = (1.0 + g * g - f * f) / (2.0 * g);
                    scatterDirection(rng, costheta, dx, dy, dz);
                }
            }
            nScatt++;
            // the scattered packet goes back to its record (new direction, one more scattering, stream position): the walk
            // keeps none of it in registers
            pk.kx = dx; pk.ky = dy; pk.kz = dz; pk.nscatt++; pk.rngCtr = rng.c2; pk.fresh = 0;
            storePacket(q, pk);
        }
        // ---- fillOpticalDepth + simulateescapeandabsorption, :286-288, :438-515 ----
        labs = P.labs ? P.labs + (size_t)ell * P.med.Ncells : nullptr;
        kext0 = __ldg(P.med.kext + ell);
        const double albedo = kext0 > 0 ? __ldg(P.med.ksca + ell) / kext0 : 0.0;        // DustMix::albedo(ell) (DustMix.cpp:55-90)
        Labs0 = SINGLE ? (1.0 - albedo) * L : L;
        tau = 0; E = 1.0; Lsca = 0;
        pendM = -1; pendDs = 0; pendRho = 0;
        nPaths++; walked = true;
        return 1;
    }
    __device__ __forceinline__ int cellHint() const { return hint; }
    __device__ __forceinline__ void noteStart(int loc) { if (KIND != GRID_CART) P.pool[slot].hint = loc; }
    __device__ __forceinline__ bool outside(double) { nSeg++; return true; }   // rho(-1,h) = 0: dtau = 0, nothing absorbed
    // escape + absorption of the parked segment in a one-component medium (MonteCarloSimulation.cpp:446-470)
    __device__ __forceinline__ void absorbPending()
    {
        if (pendM < 0) return;
        double dtau = (kext0 * pendRho) * pendDs;
        if (STORE)
        {
            double x = expm1Small(-dtau);
#ifdef SKG_EXP_NOATOMIC
            Lsca += Labs0 * (E * (-x));         // experiment: no absorption atomics
#else
            atomicAdd(labs + pendM, Labs0 * (E * (-x)));
#endif
            E += E * x;
            nAbs++;
        }
        tau += dtau;
        pendM = -1;
    }
    __device__ __forceinline__ bool segment(int m, double ds)
    {
        nSeg++;
        if (SINGLE)
        {
            absorbPending();
            pendM = m; pendDs = ds; pendRho = __ldg(P.med.rho + m);
            return true;
        }
        return segmentMulti(m, ds);
    }
    template<int U> __device__ __forceinline__ bool segmentU(int m, double ds) { return segment(m, ds); }
    template<int U> __device__ __forceinline__ void idleU() {}
    __device__ __forceinline__ bool segmentMulti(int m, double ds)
    {
        const int Ncomp = P.med.Ncomp;
        double ksca = 0.0, kext = 0.0, krr = 0.0;
        for (int h = 0; h < Ncomp; h++)
        {
            double rho = __ldg(P.med.rho + (size_t)m * Ncomp + h);
            ksca += rho * __ldg(P.med.ksca + (size_t)h * P.med.Nlambda + ell);
            double ke = __ldg(P.med.kext + (size_t)h * P.med.Nlambda + ell);
            kext += rho * ke;
            krr += ke * rho;
        }
        double alb = (kext > 0.0) ? ksca / kext : 0.0;
        double dtau = krr * ds;
        double x = expm1Small(-dtau);
        double Lintm = Labs0 * E * (-x);        // Labs0 = L for several components
        E += E * x;
        Lsca += alb * Lintm;
        if (STORE) { atomicAdd(labs + m, (1.0 - alb) * Lintm); nAbs++; }
        tau += dtau;
        return true;
    }
    __device__ __forceinline__ void finish()
    {
        if (!walked) return;
        if (SINGLE) absorbPending();
        const Packet* q = P.pool + slot;
        const double taupath = tau;
        double L = q->L;
        if (SINGLE) { const double albedo = kext0 > 0 ? __ldg(P.med.ksca + ell) / kext0 : 0.0; L = L * albedo * (-expm1(-taupath)); }
        else L = Lsca;
        // ---- termination test, :289 ----
        const double Lthreshold = __ldg(P.Ltot + ell) / P.Lscale / P.minWeightReduction;
        survive = !(L <= 0 || (L <= Lthreshold && q->nscatt >= P.minfs));
        if (survive)
        {
            // ---- simulatepropagation, :519-533: sample the interaction optical depth, weight for the bias ----
            Philox rng; rng.init(P.seed, q->id, P.rngKind); rng.c2 = q->rngCtr;
            double t = 0;
            if (taupath != 0.0)
            {
                if (P.xi == 0.0) t = exponCutoff(rng, taupath);
                else
                {
                    double X = rng.uniform();
                    t = (X < P.xi) ? rng.uniform() * taupath : exponCutoff(rng, taupath);
                    double p = -exp(-t) / expm1(-taupath);
                    double qq = (1.0 - P.xi) * p + P.xi / taupath;
                    L = L * (p / qq);
                }
            }
            target = t;
            rngOut = rng.c2;
        }
        Lout = L;
    }
    __device__ __forceinline__ void collective(bool fin)
    {
        // survivors move on to the next pool, compact; packets that ended simply are not copied
        const int pos = warpAppendPosition(fin && survive, counts);
        if (pos >= 0)
        {
            Packet pk = loadPacket(P.pool + slot);
            pk.L = Lout; pk.target = target; pk.rngCtr = rngOut; pk.fresh = 0;
            storePacket(P.poolNext + pos, pk);
            if constexpr (POL) P.polNext[pos] = P.pol[slot];
        }
    }
    __device__ __forceinline__ void periodic() {}
};

template<int KIND, bool SINGLE, bool STORE, bool POL>
__global__ void __launch_bounds__(128, KIND == GRID_CART ? SKG_ABSORB_MINBLOCKS : SKG_OTHER_MINBLOCKS) absorbStage(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P, Counters* ctr, bool cartSmem,
                                                   int nAlive, int* __restrict__ counts, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    AbsorbJob<KIND, SINGLE, STORE, POL> job(G, cart, P, counts);
    runJobs<KIND>(G, cart, ctr, job, nAlive, work, P.refill);
    flushStats(ctr, job.nSeg, job.nPaths, job.nScatt, 0, job.nAbs, 0);
}

// re-walk to the sampled interaction optical depth and move the packet there:
// DustGridPath::pathlength (DustGridPath.cpp:162-173) evaluated on the fly + PhotonPackage::propagate (PhotonPackage.cpp:93-96)
template<int KIND, bool SINGLE> struct PropagateJob
{
    static constexpr bool kCartFast = SKG_MC_FAST; static constexpr bool kCartRegBorders = SKG_CART_REGBORDERS, kCartTinySelect = false; static constexpr bool kTreeHints = SKG_TREE_HINTS_MC, kCartRhoAhead = SKG_MC_RHO_AHEAD; static constexpr int kBatches = SKG_MC_BATCHES;
    const GridSetMC& G; const McDev& P;
    double rx, ry, rz, dx, dy, dz;
    double target, sPrev, tauPrev, result; bool found; int slot, ell, hint;
    int foundM;                                 // the cell in which the interaction takes place (-1: the path ended first)
    double kext0, pendDs, pendRho; int pendM; static constexpr bool single = SINGLE;      // one-component media: gather now, test one crossing later
    unsigned nSeg = 0, nPaths = 0;
    __device__ PropagateJob(const GridSetMC& G_, const McDev& P_) : G(G_), P(P_) {}
    __device__ __forceinline__ int begin(int item)
    {
        Packet* q = P.poolNext;      // the survivors the absorb stage just compacted
        slot = item;
        const Packet pk = loadPacket(q + slot);
        target = pk.target;
        if (!(target > 0)) return 0;
        ell = pk.ell; hint = KIND == GRID_CART ? -1 : pk.hint;
        rx = pk.x; ry = pk.y; rz = pk.z; dx = pk.kx; dy = pk.ky; dz = pk.kz;
        sPrev = 0; tauPrev = 0; result = 0; found = false; foundM = -1;
        kext0 = single ? __ldg(P.med.kext + ell) : 0.0; pendM = -1; pendDs = 0; pendRho = 0;
        nPaths++;
        return 1;
    }
    __device__ __forceinline__ int cellHint() const { return hint; }
    __device__ __forceinline__ void noteStart(int) {}        // the packet is about to move: finish() records where it ends up
    __device__ __forceinline__ bool outside(double ds) { nSeg++; sPrev += ds; return true; }       // only before the first cell
    __device__ __forceinline__ bool segment(int m, double ds)
    {
        nSeg++;
        if (single)
        {
            const bool cont = pendM >= 0 ? test((kext0 * pendRho) * pendDs, pendDs, pendM) : true;
            pendRho = __ldg(P.med.rho + m); pendM = m; pendDs = ds;
            return cont;
        }
        return test(KappaRho{P.med.rho, P.med.kext + ell, P.med.Ncomp, P.med.Nlambda}(m) * ds, ds, m);
    }
    template<int U> __device__ __forceinline__ bool segmentU(int m, double ds) { return segment(m, ds); }
    template<int U> __device__ __forceinline__ void idleU() {}
    __device__ __forceinline__ bool test(double dtau, double ds, int m)
    {
        double sNew = sPrev + ds;
        double tauNew = tauPrev + dtau;
        if (target < tauNew)
        {
            result = sPrev + ((target - tauPrev) / (tauNew - tauPrev)) * (sNew - sPrev);     // NR::interpolate_linlin
            found = true; foundM = m;
            return false;
        }
        sPrev = sNew; tauPrev = tauNew;
        return true;
    }
    __device__ __forceinline__ void finish()
    {
        Packet* q = P.poolNext;      // the survivors the absorb stage just compacted
        if (single && pendM >= 0 && !found) test((kext0 * pendRho) * pendDs, pendDs, pendM);
        const double s = found ? result : sPrev;
        // PhotonPackage::propagate: r += s k, with r and k from the record (not kept in registers over the walk)
        Packet* w = q + slot;
        const double px = w->x, py = w->y, pz = w->z, kx = w->kx, ky = w->ky, kz = w->kz;
        w->x = px + s * kx; w->y = py + s * ky; w->z = pz + s * kz;
        // where the packet now lies, for the traversals that start from there (peel-off, escape + absorption, propagation):
        // the leaf node of the interaction cell on tree / adaptive-mesh grids, the cell itself on Voronoi grids
        if (KIND != GRID_CART)
        {
            int loc = -1;
            if (foundM >= 0) loc = KIND == GRID_TREE ? __ldg(G.tree.cellNode + foundM) : (KIND == GRID_AMESH ? __ldg(G.amesh.cellNode + foundM) : foundM);
            w->hint = loc;
        }
    }
    __device__ __forceinline__ void collective(bool) {}
    __device__ __forceinline__ void periodic() {}
};

template<int KIND, bool SINGLE>
__global__ void __launch_bounds__(128, KIND == GRID_CART ? SKG_PROP_MINBLOCKS : SKG_OTHER_MINBLOCKS) propagateStage(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P, Counters* ctr, bool cartSmem,
                                                      int nSurv, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    PropagateJob<KIND, SINGLE> job(G, P);
    runJobs<KIND>(G, cart, ctr, job, nSurv, work, P.propRefill);
    flushStats(ctr, job.nSeg, job.nPaths, 0, 0, 0, 0);
    flushStageSegments(&ctr->propSegments, job.nSeg);
}

// Labs is wavelength-major on the device; the host interface is DustSystem's (m, ell) row-major table
__global__ void transposeLabs(const double* __restrict__ src, double* __restrict__ dst, int Ncells, int Nlambda)
{
    __shared__ double tile[32][33];
    int m0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
    { int l = l0 + r, m = m0 + threadIdx.x; if (l < Nlambda && m < Ncells) tile[r][threadIdx.x] = src[(size_t)l * Ncells + m]; }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
    { int m = m0 + r, l = l0 + threadIdx.x; if (l < Nlambda && m < Ncells) dst[(size_t)m * Nlambda + l] = tile[threadIdx.x][r]; }
}

// ---- host side -------------------------------------------------------------------------------------------

// validates one geometry description and moves its tables to the device
static SourceDev makeSourceDev(Engine& e, const skg_source& c, std::vector<DevBuf*>& bufs, bool needProfile)
{
    SourceDev s{};
    s.geometry = c.geometry;
    for (int j = 0; j < 8; j++) s.p[j] = c.p[j];
    if (c.geometry == SKG_GEOM_EXPDISK)
    {
        if (!(c.p[0] > 0) || !(c.p[1] > 0)) throw Error("The radial scale length hR and axial scale height hz should be positive");  // ExpDiskGeometry.cpp:27-28
        // ExpDiskGeometry::setupSelfBefore, ExpDiskGeometry.cpp:30-42
        const double hR = c.p[0], hz = c.p[1], Rmax = c.p[2], zmax = c.p[3], Rmin = c.p[4];
        double intz = zmax > 0 ? -2.0 * hz * std::expm1(-zmax / hz) : 2.0 * hz;
        double tmin = Rmin > 0 ? std::exp(-Rmin / hR) * (1.0 + Rmin / hR) : 1.0;
        double tmax = Rmax > 0 ? std::exp(-Rmax / hR) * (1.0 + Rmax / hR) : 0.0;
        s.rho0 = 1.0 / (hR * hR * (tmin - tmax) * 2.0 * M_PI * intz);
    }
    else if (c.geometry == SKG_GEOM_SERSIC)
    {
        if (!(c.p[0] > 0)) throw Error("the effective radius should be positive");
        if (c.ntab < 2 || !c.rv || !c.Xv) throw Error("Sersic geometry needs the tabulated inverse mass function");
        if (needProfile && !c.Sv) throw Error("Sersic geometry needs the tabulated profile S(s) for density sampling");
        if (c.p[1] == 0) s.p[1] = 1.0;
        DevBuf* a = e.takeBuf(sizeof(double) * c.ntab); DevBuf* b = e.takeBuf(sizeof(double) * c.ntab); bufs.push_back(a); bufs.push_back(b);
        a->upload(c.rv, sizeof(double) * c.ntab, e.stream); b->upload(c.Xv, sizeof(double) * c.ntab, e.stream);
        s.ntab = c.ntab; s.rv = a->as<double>(); s.Xv = b->as<double>();
        if (c.Sv) { DevBuf* d = e.takeBuf(sizeof(double) * c.ntab); bufs.push_back(d); d->upload(c.Sv, sizeof(double) * c.ntab, e.stream); s.Sv = d->as<double>(); }
        s.rho0 = 1.0 / (c.p[0] * c.p[0] * c.p[0]);      // SersicGeometry.cpp:44
    }
    else throw Error("unsupported source geometry (no CPU fallback): " + std::to_string(c.geometry));
    s.spiral_arms = c.spiral_arms; s.spiral_index = c.spiral_index; s.spiral_pitch = c.spiral_pitch;
    s.spiral_radius = c.spiral_radius; s.spiral_phase = c.spiral_phase; s.spiral_weight = c.spiral_weight;
    if (c.spiral_arms > 0)
    {
        // SpiralStructureGeometryDecorator::setupSelfBefore, SpiralStructureGeometryDecorator.cpp:24-40
        if (c.spiral_pitch <= 0 || c.spiral_pitch >= M_PI / 2.) throw Error("The pitch angle should be between 0 and 90 degrees");
        if (c.spiral_radius <= 0) throw Error("The radius zero-point should be positive");
        if (c.spiral_weight <= 0 || c.spiral_weight > 1.) throw Error("The weight of the spiral perturbation should be between 0 and 1");
        if (c.spiral_index < 0 || c.spiral_index > 10) throw Error("The arm-interarm size ratio index should be between 0 and 10");
        s.spiral_tanp = std::tan(c.spiral_pitch);
        s.spiral_cn = std::sqrt(M_PI) * std::tgamma(c.spiral_index + 1.0) / std::tgamma(c.spiral_index + 0.5);
        s.spiral_c = 1.0 + (s.spiral_cn - 1.0) * c.spiral_weight;
    }
    return s;
}

void mcSetSources(Engine& e, int Ncomp, const skg_source* comps, int Nlambda, const double* L, double emissionBias)
{
    if (Ncomp < 1 || !comps || Nlambda < 1 || !L) throw Error("skg_sources: bad arguments");
    e.recycle(e.sourceBufs);                       // (the wavelength count is checked against the medium when a phase starts)
    e.sources.clear();
    for (int h = 0; h < Ncomp; h++) e.sources.push_back(makeSourceDev(e, comps[h], e.sourceBufs, false));
    e.sourcesDev.upload(e.sources.data(), sizeof(SourceDev) * Ncomp, e.stream);
    // StellarSystem::setupSelfAfter, StellarSystem.cpp:35-52: total luminosities and per-wavelength CDFs
    e.lumHost.assign(L, L + (size_t)Ncomp * Nlambda);
    e.lumTotHost.assign(Nlambda, 0.0);
    std::vector<double> cdf((size_t)Nlambda * (Ncomp + 1), 0.0);
    for (int ell = 0; ell < Nlambda; ell++)
    {
        for (int h = 0; h < Ncomp; h++) e.lumTotHost[ell] += L[(size_t)h * Nlambda + ell];
        double* X = cdf.data() + (size_t)ell * (Ncomp + 1);
        for (int h = 0; h < Ncomp; h++) X[h + 1] = X[h] + L[(size_t)h * Nlambda + ell];     // NR::cdf, NR.hpp:404-409
        double norm = X[Ncomp];
        for (int h = 0; h <= Ncomp; h++) X[h] /= norm;
    }
    e.lumDev.upload(e.lumHost.data(), sizeof(double) * e.lumHost.size(), e.stream);
    e.lumTotDev.upload(e.lumTotHost.data(), sizeof(double) * Nlambda, e.stream);
    e.lumCdfDev.upload(cdf.data(), sizeof(double) * cdf.size(), e.stream);
    e.Nsources = Ncomp; e.NlambdaSrc = Nlambda; e.emissionBias = emissionBias;
    e.sync();
}

// DustMix::addpolarization + the derived tables of DustMix::setupSelfAfter (DustMix.cpp:96-123)
void mcSetPolarization(Engine& e, int Ntheta, const double* S11, const double* S12, const double* S33, const double* S34)
{
    if (!e.med.rho) throw Error("skg_medium_polarization needs skg_medium first");
    if (Ntheta < 2 || !S11 || !S12 || !S33 || !S34) throw Error("skg_medium_polarization: bad arguments");
    const int C = e.med.Ncomp, Nl = e.med.Nlambda; const size_t n = (size_t)C * Nl * Ntheta;
    const double* src[4] = {S11, S12, S33, S34};
    for (int q = 0; q < 4; q++) e.mueller[q].upload(src[q], sizeof(double) * n, e.stream);
    std::vector<double> thetaX(n, 0.0), pfnorm((size_t)C * Nl, 0.0);
    const double dt = M_PI / (Ntheta - 1);
    for (int h = 0; h < C; h++) for (int ell = 0; ell < Nl; ell++)
    {
        const double* s11 = S11 + ((size_t)h * Nl + ell) * Ntheta; double* X = thetaX.data() + ((size_t)h * Nl + ell) * Ntheta;
        // NR::cdf(_thetaXvv[ell], _Ntheta-1, t -> S11(ell,t+1) sin(theta_{t+1}) dt), NR.hpp:388-394
        X[0] = 0; for (int t = 0; t < Ntheta - 1; t++) X[t + 1] = X[t] + s11[t + 1] * std::sin((t + 1) * dt) * dt;
        const double norm = X[Ntheta - 1];
        if (!(norm > 0)) throw Error("the Mueller coefficient S11 must be positive");
        for (int t = 0; t < Ntheta; t++) X[t] /= norm;
        double sum = 0; for (int t = 0; t < Ntheta; t++) sum += s11[t] * std::sin(t * dt) * dt;
        pfnorm[(size_t)h * Nl + ell] = 2.0 / sum;
    }
    e.thetaX.upload(thetaX.data(), sizeof(double) * n, e.stream); e.pfnorm.upload(pfnorm.data(), sizeof(double) * pfnorm.size(), e.stream);
    e.med.S11 = e.mueller[0].as<double>(); e.med.S12 = e.mueller[1].as<double>(); e.med.S33 = e.mueller[2].as<double>(); e.med.S34 = e.mueller[3].as<double>();
    e.med.thetaX = e.thetaX.as<double>(); e.med.pfnorm = e.pfnorm.as<double>(); e.med.Ntheta = Ntheta;
    e.sync();
}

void mcSetInstruments(Engine& e, int n, const skg_instrument* instr)
{
    if (n < 0 || (n > 0 && !instr)) throw Error("skg_instruments: bad arguments");
    if (!e.med.Nlambda) throw Error("skg_instruments needs skg_medium first (number of wavelengths)");
    e.recycle(e.instrBufs); e.instr.clear();
    std::vector<PerspDev> persp;
    for (int i = 0; i < n; i++)
    {
        const skg_instrument& s = instr[i];
        InstrDev d{};
        d.kind = s.kind;
        if (s.kind < SKG_INSTR_FRAME || s.kind > SKG_INSTR_PERSPECTIVE) throw Error("unsupported instrument kind");
        if (s.kind == SKG_INSTR_PERSPECTIVE)
        {
            // PerspectiveInstrument::setupSelfBefore, PerspectiveInstrument.cpp:36-108
            if (s.Nxp <= 0 || s.Nyp <= 0) throw Error("Number of pixels was not set");
            if (s.fovxp <= 0) throw Error("Viewport width was not set");
            if (s.upX == 0 && s.upY == 0 && s.upZ == 0) throw Error("Upwards direction was not set");
            if (s.focal <= 0) throw Error("Focal length was not set");
            const double Vx = s.viewX, Vy = s.viewY, Vz = s.viewZ, Cx = s.crossX, Cy = s.crossY, Cz = s.crossZ, Ux = s.upX, Uy = s.upY, Uz = s.upZ, Fe = s.focal;
            const double Gn = std::sqrt((Vx - Cx) * (Vx - Cx) + (Vy - Cy) * (Vy - Cy) + (Vz - Cz) * (Vz - Cz));
            if (Gn < 1e-20) throw Error("Crosshair is too close to viewport origin");
            const double a = (Vx - Cx) / Gn, b = (Vy - Cy) / Gn, c = (Vz - Cz) / Gn;
            PerspDev V{}; V.Nx = s.Nxp; V.Ny = s.Nyp; V.s = s.fovxp / s.Nxp;
            V.Ex = Vx + Fe * a; V.Ey = Vy + Fe * b; V.Ez = Vz + Fe * c;
            // HomogeneousTransform (HomogeneousTransform.cpp): row-vector convention, every step post-multiplied
            struct H
            {
                double M[4][4];
                H() { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i][j] = i == j ? 1. : 0.; }
                void cat(const H& t)
                { H c0(*this); for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i][j] = c0.M[i][0] * t.M[0][j] + c0.M[i][1] * t.M[1][j] + c0.M[i][2] * t.M[2][j] + c0.M[i][3] * t.M[3][j]; }
                void translate(double x, double y, double z) { H t; t.M[3][0] = x; t.M[3][1] = y; t.M[3][2] = z; cat(t); }
                void scale(double x, double y, double z) { H t; t.M[0][0] = x; t.M[1][1] = y; t.M[2][2] = z; cat(t); }
                void rotateX(double co, double si) { H t; t.M[1][1] = co; t.M[2][2] = co; t.M[1][2] = -si; t.M[2][1] = si; cat(t); }
                void rotateY(double co, double si) { H t; t.M[0][0] = co; t.M[2][2] = co; t.M[0][2] = -si; t.M[2][0] = si; cat(t); }
                void rotateZ(double co, double si) { H t; t.M[0][0] = co; t.M[1][1] = co; t.M[0][1] = -si; t.M[1][0] = si; cat(t); }
                void perspectiveZ(double f) { H t; t.M[2][3] = 1. / f; t.M[3][2] = -f; t.M[3][3] = 0.; cat(t); }
            } T;
            T.translate(-V.Ex, -V.Ey, -V.Ez);
            double v = std::sqrt(b * b + c * c);
            if (v > 0.3)
            {
                T.rotateX(c / v, -b / v); T.rotateY(v, -a);
                const double k = (b * b + c * c) * Ux - a * b * Uy - a * c * Uz, l = c * Uy - b * Uz, u = std::sqrt(k * k + l * l);
                T.rotateZ(l / u, -k / u);
            }
            else
            {
                v = std::sqrt(a * a + c * c);
                T.rotateY(c / v, -a / v); T.rotateX(v, -b);
                const double k = c * Ux - a * Uz, l = (a * a + c * c) * Uy - a * b * Ux - b * c * Uz, u = std::sqrt(k * k + l * l);
                T.rotateZ(l / u, -k / u);
            }
            T.scale(1., 1., -1.);
            T.perspectiveZ(Fe);
            T.scale(1. / V.s, 1. / V.s, 1.);
            T.translate(s.Nxp / 2., s.Nyp / 2., 0);
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) V.M[i][j] = T.M[i][j];
            d.Nxp = s.Nxp; d.Nyp = s.Nyp; d.frames = nullptr; d.mfPixels = 0; d.mfTotal = d.mfComp0 = -1; d.mfNcomp = 0;
            d.frameCount = (long long)s.Nxp * s.Nyp * e.med.Nlambda;
            DevBuf* f = e.takeBuf(sizeof(double) * (size_t)d.frameCount); e.instrBufs.push_back(f);
            f->ensure(sizeof(double) * (size_t)d.frameCount); SKG_CUDA(cudaMemsetAsync(f->p, 0, sizeof(double) * (size_t)d.frameCount, e.stream));
            d.frame = f->as<double>(); V.frame = d.frame;
            persp.push_back(V);
            e.instr.push_back(d);
            continue;
        }
        if (s.kind == SKG_INSTR_FULL && (s.scatteringLevels < 0 || s.scatteringLevels > 1000)) throw Error("invalid number of scattering levels");
        if (s.distance <= 0) throw Error("Distance was not set");                    // DistantInstrument.cpp:32
        // DistantInstrument::setupSelfBefore, DistantInstrument.cpp:27-50
        d.costheta = std::cos(s.inclination); d.sintheta = std::sin(s.inclination);
        d.cosphi = std::cos(s.azimuth); d.sinphi = std::sin(s.azimuth);
        d.cospa = std::cos(s.positionAngle); d.sinpa = std::sin(s.positionAngle);
        {
            const double eps = 1e-8; double theta = s.inclination, phi = s.azimuth;      // Direction(theta,phi), Direction.cpp:12-38
            if (theta < -eps || theta > M_PI + eps) throw Error("Theta should be between 0 and pi.");
            if (theta <= eps) { d.kobsx = 0; d.kobsy = 0; d.kobsz = 1; }
            else if (theta >= M_PI - eps) { d.kobsx = 0; d.kobsy = 0; d.kobsz = -1; }
            else { double st = std::sin(theta); d.kobsx = st * std::cos(phi); d.kobsy = st * std::sin(phi); d.kobsz = std::cos(theta); }
        }
        // DistantInstrument::bfky, DistantInstrument.cpp:47-49
        d.kyx = -d.cosphi * d.costheta * d.cospa - d.sinphi * d.sinpa; d.kyy = -d.sinphi * d.costheta * d.cospa + d.cosphi * d.sinpa; d.kyz = d.sintheta * d.cospa;
        d.frameCount = 0; d.frames = nullptr; d.mfPixels = 0; d.mfTotal = d.mfComp0 = -1; d.mfNcomp = 0;
        if (s.kind == SKG_INSTR_MULTIFRAME)
        {
            // MultiFrameInstrument::setupSelfBefore (MultiFrameInstrument.cpp:22-29) + InstrumentFrame::setupSelfBefore (InstrumentFrame.cpp:29-44, :62-71):
            // one frame per wavelength; the arrays of all frames share one allocation, slab by slab
            if (!s.frames) throw Error("Number of instrument frames must equal number of wavelengths");
            std::vector<FrameDev> fr(e.med.Nlambda);
            long long off = 0;
            for (int ell = 0; ell < e.med.Nlambda; ell++)
            {
                const skg_instrument_frame& f = s.frames[ell];
                if (f.Nxp <= 0 || f.Nyp <= 0) throw Error("Number of pixels was not set");
                if (f.fovxp <= 0 || f.fovyp <= 0) throw Error("Field of view was not set");
                fr[ell] = FrameDev{f.Nxp, f.Nyp, f.xpc - 0.5 * f.fovxp, f.ypc - 0.5 * f.fovyp, f.fovxp / f.Nxp, f.fovyp / f.Nyp, off};
                off += (long long)f.Nxp * f.Nyp;
            }
            int slabs = 0;
            d.mfPixels = off; d.mfNcomp = (int)e.sources.size();
            if (s.writeTotal) d.mfTotal = slabs++;
            if (s.writeStellarComps)
            {
                if (d.mfNcomp < 1) throw Error("a MultiFrameInstrument that records the stellar components needs skg_sources first");
                d.mfComp0 = slabs; slabs += d.mfNcomp;
            }
            DevBuf* fb = e.takeBuf(sizeof(FrameDev) * fr.size()); e.instrBufs.push_back(fb);
            fb->upload(fr.data(), sizeof(FrameDev) * fr.size(), e.stream); SKG_CUDA(cudaStreamSynchronize(e.stream));      // (fr is a local)
            d.frames = fb->as<FrameDev>();
            d.frameCount = (long long)slabs * off;
            if (d.frameCount > 0)
            {
                DevBuf* f = e.takeBuf(sizeof(double) * (size_t)d.frameCount); e.instrBufs.push_back(f);
                f->ensure(sizeof(double) * (size_t)d.frameCount); SKG_CUDA(cudaMemsetAsync(f->p, 0, sizeof(double) * (size_t)d.frameCount, e.stream));
                d.frame = f->as<double>();
            }
            e.instr.push_back(d);
            continue;
        }
        if (s.kind != SKG_INSTR_SED)
        {
            // SingleFrameInstrument::setupSelfBefore, SingleFrameInstrument.cpp:26-42
            if (s.Nxp <= 0 || s.Nyp <= 0) throw Error("Number of pixels was not set");
            if (s.fovxp <= 0 || s.fovyp <= 0) throw Error("Field of view was not set");
            d.Nxp = s.Nxp; d.Nyp = s.Nyp;
            d.xpmin = s.xpc - 0.5 * s.fovxp; d.xpsiz = s.fovxp / s.Nxp;
            d.ypmin = s.ypc - 0.5 * s.fovyp; d.ypsiz = s.fovyp / s.Nyp;
            if (s.kind == SKG_INSTR_FULL) { d.Nscatt = s.scatteringLevels; d.pol = e.med.Ntheta > 0 ? 1 : 0; d.Nchan = 5 + d.Nscatt + (d.pol ? 3 : 0); }
            size_t bytes = sizeof(double) * (size_t)s.Nxp * s.Nyp * e.med.Nlambda * (s.kind == SKG_INSTR_FULL ? d.Nchan : 1);
            DevBuf* f = e.takeBuf(bytes); e.instrBufs.push_back(f);
            f->ensure(bytes); SKG_CUDA(cudaMemsetAsync(f->p, 0, bytes, e.stream));
            if (s.kind == SKG_INSTR_FULL) d.chanFrame = f->as<double>(); else { d.frame = f->as<double>(); d.frameCount = (long long)s.Nxp * s.Nyp * e.med.Nlambda; }
        }
        if (s.kind != SKG_INSTR_FRAME)
        {
            size_t bytes = sizeof(double) * e.med.Nlambda * (s.kind == SKG_INSTR_FULL ? d.Nchan : 1);
            DevBuf* f = e.takeBuf(bytes); e.instrBufs.push_back(f);
            f->ensure(bytes); SKG_CUDA(cudaMemsetAsync(f->p, 0, bytes, e.stream));
            if (s.kind == SKG_INSTR_FULL) d.chanSed = f->as<double>(); else d.sed = f->as<double>();
        }
        e.instr.push_back(d);
    }
    if (n > 0) e.instrDev.upload(e.instr.data(), sizeof(InstrDev) * (size_t)n, e.stream);
    // observer groups: instruments with the same line of sight share one peel-off traversal per event
    std::vector<InstrDev> grouped; std::vector<ObsGroup> groups; std::vector<char> used(n, 0);
    e.Npersp = (int)persp.size();
    if (!persp.empty()) { e.perspDev.upload(persp.data(), sizeof(PerspDev) * persp.size(), e.stream); SKG_CUDA(cudaStreamSynchronize(e.stream)); }     // (persp is a local)
    for (int i = 0; i < n; i++)
    {
        if (e.instr[i].kind == SKG_INSTR_PERSPECTIVE) used[i] = 1;       // not a line of sight: one ray per packet towards the eye
        if (used[i]) continue;
        ObsGroup g; g.kx = e.instr[i].kobsx; g.ky = e.instr[i].kobsy; g.kz = e.instr[i].kobsz; g.first = (int)grouped.size(); g.count = 0;
        for (int j = i; j < n; j++)
            if (!used[j] && e.instr[j].kobsx == g.kx && e.instr[j].kobsy == g.ky && e.instr[j].kobsz == g.kz)
            { used[j] = 1; grouped.push_back(e.instr[j]); g.count++; }
        groups.push_back(g);
    }
    e.instrPol = e.med.Ntheta > 0;
    e.maxGroupCount = 0; for (const ObsGroup& g : groups) e.maxGroupCount = std::max(e.maxGroupCount, g.count);
    e.Ngroups = (int)groups.size(); e.instrNlambda = e.med.Nlambda; e.accInstr = Engine::ACC_ZERO;
    if (!grouped.empty()) e.instrGroupedDev.upload(grouped.data(), sizeof(InstrDev) * grouped.size(), e.stream);
    if (!groups.empty()) e.groupsDev.upload(groups.data(), sizeof(ObsGroup) * groups.size(), e.stream);
    e.sync();
}

// every accumulator back to zero: detector arrays, the stellar AND the dust absorption table (the reference starts a
// simulation with _Labsstelvv = _Labsdustvv = 0, PanDustSystem.cpp:100-118)
void mcResetResults(Engine& e)
{
    const size_t Nl = (size_t)e.instrNlambda;
    for (const InstrDev& d : e.instr)
    {
        if (d.frame) SKG_CUDA(cudaMemsetAsync(d.frame, 0, sizeof(double) * (size_t)d.frameCount, e.stream));
        if (d.sed) SKG_CUDA(cudaMemsetAsync(d.sed, 0, sizeof(double) * Nl, e.stream));
        if (d.chanFrame) SKG_CUDA(cudaMemsetAsync(d.chanFrame, 0, sizeof(double) * (size_t)d.Nxp * d.Nyp * Nl * d.Nchan, e.stream));
        if (d.chanSed) SKG_CUDA(cudaMemsetAsync(d.chanSed, 0, sizeof(double) * Nl * d.Nchan, e.stream));
    }
    if (e.labs.p && e.labsCount) SKG_CUDA(cudaMemsetAsync(e.labs.p, 0, sizeof(double) * e.labsCount, e.stream));
    if (e.labsDust.p && e.labsCount) SKG_CUDA(cudaMemsetAsync(e.labsDust.p, 0, sizeof(double) * e.labsCount, e.stream));
    e.accLabs = e.accLabsDust = e.accInstr = Engine::ACC_ZERO;
    e.sync();
}

// block-wise sum of an array into *out (fp64 atomicAdd of one partial sum per CTA)
__global__ void __launch_bounds__(256) sumArray(const double* __restrict__ a, size_t n, double* __restrict__ out)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += a[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; w++) t += part[w]; atomicAdd(out, t); }
}

// PanDustSystem::Labsstellartot / Labsdusttot (PanDustSystem.cpp:351-366), the rank-local part
double mcLabsTotal(Engine& e, int which)
{
    DevBuf& src = which ? e.labsDust : e.labs;
    if (!src.p || e.labsCount == 0) return 0.0;
    e.scalarDev.ensure(sizeof(double));
    SKG_CUDA(cudaMemsetAsync(e.scalarDev.p, 0, sizeof(double), e.stream));
    sumArray<<<e.smCount * 4, 256, 0, e.stream>>>(src.as<double>(), (size_t)e.labsCount, e.scalarDev.as<double>());
    e.launches++; SKG_CUDA(cudaGetLastError());
    double total = 0;
    e.readSmall(&total, e.scalarDev.p, sizeof(double));
    return total;
}

void mcTransposeLabs(Engine& e, const double* src, double* dst)
{
    if (e.Ncells <= 0 || e.labsCount % e.Ncells) throw Error("the absorption table does not belong to the current grid");
    const int Nc = e.Ncells, Nl = (int)(e.labsCount / e.Ncells);       // the table's own shape, not the sources'
    dim3 grid((Nc + 31) / 32, (Nl + 31) / 32), block(32, 8);
    transposeLabs<<<grid, block, 0, e.stream>>>(src, dst, Nc, Nl);
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void mcFetchLabs(Engine& e, double* host, int add, int which)
{
    DevBuf& src = which ? e.labsDust : e.labs;
    if (!src.p || e.labsCount == 0) throw Error(which ? "absorption of dust emission was not stored" : "absorption rates were not stored");
    // (the table's own shape: sources left over from an earlier run may have another number of wavelengths)
    if (e.Ncells <= 0 || e.labsCount % e.Ncells) throw Error("the absorption table does not belong to the current grid");
    int Nc = e.Ncells, Nl = (int)(e.labsCount / e.Ncells);
    e.labsT.ensure(sizeof(double) * e.labsCount);
    dim3 grid((Nc + 31) / 32, (Nl + 31) / 32), block(32, 8);
    transposeLabs<<<grid, block, 0, e.stream>>>(src.as<double>(), e.labsT.as<double>(), Nc, Nl);
    e.launches++; SKG_CUDA(cudaGetLastError());
    if (!add) { SKG_CUDA(cudaMemcpyAsync(host, e.labsT.p, sizeof(double) * e.labsCount, cudaMemcpyDeviceToHost, e.stream)); e.sync(); return; }
    std::vector<double> tmp(e.labsCount);
    SKG_CUDA(cudaMemcpyAsync(tmp.data(), e.labsT.p, sizeof(double) * e.labsCount, cudaMemcpyDeviceToHost, e.stream)); e.sync();
    for (int64_t i = 0; i < e.labsCount; i++) host[i] += tmp[i];
}

void mcLabsBolometric(Engine& e, double* host)
{
    if ((!e.labs.p && !e.labsDust.p) || e.labsCount == 0) throw Error("absorption rates were not stored");
    int Nl = (int)(e.labsCount / e.Ncells);
    e.scratchTau.ensure(sizeof(double) * e.Ncells);
    sumOverWavelengths<<<(e.Ncells + 127) / 128, 128, 0, e.stream>>>(e.labs.as<double>(), e.labsDust.as<double>(), e.scratchTau.as<double>(), e.Ncells, Nl);
    e.launches++; SKG_CUDA(cudaGetLastError());
    SKG_CUDA(cudaMemcpyAsync(host, e.scratchTau.p, sizeof(double) * e.Ncells, cudaMemcpyDeviceToHost, e.stream)); e.sync();
}

__global__ void publishCounts(const int* __restrict__ counts, int* hostCounts)
{
    if (threadIdx.x < 2) { reinterpret_cast<volatile int*>(hostCounts)[threadIdx.x] = counts[threadIdx.x]; __threadfence_system(); }
}

template<int KIND>
static void shootWavefront(Engine& e, const GridSetMC& G, McDev& P, unsigned long long total, int pool, size_t smem, bool cartSmem)
{
    Packet* poolA = e.mcPool.as<Packet>(); Packet* poolB = poolA + pool;
    PolState* polA = P.med.Ntheta > 0 ? e.mcPolPool.as<PolState>() : nullptr; PolState* polB = polA ? polA + pool : nullptr;
    const bool single = P.med.Ncomp == 1, store = P.labs != nullptr, pol = P.med.Ntheta > 0;
    int* counts = e.mcCounts.as<int>();      // [0] survivors, [2..4] work counters of the three traversal stages
    auto blocksFor = [&](long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 127) / 128, (long long)e.smCount * 16)); };
    for (cudaEvent_t& ev : e.mcEvents) if (!ev) SKG_CUDA(cudaEventCreate(&ev));
    cudaEvent_t* ev = e.mcEvents;            // 0..3: boundaries launch | peel | absorb | end; 4..5: around propagate
    for (double& t : e.stageMs) t = 0;
    e.mcIterations = 0;
    unsigned long long launched = 0;
    int nAlive = 0;                          // packets in flight: poolA[0, nAlive), survivors first, then the newly launched
    int* hostCounts = e.mcHostCounts;
    bool propagatePending = false;
    auto addMs = [&](int stage, cudaEvent_t a, cudaEvent_t b) { float ms = 0; SKG_CUDA(cudaEventElapsedTime(&ms, a, b)); e.stageMs[stage] += ms; };
    while (true)
    {
        P.pool = poolA; P.poolNext = poolB; P.pol = polA; P.polNext = polB;
        int nLaunch = (int)std::min<unsigned long long>((unsigned long long)(pool - nAlive), total - launched);
        SKG_CUDA(cudaEventRecord(ev[0], e.stream));
        if (nLaunch > 0)
        {
            if (P.phase == SKG_PHASE_STELLAR) launchStage<<<blocksFor(nLaunch), 128, 0, e.stream>>>(P, e.ctr(), nLaunch, launched, nAlive);
            else launchDustStage<KIND><<<blocksFor(nLaunch), 128, 0, e.stream>>>(G, P, e.ctr(), nLaunch, launched, nAlive);
            e.launches++;
            nAlive += nLaunch; launched += nLaunch;
        }
        if (nAlive == 0) break;
        SKG_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int), e.stream));
        SKG_CUDA(cudaEventRecord(ev[1], e.stream));
        if (P.Ngroups > 0 && P.phase != SKG_PHASE_DUST_SELFABS)
        {
            const int nb = blocksFor((long long)nAlive * P.Ngroups);
            if (pol) { if (single) peelStage<KIND, true, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 2);
                       else peelStage<KIND, false, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 2); }
            else if (single) peelStage<KIND, true, false><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 2);
            else peelStage<KIND, false, false><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 2);
            e.launches++;
        }
        if (P.Npersp > 0 && P.phase != SKG_PHASE_DUST_SELFABS)
        {
            // PerspectiveInstruments: one more peel-off ray per packet each, towards the eye
            const int nb = blocksFor((long long)nAlive * P.Npersp);
            if (pol) { if (single) peelStage<KIND, true, true, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 6);
                       else peelStage<KIND, false, true, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 6); }
            else if (single) peelStage<KIND, true, false, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 6);
            else peelStage<KIND, false, false, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 6);
            e.launches++;
        }
        SKG_CUDA(cudaEventRecord(ev[2], e.stream));
        {
            const int nb = blocksFor(nAlive);
#define SKG_ABSORB(S1, S2, S3) absorbStage<KIND, S1, S2, S3><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts, counts + 3)
            if (pol) { if (single && store) SKG_ABSORB(true, true, true); else if (single) SKG_ABSORB(true, false, true);
                       else if (store) SKG_ABSORB(false, true, true); else SKG_ABSORB(false, false, true); }
            else if (single && store) SKG_ABSORB(true, true, false);
            else if (single) SKG_ABSORB(true, false, false);
            else if (store) SKG_ABSORB(false, true, false);
            else SKG_ABSORB(false, false, false);
#undef SKG_ABSORB
            e.launches++;
        }
        if (P.contScatt && P.Ngroups > 0 && P.phase != SKG_PHASE_DUST_SELFABS && P.med.rho)
        {
            // continuous scattering: every segment of every packet's path peels off.  After the absorb stage, which has scattered the
            // packets that come from an interaction (new direction, Stokes state and scattering count are back in their records of
            // the current pool, whose luminosities are still those before escape and absorption): the reference walks the NEW
            // direction, fillOpticalDepth -> continuouspeeloffscattering (MonteCarloSimulation.cpp:286-287).  Timed with the absorb stage.
            const int nb = blocksFor((long long)nAlive * P.Ngroups);
            if (pol) { if (single) contPeelStage<KIND, true, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 5);
                       else contPeelStage<KIND, false, true><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 5); }
            else if (single) contPeelStage<KIND, true, false><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 5);
            else contPeelStage<KIND, false, false><<<nb, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nAlive, counts + 5);
            e.launches++;
        }
        SKG_CUDA(cudaEventRecord(ev[3], e.stream));
        // the survivor count travels as a store into mapped page-locked memory, not as a copy: a copy would queue on the
        // device-to-host copy engine behind a result transfer of skg_fetch_snapshot_async that is still in flight
        publishCounts<<<1, 32, 0, e.stream>>>(counts, hostCounts); e.launches++;
        SKG_CUDA(cudaGetLastError());
        e.sync();
        if (propagatePending) { addMs(3, ev[4], ev[5]); propagatePending = false; }
        addMs(0, ev[0], ev[1]); addMs(1, ev[1], ev[2]); addMs(2, ev[2], ev[3]);
        e.mcIterations++;
        int nSurv = hostCounts[0];
        if (nSurv > 0)
        {
            SKG_CUDA(cudaEventRecord(ev[4], e.stream));
            if (single) propagateStage<KIND, true><<<blocksFor(nSurv), 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nSurv, counts + 4);
            else propagateStage<KIND, false><<<blocksFor(nSurv), 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem, nSurv, counts + 4);
            e.launches++;
            SKG_CUDA(cudaEventRecord(ev[5], e.stream));
            propagatePending = true;
        }
        std::swap(poolA, poolB); std::swap(polA, polB);
        nAlive = nSurv;
    }
    SKG_CUDA(cudaGetLastError());
    e.sync();
    if (propagatePending) addMs(3, ev[4], ev[5]);
}

// the per-wavelength cumulative distributions of the cell luminosities (NR::cdf) and their totals
static void mcBuildCdfs(Engine& e, int Nlambda, int Ncells)
{
    size_t tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp, e.dustLv.as<double>(), e.dustCdf.as<double>() + 1, Ncells, e.stream);
    e.scratchCub.ensure(tmp);
    for (int ell = 0; ell < Nlambda; ell++)
    {
        double* cdf = e.dustCdf.as<double>() + (size_t)ell * (Ncells + 1);
        SKG_CUDA(cub::DeviceScan::InclusiveSum(e.scratchCub.p, tmp, e.dustLv.as<double>() + (size_t)ell * Ncells, cdf + 1, Ncells, e.stream));
        captureTotal<<<1, 1, 0, e.stream>>>(cdf, e.dustLtot.as<double>() + ell, Ncells);
        normalizeCdf<<<std::min(148 * 8, (Ncells + 255) / 256), 256, 0, e.stream>>>(cdf, e.dustLtot.as<double>() + ell, Ncells);
        e.launches += 4;
    }
    SKG_CUDA(cudaGetLastError());
}

// ---- dust emission spectra (DustLib::calculate for AllCellsDustLib + GreyBodyDustEmissivity) ----------------------
struct DustLibDev
{
    const double* labsStel; const double* labsDust;     // [Nlambda*Ncells] wavelength-major, either may be null
    const double* rho; const double* vol; const double* kabs; const double* lambda; const double* dlambda;
    const double* Tv; const double* planckabs;          // [NT+1], [Ncomp*(NT+1)]
    double* out;                                        // Lcell [Nlambda*Ncells]
    int Ncells, Ncomp, Nlambda, NT;
};

// PlanckFunction::operator(), PlanckFunction.cpp:24-33 (Units.cpp constants)
__device__ __forceinline__ double planckB(double lambda, double T)
{
    const double h = 6.62606957e-34, c = 2.99792458e8, k = 1.3806488e-23;
    double x = h * c / (lambda * k * T);
    return 2.0 * h * c * c / pow(lambda, 5) / (exp(x) - 1.0);
}

// one thread per cell; the absorption tables are wavelength-major, so every pass over ell is coalesced over the cells
__global__ void __launch_bounds__(128) dustLibraryKernel(const __grid_constant__ DustLibDev D)
{
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < D.Ncells; m += gridDim.x * blockDim.x)
    {
        const size_t N = D.Ncells;
        const double fac = 4.0 * M_PI * D.vol[m];
        double T[8]; double Labsbol = 0;
        for (int h = 0; h < D.Ncomp; h++) T[h] = 0;
        // mean intensity and the Planck-weighted absorption of every component (DustMix::equilibrium)
        double pa[8];
        for (int h = 0; h < D.Ncomp; h++) pa[h] = 0;
        for (int ell = 0; ell < D.Nlambda; ell++)
        {
            double L = 0;
            if (D.labsStel) L += D.labsStel[ell * N + m];
            if (D.labsDust) L += D.labsDust[ell * N + m];
            Labsbol += L;
        }
        // PanDustSystem::Labs(m) sums the stellar table first, then the dust table; keep that order for the bolometric value
        {
            double a = 0, b = 0;
            if (D.labsStel) for (int ell = 0; ell < D.Nlambda; ell++) a += D.labsStel[ell * N + m];
            if (D.labsDust) for (int ell = 0; ell < D.Nlambda; ell++) b += D.labsDust[ell * N + m];
            Labsbol = a + b;
        }
        for (int ell = 0; ell < D.Nlambda; ell++)
        {
            double kabsrho = 0.0;
            for (int h = 0; h < D.Ncomp; h++) kabsrho += D.kabs[h * D.Nlambda + ell] * D.rho[(size_t)m * D.Ncomp + h];
            double L = 0;
            if (D.labsStel) L += D.labsStel[ell * N + m];
            if (D.labsDust) L += D.labsDust[ell * N + m];
            double J = L / (kabsrho * fac) / D.dlambda[ell];
            if (!isfinite(J)) J = 0.0;                              // DustSystem.cpp:951-952
            for (int h = 0; h < D.Ncomp; h++) pa[h] += D.kabs[h * D.Nlambda + ell] * J * D.dlambda[ell];
        }
        for (int h = 0; h < D.Ncomp; h++)
        {
            // DustMix::invplanckabs: NR::locate_clip on the table + linear interpolation
            const double* tab = D.planckabs + (size_t)h * (D.NT + 1);
            int p = locateClip(tab, pa[h], D.NT + 1);
            T[h] = D.Tv[p] + ((pa[h] - tab[p]) / (tab[p + 1] - tab[p])) * (D.Tv[p + 1] - D.Tv[p]);
        }
        // emission spectrum, converted to luminosities and normalised; written straight into the output, then rescaled
        double total = 0;
        for (int ell = 0; ell < D.Nlambda; ell++)
        {
            double ev = 0;
            for (int h = 0; h < D.Ncomp; h++)
            {
                double e1 = D.kabs[h * D.Nlambda + ell] * planckB(D.lambda[ell], T[h]);
                ev += D.Ncomp > 1 ? e1 * D.rho[(size_t)m * D.Ncomp + h] : e1;
            }
            double Lv = ev * D.dlambda[ell];
            if (!isfinite(Lv)) Lv = 0.0;
            D.out[ell * N + m] = Lv;
            total += Lv;
        }
        const bool emit = Labsbol > 0.0 && total > 0;              // PanMonteCarloSimulation.cpp:196-197
        for (int ell = 0; ell < D.Nlambda; ell++) D.out[ell * N + m] = emit ? Labsbol * (D.out[ell * N + m] / total) : 0.0;
    }
}

void mcDustLibrary(Engine& e, const double* volumes, const double* kappaabs, const double* lambda, const double* dlambda)
{
    if (!e.med.rho) throw Error("skg_dust_library needs skg_medium first");
    const int Nl = e.med.Nlambda, C = e.med.Ncomp, N = e.Ncells;
    if (C > 8) throw Error("at most 8 dust components are supported");
    e.libVol.upload(volumes, sizeof(double) * N, e.stream); e.libKabs.upload(kappaabs, sizeof(double) * C * Nl, e.stream);
    e.libLambda.upload(lambda, sizeof(double) * Nl, e.stream); e.libDlambda.upload(dlambda, sizeof(double) * Nl, e.stream);
    // temperature grid and Planck-integrated absorption coefficients, DustMix.cpp:238-262
    const int NT = 1000;
    std::vector<double> Tv(NT + 1), pab((size_t)C * (NT + 1), 0.0);
    { double q = std::pow(500., 1. / (NT - 1)), qn = std::pow(q, NT); for (int i = 0; i <= NT; ++i) Tv[i] = 0. + (1. - std::pow(q, i)) / (1. - qn) * 5000.; }
    const double hh = 6.62606957e-34, cc = 2.99792458e8, kk = 1.3806488e-23;
    for (int p = 1; p <= NT; p++) for (int h = 0; h < C; h++)
    {
        double planckabs = 0.0;
        for (int ell = 0; ell < Nl; ell++)
        {
            double x = hh * cc / (lambda[ell] * kk * Tv[p]);
            double B = 2.0 * hh * cc * cc / std::pow(lambda[ell], 5) / (std::exp(x) - 1.0);
            planckabs += kappaabs[h * Nl + ell] * B * dlambda[ell];
        }
        pab[(size_t)h * (NT + 1) + p] = planckabs;
    }
    e.libTv.upload(Tv.data(), sizeof(double) * Tv.size(), e.stream); e.libPlanckabs.upload(pab.data(), sizeof(double) * pab.size(), e.stream);
    e.haveDustLib = true;
    e.sync();
}

double* mcDustCellLuminosities(Engine& e)
{
    if (!e.haveDustLib) throw Error("skg_dust_library has not been called");
    if ((!e.labs.p && !e.labsDust.p) || e.labsCount == 0) throw Error("absorption rates were not stored");
    DustLibDev D{};
    D.labsStel = e.labs.as<double>(); D.labsDust = e.labsDust.as<double>();
    D.rho = e.med.rho; D.vol = e.libVol.as<double>(); D.kabs = e.libKabs.as<double>(); D.lambda = e.libLambda.as<double>(); D.dlambda = e.libDlambda.as<double>();
    D.Tv = e.libTv.as<double>(); D.planckabs = e.libPlanckabs.as<double>();
    D.Ncells = e.Ncells; D.Ncomp = e.med.Ncomp; D.Nlambda = e.med.Nlambda; D.NT = 1000;
    e.dustLvOut.ensure(sizeof(double) * (size_t)D.Ncells * D.Nlambda);
    D.out = e.dustLvOut.as<double>();
    dustLibraryKernel<<<std::min(e.smCount * 16, (D.Ncells + 127) / 128), 128, 0, e.stream>>>(D);
    e.launches++; SKG_CUDA(cudaGetLastError());
    e.sync();
    return D.out;
}

// common driver of the three shooting phases
static void runPhase(Engine& e, const skg_mc_params& p, int phase, double dustBias, const std::vector<double>& LtotHost, skg_mc_stats* stats)
{
    int Nlambda = e.med.Nlambda ? e.med.Nlambda : e.NlambdaSrc;
    if (p.ellBegin < 0 || p.ellEnd > Nlambda || p.ellBegin > p.ellEnd) throw Error("wavelength range out of bounds");
    if (p.scattBias < 0 || p.scattBias > 1) throw Error("scattBias should be between 0 and 1");
    if (e.med.Ncomp > 8) throw Error("at most 8 dust components are supported");
    if (e.med.rho && e.med.Ncells != e.Ncells) throw Error("the medium has " + std::to_string(e.med.Ncells) + " cells but the grid has " + std::to_string(e.Ncells) + ": call skg_medium again");
    if (!e.instr.empty() && e.instrPol != (e.med.Ntheta > 0)) throw Error("the instruments were set up before the polarisation of the medium changed: call skg_instruments again");
    if (!e.instr.empty() && e.instrNlambda != Nlambda) throw Error("the instruments were set up for " + std::to_string(e.instrNlambda) + " wavelengths but the medium has " + std::to_string(Nlambda) + ": call skg_instruments again");
    if (!(p.packages >= 0) || p.packages > 1e15) throw Error("Number of photon packages is negative or larger than implementation limit of 1e15");
    McDev P{};
    P.med = e.med; if (!e.med.rho) { P.med.Nlambda = Nlambda; P.med.Ncomp = 0; }
    P.phase = phase; P.rngKind = (unsigned)phase;
    // parked lanes at which a warp refills, measured per stage (profiles/README.md): absorb 24 / propagate 18 / peel 28 on the
    // Cartesian grid (issue-bound crossings: a refill interrupts every walking lane of the warp), 14 / 14 / 16 on the tree and
    // adaptive-mesh grids, 14 / 14 / 8 on the Voronoi grid
    P.refill = e.gridKind == GRID_CART ? 24 : 14; P.propRefill = e.gridKind == GRID_CART ? 18 : 14;
    if (const char* v = getenv("SKG_REFILL")) P.refill = P.propRefill = std::max(1, std::min(32, atoi(v)));
    // measured (profiles/README.md): on the Cartesian grid the peel-off stage is fastest when a warp waits for 28 parked lanes
    // (its crossings are issue-bound), on the tree / adaptive-mesh grids at 16 and on the Voronoi grid at 8 (long crossings:
    // more walking lanes per step cost nothing, idle ones do)
    P.peelRefill = e.gridKind == GRID_CART ? 28 : (e.gridKind == GRID_VORO ? 8 : 16);
    if (const char* v = getenv("SKG_PEEL_REFILL")) P.peelRefill = std::max(1, std::min(32, atoi(v)));
    P.sources = e.sourcesDev.as<SourceDev>(); P.Nsources = e.Nsources;
    P.L = e.lumDev.as<double>(); P.Lcdf = e.lumCdfDev.as<double>();
    P.Ltot = phase == SKG_PHASE_STELLAR ? e.lumTotDev.as<double>() : e.dustLtot.as<double>();
    P.emissionBias = e.emissionBias;
    P.instr = e.instrGroupedDev.as<InstrDev>(); P.Ninstr = (int)e.instr.size();
    P.groups = e.groupsDev.as<ObsGroup>(); P.Ngroups = e.Ngroups; P.maxGroupCount = e.maxGroupCount;
    P.persp = e.perspDev.as<PerspDev>(); P.Npersp = e.Npersp;
    if (P.Npersp > 0 && p.continuousScattering) throw Error("continuous scattering with a PerspectiveInstrument is not supported");
    P.dustLv = e.dustLv.as<double>(); P.dustCdf = e.dustCdf.as<double>(); P.dustBias = dustBias;
    const bool store = phase == SKG_PHASE_STELLAR ? p.storeAbsorption != 0 : phase == SKG_PHASE_DUST_SELFABS;
    if (store)
    {
        if (!e.med.rho) throw Error("absorption rates can only be stored with a dust system");
        int64_t count = (int64_t)e.Ncells * Nlambda;
        if (e.labsCount != count)
        {
            // (re)allocate both tables consistently
            e.labs.release(); e.labsDust.release(); e.labsCount = count; e.accLabs = e.accLabsDust = Engine::ACC_ZERO;
        }
        DevBuf& tab = phase == SKG_PHASE_STELLAR ? e.labs : e.labsDust;
        if (!tab.p) { tab.ensure(sizeof(double) * count); SKG_CUDA(cudaMemsetAsync(tab.p, 0, sizeof(double) * count, e.stream)); }
        P.labs = tab.as<double>();
        e.touched(phase == SKG_PHASE_STELLAR ? e.accLabs : e.accLabsDust);
    }
    if (!e.instr.empty() && phase != SKG_PHASE_DUST_SELFABS) e.touched(e.accInstr);
    P.NppInt = (unsigned long long)std::ceil(p.packages);
    P.Lscale = p.luminosityScale > 0 ? p.luminosityScale : (double)P.NppInt;
    P.minWeightReduction = p.minWeightReduction; P.minfs = p.minScattEvents; P.xi = p.scattBias; P.contScatt = p.continuousScattering != 0;
    P.seed = p.seed; P.streamOffset = p.streamOffset;

    // wavelengths with luminosity, in shooting order (the reference skips the others, MonteCarloSimulation.cpp:269,298)
    std::vector<int> ells;
    for (int ell = p.ellBegin; ell < p.ellEnd; ell++) if (LtotHost[ell] > 0) ells.push_back(ell);
    unsigned long long total = P.NppInt * (unsigned long long)ells.size();

    Counters before = e.readCounters();
    cudaEvent_t ev0, ev1; SKG_CUDA(cudaEventCreate(&ev0)); SKG_CUDA(cudaEventCreate(&ev1));
    SKG_CUDA(cudaEventRecord(ev0, e.stream));
    if (total > 0)
    {
        e.mcEllList.upload(ells.data(), sizeof(int) * ells.size(), e.stream);
        P.ellList = e.mcEllList.as<int>();
        // the pool of in-flight packets
        int pool = p.poolPackets > 0 ? p.poolPackets : (1 << 22);
        pool = (int)std::min<unsigned long long>((unsigned long long)pool, total);
        pool = std::max(pool, 1);
        if ((long long)pool * std::max(1, e.Ngroups) > 2000000000LL) throw Error("packet pool times observer directions exceeds the 32-bit work index: lower poolPackets");
        e.mcPool.ensure(2 * sizeof(Packet) * (size_t)pool);       // two pools: the stages ping-pong between them
        if (P.med.Ntheta > 0) e.mcPolPool.ensure(2 * sizeof(PolState) * (size_t)pool);
        e.mcCounts.ensure(sizeof(int) * 8);
        if (!e.mcHostCounts) SKG_CUDA(cudaMallocHost(&e.mcHostCounts, 2 * sizeof(int)));

        size_t smem = 0; bool cartSmem = false;
        if (e.gridKind == GRID_CART)
        {
            size_t need = sizeof(double) * SKG_CART_SMEM_DOUBLES(e.cart);
            if (need > SKG_CART_SMEM_MAX) throw Error("CartesianDustGrid: more than 8189 mesh borders in total are not supported");
            smem = need; cartSmem = true;
            if (!e.attrStages)
            {
                const int cap = 96 * 1024;
#define SKG_ATTR(F) SKG_CUDA(cudaFuncSetAttribute(F, cudaFuncAttributeMaxDynamicSharedMemorySize, cap))
                SKG_ATTR((peelStage<GRID_CART, true, false>)); SKG_ATTR((peelStage<GRID_CART, false, false>));
                SKG_ATTR((peelStage<GRID_CART, true, true>)); SKG_ATTR((peelStage<GRID_CART, false, true>));
                SKG_ATTR((contPeelStage<GRID_CART, true, false>)); SKG_ATTR((contPeelStage<GRID_CART, false, false>));
                SKG_ATTR((contPeelStage<GRID_CART, true, true>)); SKG_ATTR((contPeelStage<GRID_CART, false, true>));
                SKG_ATTR((absorbStage<GRID_CART, true, true, false>)); SKG_ATTR((absorbStage<GRID_CART, true, false, false>));
                SKG_ATTR((absorbStage<GRID_CART, false, true, false>)); SKG_ATTR((absorbStage<GRID_CART, false, false, false>));
                SKG_ATTR((absorbStage<GRID_CART, true, true, true>)); SKG_ATTR((absorbStage<GRID_CART, true, false, true>));
                SKG_ATTR((absorbStage<GRID_CART, false, true, true>)); SKG_ATTR((absorbStage<GRID_CART, false, false, true>));
#undef SKG_ATTR
                SKG_CUDA(cudaFuncSetAttribute(propagateStage<GRID_CART, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                SKG_CUDA(cudaFuncSetAttribute(propagateStage<GRID_CART, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                e.attrStages = true;
            }
        }
        GridSetMC G; G.cart = e.cart; G.tree = e.tree; G.amesh = e.amesh; G.voro = e.voro; G.sym = e.sym;
        switch (e.gridKind)
        {
        case GRID_CART: shootWavefront<GRID_CART>(e, G, P, total, pool, smem, cartSmem); break;
        case GRID_TREE: shootWavefront<GRID_TREE>(e, G, P, total, pool, 0, false); break;
        case GRID_AMESH: shootWavefront<GRID_AMESH>(e, G, P, total, pool, 0, false); break;
        case GRID_VORO: shootWavefront<GRID_VORO>(e, G, P, total, pool, 0, false); break;
        case GRID_SYM: shootWavefront<GRID_SYM>(e, G, P, total, pool, 0, false); break;
        default:
            if (e.med.rho || phase != SKG_PHASE_STELLAR) throw Error("no dust grid has been set");
            shootWavefront<GRID_CART>(e, G, P, total, pool, 0, false);      // no dust: only launch + emission peel-off run
        }
    }
    SKG_CUDA(cudaEventRecord(ev1, e.stream));
    e.sync();
    float ms = 0; SKG_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    Counters after = e.readCounters();
    if (after.errors != before.errors) throw Error("the shooting kernels met a condition on which the reference throws a fatal error (" + std::to_string(after.errors - before.errors) + " times)");
    if (stats)
    {
        stats->packets = after.packets - before.packets; stats->pathSegments = after.segments - before.segments;
        stats->paths = after.paths - before.paths; stats->scatterings = after.scatterings - before.scatterings;
        stats->kernel_ms = ms;
        stats->absorbSegments = after.absorbSegments - before.absorbSegments; stats->detections = after.detections - before.detections;
        stats->launch_ms = e.stageMs[0]; stats->peel_ms = e.stageMs[1]; stats->absorb_ms = e.stageMs[2]; stats->propagate_ms = e.stageMs[3];
        stats->iterations = e.mcIterations;
        stats->peelSegments = after.peelSegments - before.peelSegments; stats->propagateSegments = after.propSegments - before.propSegments;
    }
}

// Geometry::density for the supported geometries
static __device__ double geometryDensity(const SourceDev& s, double x, double y, double z)
{
    double rho;
    if (s.geometry == SKG_GEOM_EXPDISK)
    {
        // ExpDiskGeometry::density, ExpDiskGeometry.cpp:117-129
        const double hR = s.p[0], hz = s.p[1], Rmax = s.p[2], zmax = s.p[3], Rmin = s.p[4];
        double R = sqrt(x * x + y * y), absz = fabs(z);
        if ((Rmax > 0.0 && R > Rmax) || (zmax > 0.0 && absz > zmax) || R < Rmin) rho = 0.0;
        else rho = s.rho0 * exp(-R / hR) * exp(-absz / hz);
    }
    else
    {
        // SpheroidalGeometryDecorator::density (:47-52) around SersicGeometry::density (:66-70), SersicFunction::operator() (:82-93)
        const double reff = s.p[0], q = s.p[1];
        double zz = z / q;
        double sr = sqrt(x * x + y * y + zz * zz) / reff;
        int Ns = s.ntab; double S;
        if (sr <= s.rv[0]) S = s.Sv[0];
        else if (sr >= s.rv[Ns - 1]) S = s.Sv[Ns - 1];
        else { int i = locateClip(s.rv, sr, Ns); S = interpLogLog(sr, s.rv[i], s.rv[i + 1], s.Sv[i], s.Sv[i + 1]); }
        rho = s.rho0 * S / q;
    }
    if (s.spiral_arms > 0)
    {
        // SpiralStructureGeometryDecorator::density + perturbation, SpiralStructureGeometryDecorator.cpp:49-58,224-229
        double R = sqrt(x * x + y * y), phi = atan2(y, x);
        double gamma = log(R / s.spiral_radius) / s.spiral_tanp + s.spiral_phase + 0.5 * M_PI / s.spiral_arms;
        rho *= (1.0 - s.spiral_weight) + s.spiral_weight * s.spiral_cn * pow(sin(0.5 * s.spiral_arms * (gamma - phi)), 2 * s.spiral_index);
    }
    return rho;
}

// DustSystem::setSampleDensityBody, DustSystem.cpp:152-177: one thread per cell
template<int KIND>
__global__ void __launch_bounds__(128) sampleDensityKernel(const __grid_constant__ GridSetMC G, const SourceDev* __restrict__ geoms, const double* __restrict__ norm,
                                                           int Ncells, int Ncomp, int sampleCount, unsigned long long seed, double* __restrict__ rho, Counters* ctr)
{
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Ncells; m += gridDim.x * blockDim.x)
    {
        Philox rng; rng.init(seed, (unsigned long long)m, 7u);
        double sum[8];
        for (int h = 0; h < Ncomp; h++) sum[h] = 0;
        for (int n = 0; n < sampleCount; n++)
        {
            double x, y, z;
            if (!randomPositionInCell<KIND>(G, m, rng, x, y, z)) { atomicAdd(&ctr->errors, 1ull); break; }
            for (int h = 0; h < Ncomp; h++) sum[h] += norm[h] * geometryDensity(geoms[h], x, y, z);
        }
        for (int h = 0; h < Ncomp; h++) rho[(size_t)m * Ncomp + h] = sum[h] / sampleCount;
    }
}

void mcSampleDensity(Engine& e, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* rho)
{
    if (e.gridKind == GRID_NONE) throw Error("no dust grid has been set");
    if (Ncomp < 1 || Ncomp > 8 || !geoms || !norm || !rho) throw Error("skg_sample_density: bad arguments");
    if (sampleCount < 1) throw Error("the number of random density samples should be positive");
    if (e.gridKind == GRID_VORO && !e.voro.cellBox) throw Error("the Voronoi grid was given without cell boxes (needed by randomPositionInCell)");
    std::vector<DevBuf*> bufs; std::vector<SourceDev> dev;
    DevBuf devGeoms, devNorm, devRho;
    try
    {
        for (int h = 0; h < Ncomp; h++) dev.push_back(makeSourceDev(e, geoms[h], bufs, true));
        devGeoms.upload(dev.data(), sizeof(SourceDev) * Ncomp, e.stream); devNorm.upload(norm, sizeof(double) * Ncomp, e.stream);
        devRho.ensure(sizeof(double) * (size_t)e.Ncells * Ncomp);
        GridSetMC G; G.cart = e.cart; G.tree = e.tree; G.amesh = e.amesh; G.voro = e.voro; G.sym = e.sym;
        Counters before = e.readCounters();
        int blocks = std::max(1, std::min((e.Ncells + 127) / 128, e.smCount * 16));
#define SKG_DENS(K) sampleDensityKernel<K><<<blocks, 128, 0, e.stream>>>(G, devGeoms.as<SourceDev>(), devNorm.as<double>(), e.Ncells, Ncomp, sampleCount, seed, devRho.as<double>(), e.ctr())
        switch (e.gridKind) { case GRID_CART: SKG_DENS(GRID_CART); break; case GRID_TREE: SKG_DENS(GRID_TREE); break;
                              case GRID_AMESH: SKG_DENS(GRID_AMESH); break; case GRID_SYM: SKG_DENS(GRID_SYM); break; default: SKG_DENS(GRID_VORO); }
        e.launches++; SKG_CUDA(cudaGetLastError());
        SKG_CUDA(cudaMemcpyAsync(rho, devRho.p, sizeof(double) * (size_t)e.Ncells * Ncomp, cudaMemcpyDeviceToHost, e.stream));
        e.sync();
        if (e.readCounters().errors != before.errors) throw Error("Can't find random position in cell");      // VoronoiMesh.cpp:606
    }
    catch (...) { for (DevBuf* b : bufs) delete b; throw; }
    for (DevBuf* b : bufs) delete b;
}

// TreeNodeSampleDensityCalculator (TreeNodeSampleDensityCalculator.cpp:25-45) for a batch of boxes: one thread per box draws
// sampleCount positions (Random::position(Box): x, y, z in this order) and averages the total density of the components;
// mass = mean density x volume -- the quantity TreeDustGrid::subdivide compares with maxMassFraction (TreeDustGrid.cpp:197-201)
// DISP: besides the mass (same operations in the same order, so the mass of a box does not depend on DISP) the spread of the
// sampled densities, TreeNodeSampleDensityCalculator::densityDispersion (TreeNodeSampleDensityCalculator.cpp:62-67)
template<bool DISP>
__global__ void __launch_bounds__(128) sampleBoxesKernel(const double* __restrict__ box, int64_t n, const SourceDev* __restrict__ geoms,
                                                         const double* __restrict__ norm, int Ncomp, int sampleCount, unsigned long long seed,
                                                         double* __restrict__ mass, double* __restrict__ dispersion)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    {
        const double* b = box + 6 * q;
        const double x0 = b[0], y0 = b[1], z0 = b[2], wx = b[3] - b[0], wy = b[4] - b[1], wz = b[5] - b[2];
        Philox rng; rng.init(seed, (unsigned long long)q, 9u);
        double sum = 0, minrho = SKG_DBL_MAX, maxrho = 0;
        for (int s = 0; s < sampleCount; s++)
        {
            const double fx = rng.uniform(), fy = rng.uniform(), fz = rng.uniform();
            const double x = x0 + fx * wx, y = y0 + fy * wy, z = z0 + fz * wz;
            double rho = 0;
            for (int h = 0; h < Ncomp; h++)
            {
                const double d = geometryDensity(geoms[h], x, y, z);
                sum += norm[h] * d;
                if (DISP) rho += norm[h] * d;
            }
            if (DISP) { minrho = fmin(minrho, rho); maxrho = fmax(maxrho, rho); }
        }
        mass[q] = sum / sampleCount * (wx * wy * wz);
        if (DISP) dispersion[q] = maxrho > 0 ? (maxrho - minrho) / maxrho : 0;
    }
}

void mcSampleBoxes(Engine& e, int64_t n, const double* box, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* mass, double* dispersion)
{
    if (n < 0 || (n > 0 && (!box || !mass)) || Ncomp < 1 || Ncomp > 8 || !geoms || !norm) throw Error("skg_sample_boxes: bad arguments");
    if (sampleCount < 1) throw Error("Number of random samples must be at least 1");       // TreeDustGrid.cpp:60
    if (n == 0) return;
    std::vector<DevBuf*> bufs; std::vector<SourceDev> dev;
    DevBuf devGeoms, devNorm, devBox, devMass, devDisp;
    try
    {
        for (int h = 0; h < Ncomp; h++) dev.push_back(makeSourceDev(e, geoms[h], bufs, true));
        devGeoms.upload(dev.data(), sizeof(SourceDev) * Ncomp, e.stream); devNorm.upload(norm, sizeof(double) * Ncomp, e.stream);
        devBox.upload(box, sizeof(double) * 6 * (size_t)n, e.stream); devMass.ensure(sizeof(double) * (size_t)n);
        if (dispersion) devDisp.ensure(sizeof(double) * (size_t)n);
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)e.smCount * 16));
        if (dispersion)
            sampleBoxesKernel<true><<<blocks, 128, 0, e.stream>>>(devBox.as<double>(), n, devGeoms.as<SourceDev>(), devNorm.as<double>(), Ncomp, sampleCount, seed, devMass.as<double>(), devDisp.as<double>());
        else
            sampleBoxesKernel<false><<<blocks, 128, 0, e.stream>>>(devBox.as<double>(), n, devGeoms.as<SourceDev>(), devNorm.as<double>(), Ncomp, sampleCount, seed, devMass.as<double>(), nullptr);
        e.launches++; SKG_CUDA(cudaGetLastError());
        SKG_CUDA(cudaMemcpyAsync(mass, devMass.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
        if (dispersion) SKG_CUDA(cudaMemcpyAsync(dispersion, devDisp.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
        e.sync();
    }
    catch (...) { for (DevBuf* b : bufs) delete b; throw; }
    for (DevBuf* b : bufs) delete b;
}

// fp64 atomic adds to pseudo-random cells of a table, the access pattern of the escape + absorption stage: the rate at which
// this GPU retires them is the ceiling of that stage (skg_selftest_atomics)
__global__ void __launch_bounds__(128) atomicRateKernel(double* __restrict__ table, unsigned cells, unsigned long long perThread, unsigned long long seed)
{
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    for (unsigned long long i = 0; i < perThread; i++)
    {
        z += 0x9E3779B97F4A7C15ull; unsigned long long t = z; t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull; t = (t ^ (t >> 27)) * 0x94D049BB133111EBull; t ^= t >> 31;
        atomicAdd(table + (unsigned)(t % cells), 1.0);
    }
}

double mcAtomicRate(Engine& e, uint64_t n, int cells)
{
    if (cells < 1 || n < 1) throw Error("skg_selftest_atomics: bad arguments");
    DevBuf table; table.ensure(sizeof(double) * (size_t)cells);
    SKG_CUDA(cudaMemsetAsync(table.p, 0, sizeof(double) * (size_t)cells, e.stream));
    const int blocks = e.smCount * 16; const unsigned long long per = std::max<unsigned long long>(1, n / ((unsigned long long)blocks * 128));
    cudaEvent_t ev0, ev1; SKG_CUDA(cudaEventCreate(&ev0)); SKG_CUDA(cudaEventCreate(&ev1));
    atomicRateKernel<<<blocks, 128, 0, e.stream>>>(table.as<double>(), (unsigned)cells, per / 8 + 1, 1ull);      // warm-up
    SKG_CUDA(cudaEventRecord(ev0, e.stream));
    atomicRateKernel<<<blocks, 128, 0, e.stream>>>(table.as<double>(), (unsigned)cells, per, 2ull);
    SKG_CUDA(cudaEventRecord(ev1, e.stream));
    e.launches += 2; SKG_CUDA(cudaGetLastError()); e.sync();
    float ms = 0; SKG_CUDA(cudaEventElapsedTime(&ms, ev0, ev1)); cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    return (double)per * blocks * 128 / (ms * 1e-3);
}

__global__ void unpackLaunches(const Packet* __restrict__ pool, int n, double* __restrict__ r, double* __restrict__ k, double* __restrict__ L)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const Packet pk = loadPacket(pool + i);
        r[3 * i] = pk.x; r[3 * i + 1] = pk.y; r[3 * i + 2] = pk.z; k[3 * i] = pk.kx; k[3 * i + 1] = pk.ky; k[3 * i + 2] = pk.kz; L[i] = pk.L;
    }
}

// n launches of StellarSystem::launch(pp, ell, L = 1) exactly as the shooting phase performs them (same kernel), for
// distribution-level checks of the samplers
void mcSampleLaunch(Engine& e, int ell, int n, uint64_t seed, double* r, double* k, double* L)
{
    if (!e.Nsources) throw Error("no sources have been set");
    if (ell < 0 || ell >= e.NlambdaSrc) throw Error("wavelength index out of range");
    if (n < 1 || !r || !k || !L) throw Error("skg_sample_launch: bad arguments");
    McDev P{};
    P.med.Nlambda = e.NlambdaSrc;
    P.sources = e.sourcesDev.as<SourceDev>(); P.Nsources = e.Nsources;
    P.L = e.lumDev.as<double>(); P.Lcdf = e.lumCdfDev.as<double>(); P.Ltot = e.lumTotDev.as<double>();
    P.emissionBias = e.emissionBias; P.phase = SKG_PHASE_STELLAR; P.rngKind = 0;
    P.NppInt = (unsigned long long)n; P.Lscale = e.lumTotHost[ell];      // unit luminosity per packet before the bias weight
    P.seed = seed; P.streamOffset = 0;
    e.mcEllList.upload(&ell, sizeof(int), e.stream); P.ellList = e.mcEllList.as<int>();
    e.mcPool.ensure(sizeof(Packet) * (size_t)n); P.pool = e.mcPool.as<Packet>(); P.poolNext = P.pool;
    e.scratchR.ensure(sizeof(double) * 3 * (size_t)n); e.scratchK.ensure(sizeof(double) * 3 * (size_t)n); e.scratchTau.ensure(sizeof(double) * (size_t)n);
    int blocks = std::max(1, std::min((n + 127) / 128, e.smCount * 16));
    launchStage<<<blocks, 128, 0, e.stream>>>(P, e.ctr(), n, 0ull, 0);
    unpackLaunches<<<blocks, 128, 0, e.stream>>>(P.pool, n, e.scratchR.as<double>(), e.scratchK.as<double>(), e.scratchTau.as<double>());
    e.launches += 2; SKG_CUDA(cudaGetLastError());
    SKG_CUDA(cudaMemcpyAsync(r, e.scratchR.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
    SKG_CUDA(cudaMemcpyAsync(k, e.scratchK.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
    SKG_CUDA(cudaMemcpyAsync(L, e.scratchTau.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
    e.sync();
}

void mcRunStellar(Engine& e, const skg_mc_params& p, skg_mc_stats* stats)
{
    if (e.gridKind == GRID_NONE && e.med.rho) throw Error("no dust grid has been set");
    if (!e.Nsources) throw Error("no sources have been set");
    if (e.med.Nlambda && e.NlambdaSrc != e.med.Nlambda) throw Error("sources and medium disagree on the number of wavelengths");
    runPhase(e, p, SKG_PHASE_STELLAR, 0.0, e.lumTotHost, stats);
}

// PanMonteCarloSimulation::dodustselfabsorptionchunk / dodustemissionchunk for every wavelength
void mcRunDust(Engine& e, const skg_mc_params& p, int phase, double emissionBias, int mem, const double* Lcell, skg_mc_stats* stats)
{
    if (phase != SKG_PHASE_DUST_SELFABS && phase != SKG_PHASE_DUST_EMISSION) throw Error("skg_run_dust: phase must be SKG_PHASE_DUST_SELFABS or SKG_PHASE_DUST_EMISSION");
    if (e.gridKind == GRID_NONE || !e.med.rho) throw Error("dust emission needs a dust grid and a medium");
    if (!Lcell) throw Error("skg_run_dust: null cell luminosities");
    if (!(emissionBias >= 0 && emissionBias < 1)) throw Error("emissionBias should be between 0 and 1");
    if (e.gridKind == GRID_VORO && !e.voro.cellBox) throw Error("the Voronoi grid was given without cell boxes (needed by randomPositionInCell)");
    const int Nlambda = e.med.Nlambda, Ncells = e.Ncells;
    const size_t count = (size_t)Nlambda * Ncells;
    if (mem == SKG_HOST) e.dustLv.upload(Lcell, sizeof(double) * count, e.stream);
    else { e.dustLv.ensure(sizeof(double) * count); SKG_CUDA(cudaMemcpyAsync(e.dustLv.p, Lcell, sizeof(double) * count, cudaMemcpyDeviceToDevice, e.stream)); }
    e.dustCdf.ensure(sizeof(double) * (size_t)Nlambda * (Ncells + 1));
    e.dustLtot.ensure(sizeof(double) * Nlambda);
    mcBuildCdfs(e, Nlambda, Ncells);
    std::vector<double> Ltot(Nlambda);
    SKG_CUDA(cudaMemcpyAsync(Ltot.data(), e.dustLtot.p, sizeof(double) * Nlambda, cudaMemcpyDeviceToHost, e.stream));
    e.sync();
    for (double& v : Ltot) if (!(v > 0) || !std::isfinite(v)) v = 0;
    runPhase(e, p, phase, emissionBias, Ltot, stats);
}

}   // namespace skg
