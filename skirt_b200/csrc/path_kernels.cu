// Batched DustGrid::path() kernels: count pass, CSR fill pass, optical depth, whichcell.
// One ray per thread in a grid-stride loop; the Cartesian borders (3 x (N+1) doubles, 2.4 KB at 100^3)
// are staged in shared memory once per CTA.
#include <cub/device/device_scan.cuh>
#include "engine.h"
#include "geom.cuh"
#include "sinks.cuh"
#include "wavefront.cuh"

namespace skg
{

struct GridSet
{
    CartGrid cart; TreeGrid tree; AMeshGrid amesh; VoroGrid voro;
};

// shared-memory staging of the Cartesian borders; returns a CartGrid view whose xv/yv/zv point to smem
__device__ __forceinline__ CartGrid stageCart(const CartGrid& g, double* smem, bool useSmem)
{
    if (!useSmem) return g;
    CartGrid s = g;
    int nx = g.Nx + 1, ny = g.Ny + 1, nz = g.Nz + 1;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) smem[i] = g.xv[i];
    for (int i = threadIdx.x; i < ny; i += blockDim.x) smem[nx + i] = g.yv[i];
    for (int i = threadIdx.x; i < nz; i += blockDim.x) smem[nx + ny + i] = g.zv[i];
    __syncthreads();
    s.xv = smem; s.yv = smem + nx; s.zv = smem + nx + ny;
    return s;
}

// ---- jobs (see wavefront.cuh) ---------------------------------------------------------------------------
struct RayJobBase
{
    const double* r; const double* k;
    double rx, ry, rz, dx, dy, dz;
    __device__ __forceinline__ void loadRay(int i)
    { rx = r[3 * (size_t)i]; ry = r[3 * (size_t)i + 1]; rz = r[3 * (size_t)i + 2]; dx = k[3 * (size_t)i]; dy = k[3 * (size_t)i + 1]; dz = k[3 * (size_t)i + 2]; }
    __device__ __forceinline__ void collective(bool) {}
    __device__ __forceinline__ void periodic() {}
};

// first pass of the batched path(): number of segments of every ray
struct CountJob : RayJobBase
{
    int* counts; int item, n;
    __device__ __forceinline__ int begin(int i) { item = i; n = 0; loadRay(i); return 1; }
    __device__ __forceinline__ bool outside(double) { n++; return true; }
    __device__ __forceinline__ bool segment(int, double) { n++; return true; }
    __device__ __forceinline__ void finish() { counts[item] = n; }
};

// DustSystem::opticaldepth(pp, distance), DustSystem.cpp:984-1000 + DustGridPath::opticalDepth, DustGridPath.hpp:97-108:
// the overshooting segment is counted in full, then the walk stops
struct TauJob : RayJobBase
{
    const int* ell; int ellStride; Medium med; const double* dist; double* out;
    KappaRho kr; double distance, sacc, tau; int item;
    __device__ __forceinline__ int begin(int i)
    {
        item = i; loadRay(i);
        kr = KappaRho{med.rho, med.kext + ell[(size_t)i * ellStride], med.Ncomp, med.Nlambda};
        distance = dist ? dist[i] : SKG_DBL_MAX; sacc = 0; tau = 0;
        return 1;
    }
    __device__ __forceinline__ bool outside(double d) { sacc += d; return !(sacc > distance); }
    __device__ __forceinline__ bool segment(int mm, double d) { sacc += d; tau += kr(mm) * d; return !(sacc > distance); }
    __device__ __forceinline__ void finish() { out[item] = tau; }
};

template<int KIND>
__global__ void __launch_bounds__(128) pathCountKernel(const __grid_constant__ GridSet G, Counters* ctr, bool cartSmem, int n,
                                                       const double* __restrict__ r, const double* __restrict__ k,
                                                       int* __restrict__ counts, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    CountJob job; job.r = r; job.k = k; job.counts = counts;
    runJobs<KIND>(G, cart, ctr, job, n, work);
}

// second pass: Segment{m, ds, s, dtau, tau} records = DustGridPath::addSegment (DustGridPath.cpp:46-53, running
// length s) + DustGridPath::fillOpticalDepth (DustGridPath.hpp:117-129, running tau), through a per-warp
// shared-memory stage.  A crossing step only parks
// (m, ds) in a ring of 8 entries per lane; every SKG_PERIOD steps each lane (1) gathers rho for all its parked
// entries at once (several independent loads in flight instead of one per step) and extends its running s and tau
// in path order, and (2) the warp writes the finished entries out together: 8 lanes per source lane, only whole
// 32-byte sectors (4 aligned doubles) except at the two ends of a path, so that L2 never has to merge partial
// sectors and DRAM sees each byte once.
#define SKG_RING 8
#define SKG_RSTRIDE 9       // doubles per lane in a ring (8 + 1 pad against bank conflicts)
struct RecordJobStaged : RayJobBase
{
    const int64_t* offsets; const int* ell; int ellStride; Medium med;
    int* m; double* ds; double* s; double* dtau; double* tau;
    // this lane's rings in shared memory
    double* rDs; double* rS; double* rDtau; double* rTau; int* rM;
    // the warp's rings (lane 0), for the cooperative write
    double* wDs; double* wS; double* wDtau; double* wTau; int* wM;
    KappaRho kr; int64_t o, f, c; double sacc, tacc; bool optical;

    __device__ __forceinline__ void bind(char* warpBase)
    {
        const int lane = threadIdx.x & 31;
        wDs = reinterpret_cast<double*>(warpBase); wS = wDs + 32 * SKG_RSTRIDE; wDtau = wS + 32 * SKG_RSTRIDE; wTau = wDtau + 32 * SKG_RSTRIDE;
        wM = reinterpret_cast<int*>(wTau + 32 * SKG_RSTRIDE);
        rDs = wDs + lane * SKG_RSTRIDE; rS = wS + lane * SKG_RSTRIDE; rDtau = wDtau + lane * SKG_RSTRIDE; rTau = wTau + lane * SKG_RSTRIDE;
        rM = wM + lane * SKG_RSTRIDE;
        o = f = c = 0; sacc = tacc = 0;
    }
    static constexpr size_t bytesPerWarp() { return (4 * sizeof(double) + sizeof(int)) * 32 * SKG_RSTRIDE; }

    __device__ __forceinline__ int begin(int i)
    {
        loadRay(i);
        o = f = c = offsets[i]; sacc = 0; tacc = 0;
        optical = ell != nullptr;
        int l = optical ? ell[(size_t)i * ellStride] : 0;
        kr = KappaRho{med.rho, med.kext + l, med.Ncomp, med.Nlambda};
        return 1;
    }
    __device__ __forceinline__ bool outside(double d) { int q = (int)(o & (SKG_RING - 1)); rM[q] = -1; rDs[q] = d; o++; return true; }
    __device__ __forceinline__ bool segment(int mm, double d) { int q = (int)(o & (SKG_RING - 1)); rM[q] = mm; rDs[q] = d; o++; return true; }

    // running s (DustGridPath::addSegment) and dtau/tau (fillOpticalDepth) for the parked entries [c, o), in path order
    __device__ __forceinline__ void compute()
    {
        while (c < o)
        {
            double krv[4], dv[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const bool valid = c + u < o;
                const int q = (int)((c + u) & (SKG_RING - 1));
                const int mm = valid ? rM[q] : -1;
                dv[u] = valid ? rDs[q] : 0.0;
                krv[u] = (optical && valid) ? kr(mm) : 0.0;         // kapparho(-1) = 0 (DustSystem.cpp:918-921)
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                if (c + u < o)
                {
                    const int q = (int)((c + u) & (SKG_RING - 1));
                    sacc += dv[u];
                    const double dt = krv[u] * dv[u];
                    tacc += dt;
                    rS[q] = sacc; rDtau[q] = dt; rTau[q] = tacc;
                }
            }
            c = (o - c > 4) ? c + 4 : o;
        }
    }

    // the warp writes entries [f, e) of every lane that has some: 8 writer lanes per source lane
    __device__ __forceinline__ void writeOut(int64_t e)
    {
        const unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        __syncwarp();
        unsigned mask = __ballot_sync(FULL, e > f);
        const int sub = lane >> 3, j = lane & 7;
        while (mask)
        {
            // the next (up to) four source lanes
            int src = -1; unsigned rem = mask;
            for (int t = 0; t < 4; t++)
            {
                int bit = rem ? __ffs(rem) - 1 : -1;
                if (t == sub) src = bit;
                if (rem) rem &= rem - 1;
            }
            mask = rem;
            const int sl = src < 0 ? 0 : src;
            const int64_t fs = __shfl_sync(FULL, f, sl), es = __shfl_sync(FULL, e, sl);
            const int64_t idx = fs + j;
            if (src >= 0 && idx < es)
            {
                const int q = (int)(idx & (SKG_RING - 1)) + src * SKG_RSTRIDE;
                m[idx] = wM[q]; ds[idx] = wDs[q]; s[idx] = wS[q]; dtau[idx] = wDtau[q]; tau[idx] = wTau[q];
            }
        }
        if (e > f) f = e;
        __syncwarp();
    }
    __device__ __forceinline__ void periodic() { compute(); writeOut(o & ~(int64_t)3); }
    __device__ __forceinline__ void finish() { compute(); }
    __device__ __forceinline__ void collective(bool fin) { writeOut(fin ? o : f); }
};

template<int KIND>
__global__ void __launch_bounds__(128) pathFillKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem,
                                                      int n, const double* __restrict__ r, const double* __restrict__ k,
                                                      const int* __restrict__ ell, int ellStride, const int64_t* __restrict__ offsets,
                                                      int* __restrict__ m, double* __restrict__ ds, double* __restrict__ s,
                                                      double* __restrict__ dtau, double* __restrict__ tau, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    size_t skip = (KIND == GRID_CART && cartSmem) ? (size_t)(G.cart.Nx + G.cart.Ny + G.cart.Nz + 3) : 0;
    RecordJobStaged job; job.r = r; job.k = k; job.offsets = offsets; job.ell = ell; job.ellStride = ellStride; job.med = med;
    job.m = m; job.ds = ds; job.s = s; job.dtau = dtau; job.tau = tau;
    job.bind(reinterpret_cast<char*>(smem + skip) + (threadIdx.x >> 5) * RecordJobStaged::bytesPerWarp());
    runJobs<KIND>(G, cart, ctr, job, n, work);
}

template<int KIND>
__global__ void __launch_bounds__(128) opticalDepthKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem,
                                                          int n, const double* __restrict__ r, const double* __restrict__ k,
                                                          const int* __restrict__ ell, int ellStride,
                                                          const double* __restrict__ dist, double* __restrict__ tau, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    TauJob job; job.r = r; job.k = k; job.ell = ell; job.ellStride = ellStride; job.med = med; job.dist = dist; job.out = tau;
    runJobs<KIND>(G, cart, ctr, job, n, work);
}

template<int KIND>
__global__ void __launch_bounds__(128) whichCellKernel(const __grid_constant__ GridSet G, Counters* ctr, int64_t n,
                                                       const double* __restrict__ r, int* __restrict__ m)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
        int res;
        if (KIND == GRID_CART) res = cartWhichCell(G.cart, x, y, z);
        else if (KIND == GRID_TREE) { int node = treeWhichNode(G.tree, x, y, z); res = node >= 0 ? G.tree.cell[node] : -1; }
        else if (KIND == GRID_AMESH) { int node = ameshWhichNode(G.amesh, x, y, z); res = node >= 0 ? G.amesh.cell[node] : -1; }
        else res = voroCellIndex(G.voro, x, y, z);
        m[i] = res;
    }
}

// ---------------------------------------------------------------------------------------------------
static GridSet gridSet(const Engine& e) { GridSet G; G.cart = e.cart; G.tree = e.tree; G.amesh = e.amesh; G.voro = e.voro; return G; }

struct LaunchCfg { int blocks; size_t smem; bool cartSmem; int* work; };
static LaunchCfg cfgFor(Engine& e, int64_t n)
{
    LaunchCfg c;
    int64_t want = (n + 127) / 128;
    int64_t cap = (int64_t)e.smCount * 16;
    c.blocks = (int)std::max<int64_t>(1, std::min(want, cap));
    c.smem = 0; c.cartSmem = false;
    if (n > 2147483647LL) throw Error("at most 2^31-1 rays per call");
    e.scratchWork.ensure(sizeof(int)); c.work = e.scratchWork.as<int>();
    SKG_CUDA(cudaMemsetAsync(c.work, 0, sizeof(int), e.stream));
    if (e.gridKind == GRID_CART)
    {
        size_t need = sizeof(double) * (size_t)(e.cart.Nx + e.cart.Ny + e.cart.Nz + 3);
        if (need <= 40 * 1024) { c.smem = need; c.cartSmem = true; }
    }
    return c;
}

#define SKG_DISPATCH(e, CALL) \
    switch ((e).gridKind) { \
    case GRID_CART:  { constexpr int K = GRID_CART;  CALL; break; } \
    case GRID_TREE:  { constexpr int K = GRID_TREE;  CALL; break; } \
    case GRID_AMESH: { constexpr int K = GRID_AMESH; CALL; break; } \
    case GRID_VORO:  { constexpr int K = GRID_VORO;  CALL; break; } \
    default: throw Error("no dust grid has been set"); }

void launchPathCount(Engine& e, int64_t n, const double* d_r, const double* d_k, int* d_counts)
{
    if (n <= 0) return;
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (pathCountKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.ctr(), c.cartSmem, (int)n, d_r, d_k, d_counts, c.work)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchPathFill(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                    const int64_t* d_offsets, int* d_m, double* d_ds, double* d_s, double* d_dtau, double* d_tau)
{
    if (n <= 0) return;
    if (d_ell && !e.med.rho) throw Error("skg_path_fill with wavelength indices needs skg_medium first");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    c.smem += 4 * RecordJobStaged::bytesPerWarp();
    static bool attr = false;
    if (!attr)
    {
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_CART>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_TREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_AMESH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_VORO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr = true;
    }
    SKG_DISPATCH(e, (pathFillKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, (int)n, d_r, d_k, d_ell, ellStride,
                                                                                d_offsets, d_m, d_ds, d_s, d_dtau, d_tau, c.work)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchOpticalDepth(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                        const double* d_dist, double* d_tau)
{
    if (n <= 0) return;
    if (!e.med.rho) throw Error("skg_opticaldepth needs skg_medium first");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (opticalDepthKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, (int)n, d_r, d_k, d_ell, ellStride,
                                                                                    d_dist, d_tau, c.work)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchWhichCell(Engine& e, int64_t n, const double* d_r, int* d_m)
{
    if (n <= 0) return;
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (whichCellKernel<K><<<c.blocks, 128, 0, e.stream>>>(G, e.ctr(), n, d_r, d_m)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

__global__ void setLastOffset(const int* counts, int64_t* offsets, int64_t n)
{
    offsets[n] = offsets[n - 1] + counts[n - 1];
}

void exclusiveScan(Engine& e, int64_t n, const int* d_counts, int64_t* d_offsets)
{
    if (n <= 0) { SKG_CUDA(cudaMemsetAsync(d_offsets, 0, sizeof(int64_t), e.stream)); return; }
    size_t tmp = 0;
    // int32 counts -> int64 offsets
    cub::DeviceScan::ExclusiveScan(nullptr, tmp, d_counts, d_offsets, cub::Sum(), (int64_t)0, n, e.stream);
    e.scratchCub.ensure(tmp);
    e.launches += 2;   // cub's scan is two kernels
    SKG_CUDA(cub::DeviceScan::ExclusiveScan(e.scratchCub.p, tmp, d_counts, d_offsets, cub::Sum(), (int64_t)0, n, e.stream));
    setLastOffset<<<1, 1, 0, e.stream>>>(d_counts, d_offsets, n);
    e.launches++; SKG_CUDA(cudaGetLastError());
}

}   // namespace skg
