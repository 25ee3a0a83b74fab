// Batched DustGrid::path() kernels: count pass, CSR fill pass, optical depth, whichcell.
// One ray per thread in a grid-stride loop; the Cartesian borders (3 x (N+1) doubles, 2.4 KB at 100^3)
// are staged in shared memory once per CTA.
#include <cub/device/device_scan.cuh>
#include "engine.h"
#include "geom.cuh"
#include "sinks.cuh"

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

template<int KIND, class Sink>
__device__ __forceinline__ void walk(const GridSet& G, const CartGrid& cart, Counters* ctr,
                                     double x, double y, double z, double kx, double ky, double kz, Sink& sink)
{
    if (KIND == GRID_CART) walkCart(cart, x, y, z, kx, ky, kz, sink);
    else if (KIND == GRID_TREE) walkTree(G.tree, ctr, x, y, z, kx, ky, kz, sink);
    else if (KIND == GRID_AMESH) walkAMesh(G.amesh, ctr, x, y, z, kx, ky, kz, sink);
    else walkVoro(G.voro, ctr, x, y, z, kx, ky, kz, sink);
}

template<int KIND>
__global__ void __launch_bounds__(128) pathCountKernel(const __grid_constant__ GridSet G, Counters* ctr, bool cartSmem, int64_t n,
                                                       const double* __restrict__ r, const double* __restrict__ k,
                                                       int* __restrict__ counts)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        CountSink sink;
        walk<KIND>(G, cart, ctr, r[3 * i], r[3 * i + 1], r[3 * i + 2], k[3 * i], k[3 * i + 1], k[3 * i + 2], sink);
        counts[i] = sink.n;
    }
}

template<int KIND>
__global__ void __launch_bounds__(128) pathFillKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem,
                                                      int64_t n, const double* __restrict__ r, const double* __restrict__ k,
                                                      const int* __restrict__ ell, int ellStride, const int64_t* __restrict__ offsets,
                                                      int* __restrict__ m, double* __restrict__ ds, double* __restrict__ s,
                                                      double* __restrict__ dtau, double* __restrict__ tau)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        int64_t o = offsets[i];
        RecordSink sink;
        sink.m = m + o; sink.ds = ds + o; sink.s = s + o; sink.dtau = dtau + o; sink.tau = tau + o;
        sink.optical = (ell != nullptr);
        int l = sink.optical ? ell[i * ellStride] : 0;
        sink.kr = KappaRho{med.rho, med.kext + l, med.Ncomp, med.Nlambda};
        walk<KIND>(G, cart, ctr, r[3 * i], r[3 * i + 1], r[3 * i + 2], k[3 * i], k[3 * i + 1], k[3 * i + 2], sink);
    }
}

template<int KIND>
__global__ void __launch_bounds__(128) opticalDepthKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem,
                                                          int64_t n, const double* __restrict__ r, const double* __restrict__ k,
                                                          const int* __restrict__ ell, int ellStride,
                                                          const double* __restrict__ dist, double* __restrict__ tau)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        TauSink sink;
        sink.kr = KappaRho{med.rho, med.kext + ell[i * ellStride], med.Ncomp, med.Nlambda};
        sink.distance = dist ? dist[i] : SKG_DBL_MAX;
        walk<KIND>(G, cart, ctr, r[3 * i], r[3 * i + 1], r[3 * i + 2], k[3 * i], k[3 * i + 1], k[3 * i + 2], sink);
        tau[i] = sink.tau;
    }
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

struct LaunchCfg { int blocks; size_t smem; bool cartSmem; };
static LaunchCfg cfgFor(const Engine& e, int64_t n)
{
    LaunchCfg c;
    int64_t want = (n + 127) / 128;
    int64_t cap = (int64_t)e.smCount * 16;
    c.blocks = (int)std::max<int64_t>(1, std::min(want, cap));
    c.smem = 0; c.cartSmem = false;
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
    SKG_DISPATCH(e, (pathCountKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.ctr(), c.cartSmem, n, d_r, d_k, d_counts)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchPathFill(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                    const int64_t* d_offsets, int* d_m, double* d_ds, double* d_s, double* d_dtau, double* d_tau)
{
    if (n <= 0) return;
    if (d_ell && !e.med.rho) throw Error("skg_path_fill with wavelength indices needs skg_medium first");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (pathFillKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, n, d_r, d_k, d_ell, ellStride,
                                                                                d_offsets, d_m, d_ds, d_s, d_dtau, d_tau)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchOpticalDepth(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                        const double* d_dist, double* d_tau)
{
    if (n <= 0) return;
    if (!e.med.rho) throw Error("skg_opticaldepth needs skg_medium first");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (opticalDepthKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, n, d_r, d_k, d_ell, ellStride,
                                                                                    d_dist, d_tau)));
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
