// Batched DustGrid::path() kernels: count pass, CSR fill pass, optical depth, whichcell.
// One ray per thread in a grid-stride loop; the Cartesian borders (3 x (N+1) doubles, 2.4 KB at 100^3)
// are staged in shared memory once per CTA.
#include <cstdlib>
#include <cub/device/device_scan.cuh>
#include "engine.h"
#include "geom.cuh"
#include "sinks.cuh"
#include "ring.cuh"
#include "wavefront.cuh"

namespace skg
{

struct GridSet
{
    CartGrid cart; TreeGrid tree; AMeshGrid amesh; VoroGrid voro; SymGrid sym;
};

// ---- jobs (see wavefront.cuh) ---------------------------------------------------------------------------
struct RayJobBase
{
    const double* r; const double* k;
    double rx, ry, rz, dx, dy, dz;
    static constexpr bool kCartRegBorders = SKG_CART_REGBORDERS, kCartTinySelect = false; static constexpr bool kTreeHints = SKG_TREE_HINTS_PATH, kCartRhoAhead = false; static constexpr int kBatches = 1;
    static constexpr bool kCartFast = false;        // the deterministic-geometry entry points are bit-exact
    __device__ __forceinline__ void loadRay(int i)
    { rx = r[3 * (size_t)i]; ry = r[3 * (size_t)i + 1]; rz = r[3 * (size_t)i + 2]; dx = k[3 * (size_t)i]; dy = k[3 * (size_t)i + 1]; dz = k[3 * (size_t)i + 2]; }
    __device__ __forceinline__ void collective(bool) {}
    __device__ __forceinline__ void periodic() {}
    __device__ __forceinline__ int cellHint() const { return -1; }      // the deterministic entry points always locate the start of a ray
    __device__ __forceinline__ void noteStart(int) {}
};
// (jobs of the deterministic-geometry kernels never run on a predicated walker, except TauJobT<true>, which forwards)

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
template<bool FAST> struct TauJobT : RayJobBase
{
    static constexpr bool kCartFast = FAST;         // FAST: the walker of the shooting stages (skg_opticaldepth_mc)
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
    template<int U> __device__ __forceinline__ bool segmentU(int mm, double d) { return segment(mm, d); }
    template<int U> __device__ __forceinline__ void idleU() {}
    __device__ __forceinline__ void finish() { out[item] = tau; }
};

template<int KIND>
__global__ void __launch_bounds__(128) pathCountKernel(const __grid_constant__ GridSet G, Counters* ctr, bool cartSmem, int refill, int n,
                                                       const double* __restrict__ r, const double* __restrict__ k,
                                                       int* __restrict__ counts, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    CountJob job; job.r = r; job.k = k; job.counts = counts;
    runJobs<KIND>(G, cart, ctr, job, n, work, refill);
}

// second pass: Segment{m, ds, s, dtau, tau} records = DustGridPath::addSegment (DustGridPath.cpp:46-53, running
// length s) + DustGridPath::fillOpticalDepth (DustGridPath.hpp:117-129, running tau).
// A crossing step only parks (m, ds) in a small per-lane ring in shared memory and starts an asynchronous copy of
// the cell's density into the same ring slot.  Every SKG_PERIOD (= 4) steps the lane turns the entries parked one
// period earlier (whose densities have landed meanwhile) into records four at a time: s and tau are extended in path order, and each
// array receives ONE 256-bit store (st.global.v4.f64, a whole 32-byte sector; 128-bit for m) at a 4-aligned
// record index -- so L2 never merges partial sectors and DRAM sees every byte once.  Only the first and last
// few records of a path (unaligned ends) are written one by one.
__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d)
{ asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory"); }
__device__ __forceinline__ double cellWord(int m) { return __longlong_as_double((long long)(unsigned)m); }    // {int m; int reserved = 0}

// The rings are lane-interleaved in shared memory -- entry q of lane l lives at [q][l] -- so that every access of a
// warp is bank-conflict free whatever ring positions its lanes are at: per warp ds[12][32], rho[12][32] (f64) and
// m[12][32] (i32).
// 1: a group of four records (160 contiguous bytes) is assembled in shared memory and leaves as ONE bulk copy of the async proxy
// (cp.async.bulk.global.shared::cta, SASS UBLKCP) instead of five 256-bit stores through the load/store pipe
#ifndef SKG_FILL_BULK
#define SKG_FILL_BULK 0
#endif
__device__ __forceinline__ void sts2F64(unsigned a, double v0, double v1) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v0), "d"(v1) : "memory"); }
__device__ __forceinline__ void bulkStore(void* gdst, unsigned ssrc, unsigned bytes)
{ asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsyncShared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct RecordJobStaged : RayJobBase
{
    static constexpr int kBatches = 1;              // the ring holds three periods
    static constexpr bool kCartRegBorders = true, kCartTinySelect = false; static constexpr bool kTreeHints = SKG_TREE_HINTS_PATH, kCartRhoAhead = true;       // the record kernel is bound by the load/store pipe
    static constexpr unsigned RHO_OFF = SKG_RING * 32 * 8, M_OFF = 2 * SKG_RING * 32 * 8;
    static constexpr size_t bytesPerWarp() { return (size_t)SKG_RING * 32 * (8 + 8 + 4); }
    static constexpr size_t bytesPerWarpAll() { return bytesPerWarp() + 32 * 4 + (SKG_FILL_BULK ? 32 * 160 : 0); }      // + the ray index of every lane (+ a 160-byte group)
    unsigned stage;             // SKG_FILL_BULK: shared-window address of this lane's 160-byte group

    const int64_t* offsets; const int* ell; int ellStride; Medium med;
    skg_segment* seg;
    int* lengths;               // one-pass mode: the number of records actually written per ray (offsets then hold capacities), or null
    unsigned rbItem;            // shared-window address of this lane's slot for the ray index (one-pass mode; not a register over the walk)
    unsigned rb, rbM;           // shared-window addresses of this lane's columns: ds entry q at rb + 256 q, rho at + RHO_OFF; m entry q at rbM + 128 q
    double* out0;               // record with relative index 0: &seg[first record of the path rounded down to a multiple of 4]
    KappaRho kr; double kext0; double sacc, tacc; bool optical, async;
    int o, f, ready;            // relative record indices: parked up to o, written up to f, densities landed up to ready
    int qo, qf;                 // ring positions (256 x slot) of o and f; relative index mod 4 == record index mod 4 == slot mod 4

    __device__ __forceinline__ void bind(char* warpBase)
    {
        const int lane = threadIdx.x & 31;
        const unsigned w = (unsigned)__cvta_generic_to_shared(warpBase);
        rb = w + 8u * lane; rbM = w + M_OFF + 4u * lane; rbItem = w + (unsigned)bytesPerWarp() + 4u * lane;
        stage = w + (unsigned)bytesPerWarp() + 128u + 160u * lane;
        o = f = ready = 0; sacc = tacc = 0; optical = async = false; kext0 = 0; qo = qf = 0; out0 = nullptr;
    }

    __device__ __forceinline__ int begin(int i)
    {
        loadRay(i);
        const int64_t first = offsets[i];
        if (lengths) stsI32(rbItem, i);
        const int a0 = (int)(first & 3);
        out0 = reinterpret_cast<double*>(seg + (first - a0));
        o = f = ready = a0; qo = qf = 256 * a0;
        sacc = 0; tacc = 0;
        optical = ell != nullptr;
        int l = optical ? ell[(size_t)i * ellStride] : 0;
        kr = KappaRho{med.rho, med.kext + l, med.Ncomp, med.Nlambda};
        async = optical && med.Ncomp == 1;
        kext0 = async ? __ldg(med.kext + l) : 0.0;
        return 1;
    }
    __device__ __forceinline__ void advance() { o++; qo = qo + 256 == 256 * SKG_RING ? 0 : qo + 256; }
    __device__ __forceinline__ bool outside(double d)
    { stsI32(rbM + (qo >> 1), -1); stsF64(rb + qo, d); stsF64(rb + RHO_OFF + qo, 0.0); advance(); return true; }     // rho(-1,h) = 0 (DustSystem.cpp:918-921)
    __device__ __forceinline__ bool segment(int mm, double d)
    {
        stsI32(rbM + (qo >> 1), mm); stsF64(rb + qo, d);
        if (async) asyncCopy8(rb + RHO_OFF + qo, med.rho + mm);
        advance();
        return true;
    }
    // KappaRho (DustSystem.cpp:465-491) of a ring entry: one component -> 0 + kext*rho == kext*rho exactly
    __device__ __forceinline__ double kapparho(unsigned q, int mm) const { return async ? kext0 * ldsVF64(rb + RHO_OFF + q) : (optical ? kr(mm) : 0.0); }

    __device__ __forceinline__ void emitOne()
    {
        const unsigned q = qf;
        const int mm = ldsVI32(rbM + (q >> 1)); const double d = ldsVF64(rb + q);
        const double dt = kapparho(q, mm) * d;
        sacc += d; tacc += dt;
        double* out = out0 + 5 * (size_t)f;
        out[0] = cellWord(mm); out[1] = d; out[2] = sacc; out[3] = dt; out[4] = tacc;
        f++; qf = q + 256 == 256 * SKG_RING ? 0 : q + 256;
    }
    // four records at a 4-aligned record index: 160 contiguous, 32-byte aligned bytes as five 256-bit stores
    __device__ __forceinline__ void emitFour()
    {
        const unsigned q = qf;              // a multiple of 4 slots: no wrap inside the group
        const unsigned qm = rbM + (q >> 1), qd = rb + q;
        const int m0 = ldsVI32(qm), m1 = ldsVI32(qm + 128), m2 = ldsVI32(qm + 256), m3 = ldsVI32(qm + 384);
        const double d0 = ldsVF64(qd), d1 = ldsVF64(qd + 256), d2 = ldsVF64(qd + 512), d3 = ldsVF64(qd + 768);
        const double k0 = kapparho(q, m0), k1 = kapparho(q + 256, m1), k2 = kapparho(q + 512, m2), k3 = kapparho(q + 768, m3);
        const double s0 = sacc + d0, s1 = s0 + d1, s2 = s1 + d2, s3 = s2 + d3;
        const double t0 = k0 * d0, t1 = k1 * d1, t2 = k2 * d2, t3 = k3 * d3;
        const double a0 = tacc + t0, a1 = a0 + t1, a2 = a1 + t2, a3 = a2 + t3;
        sacc = s3; tacc = a3;
        double* out = out0 + 5 * (size_t)f;      // 4 records x 5 words
#if SKG_FILL_BULK
        bulkWaitRead();                          // the previous group has left this lane's staging bytes
        sts2F64(stage, cellWord(m0), d0); sts2F64(stage + 16, s0, t0); sts2F64(stage + 32, a0, cellWord(m1)); sts2F64(stage + 48, d1, s1);
        sts2F64(stage + 64, t1, a1); sts2F64(stage + 80, cellWord(m2), d2); sts2F64(stage + 96, s2, t2); sts2F64(stage + 112, a2, cellWord(m3));
        sts2F64(stage + 128, d3, s3); sts2F64(stage + 144, t3, a3);
        fenceProxyAsyncShared();
        bulkStore(out, stage, 160u); bulkCommit();
        f += 4; qf = q + 1024 == 256 * SKG_RING ? 0 : q + 1024;
        return;
#endif
        store4(out, cellWord(m0), d0, s0, t0);
        store4(out + 4, a0, cellWord(m1), d1, s1);
        store4(out + 8, t1, a1, cellWord(m2), d2);
        store4(out + 12, s2, t2, a2, cellWord(m3));
        store4(out + 16, d3, s3, t3, a3);
        f += 4; qf = q + 1024 == 256 * SKG_RING ? 0 : q + 1024;
    }
    // turns the parked entries [f, upto) into records, in path order
    __device__ __forceinline__ void emit(int upto)
    {
        while (f < upto && (f & 3)) emitOne();          // unaligned head of a path
        while (f + 4 <= upto) emitFour();
        while (f < upto) emitOne();                     // tail (finish only)
    }
    // entries parked before the previous call have their density in the ring by now
    __device__ __forceinline__ void periodic() { asyncCommit(); asyncWaitAllButLatest(); emit(ready & ~3); ready = o; }
    __device__ __forceinline__ void finish()
    {
        asyncCommit(); asyncWaitAll(); emit(o); ready = o;
#if SKG_FILL_BULK
        bulkWaitRead();
#endif
        if (lengths)
        {
            // one-pass mode: report the length; the ray's slab [offsets[i], offsets[i+1]) must have held it
            const int i = ldsVI32(rbItem);
            const int64_t first = offsets[i], cap = offsets[i + 1] - first;
            const int len = o - (int)(first & 3);
            lengths[i] = len;
            if (len > cap) atomicAdd(overflow, 1ull);
        }
    }
    unsigned long long* overflow;
    __device__ __forceinline__ void collective(bool) {}
};

// 1: the record kernel really pulls the next cell's density into L1 at the end of a crossing.  Measured slower (2.94 ms against
// 2.80 ms for 4 Mi rays), so the table pointer stays null; the guarded prefetch is still compiled into the walker
// (kCartRhoAhead) because with that branch in the crossing ptxas allocates the 96 registers without spilling (2.58 ms).
#ifndef SKG_FILL_RHO_AHEAD
#define SKG_FILL_RHO_AHEAD 0
#endif
template<int KIND>
__global__ void __launch_bounds__(128) pathFillKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem, int refill,
                                                      int n, const double* __restrict__ r, const double* __restrict__ k,
                                                      const int* __restrict__ ell, int ellStride, const int64_t* __restrict__ offsets,
                                                      skg_segment* __restrict__ segments, int* work, int* __restrict__ lengths)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART)
    {
        cart = stageCart(G.cart, smem, cartSmem);
#if SKG_FILL_RHO_AHEAD
        if (ell && med.Ncomp == 1) { cart.rhoAhead = med.rho; cart.rhoAheadStride = 1; }     // the density the ring's copy fetches a crossing later
#endif
    }
    size_t skip = (KIND == GRID_CART && cartSmem) ? SKG_CART_SMEM_DOUBLES(G.cart) : 0;
    RecordJobStaged job; job.r = r; job.k = k; job.offsets = offsets; job.ell = ell; job.ellStride = ellStride; job.med = med;
    job.seg = segments; job.lengths = lengths; job.overflow = &ctr->errors;
    job.bind(reinterpret_cast<char*>(smem + ((skip + 1) & ~(size_t)1)) + (threadIdx.x >> 5) * RecordJobStaged::bytesPerWarpAll());      // 16-byte aligned
    runJobs<KIND>(G, cart, ctr, job, n, work, refill);
}

template<int KIND, bool FAST>
__global__ void __launch_bounds__(128) opticalDepthKernel(const __grid_constant__ GridSet G, const Medium med, Counters* ctr, bool cartSmem, int refill,
                                                          int n, const double* __restrict__ r, const double* __restrict__ k,
                                                          const int* __restrict__ ell, int ellStride,
                                                          const double* __restrict__ dist, double* __restrict__ tau, int* work)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART) cart = stageCart(G.cart, smem, cartSmem);
    TauJobT<FAST> job; job.r = r; job.k = k; job.ell = ell; job.ellStride = ellStride; job.med = med; job.dist = dist; job.out = tau;
    runJobs<KIND>(G, cart, ctr, job, n, work, refill);
}

// One-pass batched path() on Cartesian grids: an upper bound of every ray's number of segments WITHOUT walking it.  The
// path of CartesianDustGrid::path is monotone in the cell indices of every axis and ends when one of them leaves the grid,
// so it has (number of crossings) = |di| + |dj| + |dk| + 1 segments between its first cell and the cell in which the
// straight line leaves the grid box, plus the up to three "outside" segments of the entry (cartEnter).  The walk itself
// rounds differently from this closed form only at the level of an ulp of the coordinates: near a cell corner it may take
// the crossings in another order or end one cell over in an axis, hence one spare segment per axis.  Segments with
// ds <= 0 are dropped by addSegment (DustGridPath.cpp:46-53): the bound stays a bound.  The record kernel reports the
// true length and flags any ray whose slab was too short (never observed; skg_path_batch then returns an error).
__global__ void __launch_bounds__(128) pathCapacityKernel(const CartGrid g, int64_t n, const double* __restrict__ r, const double* __restrict__ k,
                                                          int* __restrict__ cap)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    {
        double x = r[3 * q], y = r[3 * q + 1], z = r[3 * q + 2];
        const double kx = k[3 * q], ky = k[3 * q + 1], kz = k[3 * q + 2];
        Entry en; int i, j, kk;
        if (!cartEnter(g, x, y, z, kx, ky, kz, en, i, j, kk)) { cap[q] = 0; continue; }
        int c = 0;
        for (int u = 0; u < en.n; u++) if (en.ds[u] > 0) c++;
        // where the straight line leaves the box
        const double tx = fabs(kx) > 1e-15 ? ((kx < 0 ? g.ext[0] : g.ext[1]) - x) / kx : SKG_DBL_MAX;
        const double ty = fabs(ky) > 1e-15 ? ((ky < 0 ? g.ext[2] : g.ext[3]) - y) / ky : SKG_DBL_MAX;
        const double tz = fabs(kz) > 1e-15 ? ((kz < 0 ? g.ext[4] : g.ext[5]) - z) / kz : SKG_DBL_MAX;
        const double t = fmin(tx, fmin(ty, tz));
        const int i2 = locateClip(g.xv, x + t * kx, g.Nx + 1), j2 = locateClip(g.yv, y + t * ky, g.Ny + 1), k2 = locateClip(g.zv, z + t * kz, g.Nz + 1);
        c += abs(i2 - i) + abs(j2 - j) + abs(k2 - kk) + 1 + 3;
        cap[q] = c;
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
        else if (KIND == GRID_SYM) res = symWhichCell(G.sym, x, y, z);
        else res = voroCellIndex(G.voro, x, y, z);
        m[i] = res;
    }
}

// ---------------------------------------------------------------------------------------------------
static GridSet gridSet(const Engine& e) { GridSet G; G.cart = e.cart; G.tree = e.tree; G.amesh = e.amesh; G.voro = e.voro; G.sym = e.sym; return G; }

struct LaunchCfg { int blocks; size_t smem; bool cartSmem; int* work; int refill; };
static LaunchCfg cfgFor(Engine& e, int64_t n)
{
    LaunchCfg c;
    int64_t want = (n + 127) / 128;
    int64_t cap = (int64_t)e.smCount * 16;
    if (const char* v = getenv("SKG_PATH_BLOCKS_PER_SM")) cap = (int64_t)e.smCount * std::max(1, std::min(64, atoi(v)));
    c.blocks = (int)std::max<int64_t>(1, std::min(want, cap));
    c.smem = 0; c.cartSmem = false;
    // idle lanes of a warp at which it draws new rays (runJobs): entry + point location are cheap on the Cartesian grid,
    // a chain of dependent reads on the hierarchical ones
    c.refill = e.gridKind == GRID_CART ? 12 : 8; if (const char* v = getenv("SKG_PATH_REFILL")) c.refill = std::max(1, std::min(32, atoi(v)));
    if (n > 2147483647LL) throw Error("at most 2^31-1 rays per call");
    e.scratchWork.ensure(sizeof(int)); c.work = e.scratchWork.as<int>();
    SKG_CUDA(cudaMemsetAsync(c.work, 0, sizeof(int), e.stream));
    if (e.gridKind == GRID_CART)
    {
        // the Cartesian walker reads the borders from shared memory only (CartWalker::step)
        size_t need = sizeof(double) * SKG_CART_SMEM_DOUBLES(e.cart);
        if (need > SKG_CART_SMEM_MAX) throw Error("CartesianDustGrid: more than 8189 mesh borders in total are not supported");
        c.smem = need; c.cartSmem = true;
        if (!e.attrPath)
        {
            SKG_CUDA(cudaFuncSetAttribute(pathCountKernel<GRID_CART>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            SKG_CUDA(cudaFuncSetAttribute(opticalDepthKernel<GRID_CART, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            SKG_CUDA(cudaFuncSetAttribute(opticalDepthKernel<GRID_CART, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            e.attrPath = true;
        }
    }
    return c;
}

#define SKG_DISPATCH(e, CALL) \
    switch ((e).gridKind) { \
    case GRID_CART:  { constexpr int K = GRID_CART;  CALL; break; } \
    case GRID_TREE:  { constexpr int K = GRID_TREE;  CALL; break; } \
    case GRID_AMESH: { constexpr int K = GRID_AMESH; CALL; break; } \
    case GRID_VORO:  { constexpr int K = GRID_VORO;  CALL; break; } \
    case GRID_SYM:   { constexpr int K = GRID_SYM;   CALL; break; } \
    default: throw Error("no dust grid has been set"); }

void launchPathCount(Engine& e, int64_t n, const double* d_r, const double* d_k, int* d_counts)
{
    if (n <= 0) return;
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (pathCountKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.ctr(), c.cartSmem, c.refill, (int)n, d_r, d_k, d_counts, c.work)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchPathFill(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                    const int64_t* d_offsets, skg_segment* d_segments, int* d_lengths)
{
    if (n <= 0) return;
    if (d_ell && !e.med.rho) throw Error("skg_path_fill with wavelength indices needs skg_medium first");
    if (d_ell && e.med.Ncells != e.Ncells) throw Error("the medium has " + std::to_string(e.med.Ncells) + " cells but the grid has " + std::to_string(e.Ncells) + ": call skg_medium again");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    c.smem += 4 * RecordJobStaged::bytesPerWarpAll() + 8;
    if (const char* pad = getenv("SKG_FILL_SMEM_PAD")) c.smem += (size_t)atoi(pad);      // experiment: limits resident CTAs
    if (!e.attrFill)
    {
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_CART>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_TREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_AMESH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_VORO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        if (const char* cv = getenv("SKG_FILL_CARVEOUT")) SKG_CUDA(cudaFuncSetAttribute(pathFillKernel<GRID_CART>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
        e.attrFill = true;
    }
    SKG_DISPATCH(e, (pathFillKernel<K><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, c.refill, (int)n, d_r, d_k, d_ell, ellStride,
                                                                                d_offsets, d_segments, c.work, d_lengths)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchOpticalDepth(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                        const double* d_dist, double* d_tau, bool mcWalker)
{
    if (n <= 0) return;
    if (!e.med.rho) throw Error("skg_opticaldepth needs skg_medium first");
    if (e.med.Ncells != e.Ncells) throw Error("the medium has " + std::to_string(e.med.Ncells) + " cells but the grid has " + std::to_string(e.Ncells) + ": call skg_medium again");
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    if (mcWalker)
    { SKG_DISPATCH(e, (opticalDepthKernel<K, true><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, c.refill, (int)n, d_r, d_k, d_ell, ellStride, d_dist, d_tau, c.work))); }
    else
    { SKG_DISPATCH(e, (opticalDepthKernel<K, false><<<c.blocks, 128, c.smem, e.stream>>>(G, e.med, e.ctr(), c.cartSmem, c.refill, (int)n, d_r, d_k, d_ell, ellStride, d_dist, d_tau, c.work))); }
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchPathCapacity(Engine& e, int64_t n, const double* d_r, const double* d_k, int* d_cap)
{
    if (n <= 0) return;
    if (e.gridKind != GRID_CART) throw Error("analytic path capacities exist for Cartesian grids only");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)e.smCount * 16));
    pathCapacityKernel<<<blocks, 128, 0, e.stream>>>(e.cart, n, d_r, d_k, d_cap);
    e.launches++; SKG_CUDA(cudaGetLastError());
}

void launchWhichCell(Engine& e, int64_t n, const double* d_r, int* d_m)
{
    if (n <= 0) return;
    LaunchCfg c = cfgFor(e, n); GridSet G = gridSet(e);
    SKG_DISPATCH(e, (whichCellKernel<K><<<c.blocks, 128, 0, e.stream>>>(G, e.ctr(), n, d_r, d_m)));
    e.launches++; SKG_CUDA(cudaGetLastError());
}

// compares divInvariant with the IEEE division on pseudo-random operands spanning the magnitudes the walkers see
// (numerators: differences of coordinates down to a few ulp, up to the domain size; divisors: direction cosines down
// to the 1e-15 cut) plus adversarial mantissas; counts the quotients that differ in any bit
__global__ void divisionSelfTest(unsigned long long n, unsigned long long seed, unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    {
        // splitmix64
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
        auto next = [&]() { z += 0x9E3779B97F4A7C15ull; unsigned long long t = z; t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull; t = (t ^ (t >> 27)) * 0x94D049BB133111EBull; return t ^ (t >> 31); };
        unsigned long long ma = next(), mb = next(), ex = next();
        // mantissas: random, or with long runs of ones / zeros (the hard cases of division rounding)
        if ((ex & 7) == 0) ma |= 0x000FFFFFFFFFF000ull; if ((ex & 7) == 1) mb |= 0x000FFFFFFFFFFF00ull;
        if ((ex & 7) == 2) ma &= 0xFFF0000000000FFFull; if ((ex & 7) == 3) mb &= 0xFFF00000000000FFull;
        int ea = (int)((ex >> 8) % 140) - 70;        // |a| in 2^-70 .. 2^70 times the scale below
        int eb = -(int)((ex >> 20) % 50);            // |b| in 2^-50 .. 1
        double a = __longlong_as_double((long long)((ma & 0x800FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
        double b = __longlong_as_double((long long)((mb & 0x800FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
        a = ldexp(a, ea) * 3.0e16; b = ldexp(b, eb);
        if (fabs(b) <= 1e-15) b = b < 0 ? -1e-15 * 1.0000001 : 1e-15 * 1.0000001;
        double rb = 1.0 / b;
        double q1 = a / b, q2 = divInvariant(a, b, rb);
        if (__double_as_longlong(q1) != __double_as_longlong(q2)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

unsigned long long runDivisionSelfTest(Engine& e, unsigned long long n, unsigned long long seed)
{
    e.scratchWork.ensure(sizeof(unsigned long long));
    SKG_CUDA(cudaMemsetAsync(e.scratchWork.p, 0, sizeof(unsigned long long), e.stream));
    divisionSelfTest<<<e.smCount * 16, 256, 0, e.stream>>>(n, seed, e.scratchWork.as<unsigned long long>());
    e.launches++; SKG_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    SKG_CUDA(cudaMemcpyAsync(&bad, e.scratchWork.p, sizeof(bad), cudaMemcpyDeviceToHost, e.stream));
    e.sync();
    return bad;
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
