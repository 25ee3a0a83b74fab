// Warp-level scheduling of ray jobs.
//
// Every traversal kernel of the engine (batched DustGrid::path(), optical depth, and the peel-off / absorb /
// propagate stages of the photon shooter) is "for each item: set up a ray, walk it through the grid feeding
// segments to a sink, finish".  Path lengths vary from 1 to a few hundred crossings, so a plain
// one-item-per-thread loop leaves most lanes of a warp idle while the longest path finishes.  runJobs() keeps
// the lanes busy instead: a lane whose path has ended parks until `refill` lanes of its warp are parked, then the
// parked lanes finish their items together and draw new ones from a global work counter.  The crossing step
// itself is executed by all walking lanes in lock step.  Every grid type provides a stepping walker (geom.cuh).
//
// A Job provides (per lane, state in registers):
//   int  begin(int item)         set rx..dz for the item; 0 = nothing to do, 1 = walk the ray, 2 = no walk but finish()
//   bool outside(double ds)      segment outside the grid (m = -1); false stops the walk
//   bool segment(int m, double ds)
//   bool segmentU<U>(int m, double ds)   the same for walkers that step in branch-free batches; U = position of the crossing in its
//                                batch of SKG_PERIOD (compile time: lets a job keep one pending entry per position in registers)
//   void idleU<U>()              a walking lane's crossing at position U that produced no segment
//   int  cellHint()              where the ray starts, if known from an earlier traversal (the walker's locator: tree / adaptive-mesh
//                                node, Voronoi cell), else -1: the walker then skips its point location
//   void noteStart(int locator)  called after a start without hint: the job may remember the locator for later traversals
//   void finish()                called once per item that returned 1 or 2 from begin()
//   void collective(bool fin)    called warp-uniformly after the finish() calls; fin = this lane just finished an item
//   void periodic()              called warp-uniformly after every SKG_PERIOD crossing steps
//   static constexpr bool kCartRegBorders, kCartTinySelect   variant of the Cartesian walker (geom.cuh)
//   static constexpr bool kCartFast         the shooting stages' t-parameterised Cartesian walker instead of the bit-exact one
//   static constexpr bool kCartRhoAhead     Cartesian walker: pull the next cell's density into L1 at the end of a crossing
//   static constexpr bool kTreeHints        tree walker: wall-bin table for walls with several neighbours (geom.cuh)
//   static constexpr int kBatches           batches of SKG_PERIOD crossings between two votes / periodic() calls
#pragma once
#include <type_traits>
#include "geom.cuh"

namespace skg
{

#define SKG_PERIOD 4

// shared-memory staging of the Cartesian borders; returns a CartGrid view whose xv/yv/zv point to smem
__device__ __forceinline__ CartGrid stageCart(const CartGrid& g, double* smem, bool useSmem)
{
    if (!useSmem) return g;
    CartGrid s = g;
    // layout: [pad] xv[0..Nx] [pad] [pad] yv[0..Ny] [pad] [pad] zv[0..Nz] [pad]   (SKG_CART_SMEM_DOUBLES)
    int nx = g.Nx + 1, ny = g.Ny + 1, nz = g.Nz + 1;
    double* sxv = smem + 1; double* syv = sxv + nx + 2; double* szv = syv + ny + 2;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sxv[i] = g.xv[i];
    for (int i = threadIdx.x; i < ny; i += blockDim.x) syv[i] = g.yv[i];
    for (int i = threadIdx.x; i < nz; i += blockDim.x) szv[i] = g.zv[i];
    if (threadIdx.x == 0) { sxv[-1] = sxv[nx] = syv[-1] = syv[ny] = szv[-1] = szv[nz] = 0.0; }
    __syncthreads();
    s.xv = sxv; s.yv = syv; s.zv = szv;
    s.sx = (unsigned)__cvta_generic_to_shared(sxv); s.sy = (unsigned)__cvta_generic_to_shared(syv); s.sz = (unsigned)__cvta_generic_to_shared(szv);
    s.staged = 1;
    return s;
}


template<class Walker, class GridT, class Job>
__device__ __forceinline__ void runJobsStep(const GridT& grid, Counters* ctr, Job& job, int n, int* workCounter, int refill)
{
    constexpr int kStepUnroll = Walker::kStepUnroll;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int state = 0;              // 0 idle, 1 walking, 2 walk ended (finish pending)
    bool more = true;
    Walker w; w.alive = false;
    while (true)
    {
        unsigned walking = __ballot_sync(FULL, state == 1);
        if (__popc(~walking) >= refill || walking == 0)
        {
            if (state == 2) job.finish();
            job.collective(state == 2);
            if (state == 2) state = 0;
            if (more)
            {
                const unsigned idle = __ballot_sync(FULL, state == 0);
                int base = 0;
                if (lane == 0) base = atomicAdd(workCounter, __popc(idle));
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) more = false;
                else if (state == 0)
                {
                    const int idx = base + __popc(idle & ((1u << lane) - 1));
                    if (idx < n)
                    {
                        const int b = job.begin(idx);
                        if (b == 1)
                        {
                            Entry en;
                            state = 2;
                            const int hint = job.cellHint();
                            if (w.start(grid, ctr, job.rx, job.ry, job.rz, job.dx, job.dy, job.dz, en, hint))
                            {
                                // (a ray that had to be moved into the grid starts OUTSIDE: its locator is not where the packet lies)
                                if (hint < 0 && en.n == 0) job.noteStart(w.locator());
                                bool cont = true;
                                for (int q = 0; q < en.n && cont; q++) if (en.ds[q] > 0) cont = job.outside(en.ds[q]);
                                if (cont) state = 1;
                            }
                        }
                        else if (b == 2) state = 2;
                    }
                }
            }
            walking = __ballot_sync(FULL, state == 1);
            if (walking == 0)
            {
                if (!more && __ballot_sync(FULL, state == 2) == 0) break;
                continue;
            }
        }
        // kBatches x SKG_PERIOD crossings between two warp-wide votes: a vote is a convergence point at which every outstanding
        // load of the warp is waited for, so voting once per crossing would expose the latency of each density gather
        // (lanes whose path ends inside the batch idle for at most SKG_PERIOD - 1 crossings)
        for (int batch = 0; batch < Job::kBatches; batch++)
        {
            if constexpr (Walker::kPredicated)
            {
                // branch-free crossings: lanes that are not walking execute the step predicated off
                auto crossing = [&](auto U)
                {
                    bool live = state == 1;
                    int m; double ds;
                    const bool seg = w.stepLive(grid, live, m, ds);
                    if (seg) { if (!job.template segmentU<decltype(U)::value>(m, ds)) live = false; }
                    else if (state == 1) job.template idleU<decltype(U)::value>();      // a crossing without a segment (ds <= 0)
                    if (state == 1 && !live) state = 2;
                };
                static_assert(SKG_PERIOD == 4, "the batch of crossings is written out for a period of four");
                crossing(std::integral_constant<int, 0>{}); crossing(std::integral_constant<int, 1>{});
                crossing(std::integral_constant<int, 2>{}); crossing(std::integral_constant<int, 3>{});
            }
            else
            {
#pragma unroll kStepUnroll
                for (int u = 0; u < SKG_PERIOD; u++)
                {
                    if (state == 1)
                    {
                        int m; double ds;
                        const bool seg = w.step(grid, ctr, m, ds);
                        const bool cont = seg ? job.segment(m, ds) : true;
                        // walkers that split a crossing: the read of the next node's record, started by step(), has been
                        // travelling while the job worked on the segment (a walk the job ends needs no next node)
                        if constexpr (Walker::kSplitStep) { if (cont) w.resolve(grid, ctr); }
                        if (!cont || !w.alive) state = 2;
                    }
                }
            }
        }
        job.periodic();
    }
}

template<int KIND, class Job, class Grids>
__device__ __forceinline__ void runJobs(const Grids& G, const CartGrid& cart, Counters* ctr, Job& job, int n, int* workCounter, int refill = 8)
{
    if (KIND == GRID_CART)
    {
        if constexpr (Job::kCartFast)
        {
            // (warp-uniform: a property of the grid)
            if (cart.uniform) runJobsStep<CartFastWalkerT<true>>(cart, ctr, job, n, workCounter, refill);
            else runJobsStep<CartFastWalkerT<false>>(cart, ctr, job, n, workCounter, refill);
        }
        else runJobsStep<CartWalkerT<Job::kCartRegBorders, Job::kCartTinySelect, Job::kCartRhoAhead>>(cart, ctr, job, n, workCounter, refill);
    }
    else if (KIND == GRID_TREE) runJobsStep<TreeWalkerT<Job::kTreeHints>>(G.tree, ctr, job, n, workCounter, refill);
    else if (KIND == GRID_AMESH) runJobsStep<AMeshWalker>(G.amesh, ctr, job, n, workCounter, refill);
    else if (KIND == GRID_SYM) runJobsStep<SymWalker>(G.sym, ctr, job, n, workCounter, refill);
#ifdef SKG_VORO_MC_EXACT
    else runJobsStep<VoroWalkerT<true>>(G.voro, ctr, job, n, workCounter, refill);                   // experiment: exact walker everywhere
#else
    else runJobsStep<VoroWalkerT<!Job::kCartFast>>(G.voro, ctr, job, n, workCounter, refill);      // kCartFast marks the shooting stages
#endif
}

}   // namespace skg
