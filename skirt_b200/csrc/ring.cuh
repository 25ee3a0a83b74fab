// Per-lane rings in shared memory that decouple a crossing from the use of the crossed cell's density.
//
// The density gather rho[m] of a crossing is a dependent, L2-latency read (the cell is only known once the crossing
// has been computed), and with fp64 register budgets that leave 4-5 warps per scheduler that latency cannot be
// hidden by other warps alone.  So a crossing only parks (m, ds) in a small ring and starts an asynchronous copy
// (LDGSTS) of the cell's density into the same ring slot; every SKG_PERIOD crossings the lane consumes the entries
// parked one period earlier -- whose densities have landed meanwhile -- in path order, so that all running sums
// (optical depth, absorbed luminosity) are formed in exactly the same order as without the ring.
//
// The rings are lane-interleaved -- entry q of lane l lives at [q][l] -- so that every access of a warp is free of
// bank conflicts whatever ring positions its lanes are at: per warp ds[12][32], rho[12][32] (f64) and m[12][32] (i32).
#pragma once

namespace skg
{

#define SKG_RING 12         // three periods of four: at most 11 entries are parked at any time

// asynchronous 8-byte copy global -> shared (LDGSTS)
__device__ __forceinline__ void asyncCopy8(unsigned smemDst, const double* src)
{ asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smemDst), "l"(src) : "memory"); }
__device__ __forceinline__ void asyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void asyncWaitAllButLatest() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void asyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stsF64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void stsI32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ double ldsVF64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int ldsVI32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

struct RhoRing
{
    static constexpr unsigned RHO_OFF = SKG_RING * 32 * 8, M_OFF = 2 * SKG_RING * 32 * 8;
    static constexpr unsigned WRAP = 256 * SKG_RING;
    static constexpr size_t bytesPerWarp() { return (size_t)SKG_RING * 32 * (8 + 8 + 4); }

    unsigned rb, rbM;           // shared-window addresses of this lane's columns: ds entry q at rb + 256 q, rho at + RHO_OFF; m entry q at rbM + 128 q
    int o, f, ready;            // entry counters: parked up to o, consumed up to f, densities landed up to ready
    unsigned qo, qf;            // ring positions (256 x slot) of o and f

    __device__ __forceinline__ void bind(char* warpBase)
    {
        const int lane = threadIdx.x & 31;
        const unsigned w = (unsigned)__cvta_generic_to_shared(warpBase);
        rb = w + 8u * lane; rbM = w + M_OFF + 4u * lane;
        o = f = ready = 0; qo = qf = 0;
    }
    // parks a crossing and starts the copy of its density
    __device__ __forceinline__ void park(int m, double ds, const double* rho)
    {
        stsI32(rbM + (qo >> 1), m); stsF64(rb + qo, ds);
        asyncCopy8(rb + RHO_OFF + qo, rho);
        o++; qo = qo + 256 == WRAP ? 0 : qo + 256;
    }
    // consumes the entries [f, upto) in path order: use(m, ds, rho)
    template<class F> __device__ __forceinline__ void drain(int upto, F&& use)
    {
        while (f < upto)
        {
            const int m = ldsVI32(rbM + (qf >> 1)); const double ds = ldsVF64(rb + qf); const double rho = ldsVF64(rb + RHO_OFF + qf);
            use(m, ds, rho);
            f++; qf = qf + 256 == WRAP ? 0 : qf + 256;
        }
    }
    // every SKG_PERIOD crossings (warp-uniform): the entries parked before the previous call have landed by now
    template<class F> __device__ __forceinline__ void periodic(F&& use) { asyncCommit(); asyncWaitAllButLatest(); drain(ready, use); ready = o; }
    // end of a path
    template<class F> __device__ __forceinline__ void finish(F&& use) { asyncCommit(); asyncWaitAll(); drain(o, use); ready = o; }
};

}   // namespace skg
