// Multi-GPU reduction: replaces Instrument::sumResults (Instrument.cpp:57-65) and PanDustSystem::sumResults
// (PanDustSystem.cpp:394-404), i.e. PeerToPeerCommunicator::sum_all over MPI, by NCCL all-reduce over NVLink.
#include <cstring>
#include <dlfcn.h>
#include "engine.h"

namespace skg
{

// ---- NCCL (loaded at run time: libnccl.so.2 is already in the process when torch.distributed is) ---------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int (*fnGetUniqueId)(ncclUniqueId*);
typedef int (*fnCommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*fnAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fnBroadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fnCommDestroy)(ncclComm_t);
typedef int (*fnGroup)(void);
typedef const char* (*fnErr)(int);
static struct { void* lib = nullptr; fnGetUniqueId getUniqueId; fnCommInitRank commInitRank; fnAllReduce allReduce; fnBroadcast broadcast;
                fnCommDestroy commDestroy; fnGroup groupStart, groupEnd; fnErr errString; } nccl;

static void loadNccl()
{
    if (nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (nccl.lib) break; }
    if (!nccl.lib) throw Error(std::string("cannot load NCCL: ") + dlerror());
    nccl.getUniqueId = (fnGetUniqueId)dlsym(nccl.lib, "ncclGetUniqueId");
    nccl.commInitRank = (fnCommInitRank)dlsym(nccl.lib, "ncclCommInitRank");
    nccl.allReduce = (fnAllReduce)dlsym(nccl.lib, "ncclAllReduce");
    nccl.broadcast = (fnBroadcast)dlsym(nccl.lib, "ncclBroadcast");
    nccl.commDestroy = (fnCommDestroy)dlsym(nccl.lib, "ncclCommDestroy");
    nccl.groupStart = (fnGroup)dlsym(nccl.lib, "ncclGroupStart");
    nccl.groupEnd = (fnGroup)dlsym(nccl.lib, "ncclGroupEnd");
    nccl.errString = (fnErr)dlsym(nccl.lib, "ncclGetErrorString");
    if (!nccl.getUniqueId || !nccl.commInitRank || !nccl.allReduce || !nccl.broadcast || !nccl.commDestroy || !nccl.groupStart || !nccl.groupEnd)
    { nccl.lib = nullptr; throw Error("NCCL symbols missing"); }
}
#define SKG_NCCL(call) do { int rc__ = (call); if (rc__ != 0) throw skg::Error(std::string(#call) + ": " + (nccl.errString ? nccl.errString(rc__) : "NCCL error")); } while (0)


void destroyComm(Engine& e)
{
    if (e.nccl && nccl.commDestroy) nccl.commDestroy((ncclComm_t)e.nccl);
    e.nccl = nullptr; e.rank = 0; e.nranks = 1;
}

static Engine& engineOf(skg_engine* eh)
{
    if (!eh) throw Error("null engine");
    Engine& e = *reinterpret_cast<Engine*>(eh);
    SKG_CUDA(cudaSetDevice(e.device));
    return e;
}

static const int ncclDouble = 8, ncclSum = 0;       // nccl.h: ncclFloat64 = 8, ncclSum = 0

// in-place sum over the ranks of the selected accumulators, each according to its state (engine.h)
static void allreduce(Engine& e, int which, double* elapsedMs)
{
    if (elapsedMs) *elapsedMs = 0;
    if (which & ~SKG_REDUCE_ALL) throw Error("skg_allreduce: unknown accumulator selection");
    if (!e.nccl || e.nranks <= 1) return;
    struct Sel { int bit; int* state; const char* name; };
    const Sel sel[] = {{SKG_REDUCE_LABS_STELLAR, &e.accLabs, "the stellar absorption table"}, {SKG_REDUCE_LABS_DUST, &e.accLabsDust, "the dust absorption table"},
                       {SKG_REDUCE_INSTRUMENTS, &e.accInstr, "the detector arrays"}};
    int todo = 0;
    for (const Sel& s : sel)
    {
        if (!(which & s.bit)) continue;
        if (*s.state == Engine::ACC_MIXED)
            throw Error(std::string(s.name) + " received rank-local additions after being summed over the processes: they cannot be summed again in place "
                        "(the reference sums the stellar table once, the dust table once per cycle and the detector arrays once, at write())");
        if (*s.state == Engine::ACC_LOCAL) todo |= s.bit;
    }
    if (!todo) return;
    ncclComm_t comm = (ncclComm_t)e.nccl;
    const size_t Nl = (size_t)e.instrNlambda;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (elapsedMs) { SKG_CUDA(cudaEventCreate(&ev0)); SKG_CUDA(cudaEventCreate(&ev1)); SKG_CUDA(cudaEventRecord(ev0, e.stream)); }
    SKG_NCCL(nccl.groupStart());
    if ((todo & SKG_REDUCE_LABS_STELLAR) && e.labs.p && e.labsCount) SKG_NCCL(nccl.allReduce(e.labs.p, e.labs.p, (size_t)e.labsCount, ncclDouble, ncclSum, comm, e.stream));
    if ((todo & SKG_REDUCE_LABS_DUST) && e.labsDust.p && e.labsCount) SKG_NCCL(nccl.allReduce(e.labsDust.p, e.labsDust.p, (size_t)e.labsCount, ncclDouble, ncclSum, comm, e.stream));
    if (todo & SKG_REDUCE_INSTRUMENTS)
        for (const InstrDev& d : e.instr)
        {
            const size_t Nf = (size_t)d.Nxp * d.Nyp;
            if (d.frame) SKG_NCCL(nccl.allReduce(d.frame, d.frame, (size_t)d.frameCount, ncclDouble, ncclSum, comm, e.stream));
            if (d.sed) SKG_NCCL(nccl.allReduce(d.sed, d.sed, Nl, ncclDouble, ncclSum, comm, e.stream));
            if (d.chanFrame) SKG_NCCL(nccl.allReduce(d.chanFrame, d.chanFrame, Nf * Nl * d.Nchan, ncclDouble, ncclSum, comm, e.stream));
            if (d.chanSed) SKG_NCCL(nccl.allReduce(d.chanSed, d.chanSed, Nl * d.Nchan, ncclDouble, ncclSum, comm, e.stream));
        }
    SKG_NCCL(nccl.groupEnd());
    if (elapsedMs) SKG_CUDA(cudaEventRecord(ev1, e.stream));
    e.sync();
    if (elapsedMs) { float ms = 0; SKG_CUDA(cudaEventElapsedTime(&ms, ev0, ev1)); *elapsedMs = ms; cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
    for (const Sel& s : sel) if (todo & s.bit) *s.state = Engine::ACC_GLOBAL;
}

// PanDustSystem::Labsdusttot / Labsstellartot: total of one absorption table over cells, wavelengths and processes
static double labsTotal(Engine& e, int which)
{
    const int state = which ? e.accLabsDust : e.accLabs;
    double local = mcLabsTotal(e, which);           // 0 when the table does not exist
    if (!e.nccl || e.nranks <= 1) return local;
    if (state == Engine::ACC_MIXED) throw Error("the absorption table holds a mix of summed and rank-local contributions");
    ncclComm_t comm = (ncclComm_t)e.nccl;
    e.scalarDev.ensure(sizeof(double));
    SKG_CUDA(cudaMemcpyAsync(e.scalarDev.p, &local, sizeof(double), cudaMemcpyHostToDevice, e.stream));
    // rank-local tables: sum_all of the scalar, like the reference; a table that is already the global sum holds the
    // total on every rank -- rank 0's value is handed out so that all ranks compare the very same number
    if (state == Engine::ACC_GLOBAL) SKG_NCCL(nccl.broadcast(e.scalarDev.p, e.scalarDev.p, 1, ncclDouble, 0, comm, e.stream));
    else SKG_NCCL(nccl.allReduce(e.scalarDev.p, e.scalarDev.p, 1, ncclDouble, ncclSum, comm, e.stream));
    double total = 0;
    e.readSmall(&total, e.scalarDev.p, sizeof(double));
    return total;
}

}   // namespace skg

using namespace skg;
extern "C"
{
int skg_comm_unique_id(void* out)
{
    try { if (!out) throw Error("null output"); loadNccl(); ncclUniqueId id; SKG_NCCL(nccl.getUniqueId(&id)); memcpy(out, &id, 128); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_comm_init(skg_engine* eh, int rank, int nranks, const void* uid)
{
    try
    {
        Engine& e = engineOf(eh);
        if (!uid) throw Error("null NCCL unique id");
        if (nranks < 1 || rank < 0 || rank >= nranks) throw Error("rank out of range");
        loadNccl();
        destroyComm(e);
        ncclUniqueId id; memcpy(&id, uid, 128);
        ncclComm_t comm; SKG_NCCL(nccl.commInitRank(&comm, nranks, id, rank));
        e.nccl = comm; e.rank = rank; e.nranks = nranks;
        return 0;
    }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_allreduce(skg_engine* eh, int which, double* elapsedMs)
{
    try { allreduce(engineOf(eh), which, elapsedMs); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_allreduce_results(skg_engine* eh) { return skg_allreduce(eh, SKG_REDUCE_ALL, nullptr); }
int skg_labs_dust_total(skg_engine* eh, double* total)
{
    try { if (!total) throw Error("null output"); *total = labsTotal(engineOf(eh), 1); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_labs_stellar_total(skg_engine* eh, double* total)
{
    try { if (!total) throw Error("null output"); *total = labsTotal(engineOf(eh), 0); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
}
