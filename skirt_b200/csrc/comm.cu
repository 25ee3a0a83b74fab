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
typedef int (*fnGroup)(void);
typedef const char* (*fnErr)(int);
static struct { void* lib = nullptr; fnGetUniqueId getUniqueId; fnCommInitRank commInitRank; fnAllReduce allReduce; fnGroup groupStart, groupEnd; fnErr errString; } nccl;

static void loadNccl()
{
    if (nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (nccl.lib) break; }
    if (!nccl.lib) throw Error(std::string("cannot load NCCL: ") + dlerror());
    nccl.getUniqueId = (fnGetUniqueId)dlsym(nccl.lib, "ncclGetUniqueId");
    nccl.commInitRank = (fnCommInitRank)dlsym(nccl.lib, "ncclCommInitRank");
    nccl.allReduce = (fnAllReduce)dlsym(nccl.lib, "ncclAllReduce");
    nccl.groupStart = (fnGroup)dlsym(nccl.lib, "ncclGroupStart");
    nccl.groupEnd = (fnGroup)dlsym(nccl.lib, "ncclGroupEnd");
    nccl.errString = (fnErr)dlsym(nccl.lib, "ncclGetErrorString");
    if (!nccl.getUniqueId || !nccl.commInitRank || !nccl.allReduce || !nccl.groupStart || !nccl.groupEnd) throw Error("NCCL symbols missing");
}
#define SKG_NCCL(call) do { int rc__ = (call); if (rc__ != 0) throw skg::Error(std::string(#call) + ": " + (nccl.errString ? nccl.errString(rc__) : "NCCL error")); } while (0)

}   // namespace skg

using namespace skg;
extern "C"
{
int skg_comm_unique_id(void* out)
{
    try { loadNccl(); ncclUniqueId id; SKG_NCCL(nccl.getUniqueId(&id)); memcpy(out, &id, 128); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_comm_init(skg_engine* eh, int rank, int nranks, const void* uid)
{
    try
    {
        Engine& e = *reinterpret_cast<Engine*>(eh);
        SKG_CUDA(cudaSetDevice(e.device));
        loadNccl();
        ncclUniqueId id; memcpy(&id, uid, 128);
        ncclComm_t comm; SKG_NCCL(nccl.commInitRank(&comm, nranks, id, rank));
        e.nccl = comm; e.rank = rank; e.nranks = nranks;
        return 0;
    }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
// replaces Instrument::sumResults (Instrument.cpp:57-65) and PanDustSystem::sumResults (PanDustSystem.cpp:394-404):
// one grouped in-place ncclAllReduce(double, sum) over Labs and every detector array
int skg_allreduce_results(skg_engine* eh)
{
    try
    {
        Engine& e = *reinterpret_cast<Engine*>(eh);
        SKG_CUDA(cudaSetDevice(e.device));
        if (!e.nccl || e.nranks <= 1) return 0;
        const int ncclDouble = 8, ncclSum = 0;       // nccl.h: ncclFloat64 = 8, ncclSum = 0
        ncclComm_t comm = (ncclComm_t)e.nccl;
        SKG_NCCL(nccl.groupStart());
        if (e.labs.p && e.labsCount) SKG_NCCL(nccl.allReduce(e.labs.p, e.labs.p, (size_t)e.labsCount, ncclDouble, ncclSum, comm, e.stream));
        for (const InstrDev& d : e.instr)
        {
            if (d.frame) SKG_NCCL(nccl.allReduce(d.frame, d.frame, (size_t)d.Nxp * d.Nyp * e.med.Nlambda, ncclDouble, ncclSum, comm, e.stream));
            if (d.sed) SKG_NCCL(nccl.allReduce(d.sed, d.sed, (size_t)e.med.Nlambda, ncclDouble, ncclSum, comm, e.stream));
            if (d.chanFrame) SKG_NCCL(nccl.allReduce(d.chanFrame, d.chanFrame, (size_t)d.Nxp * d.Nyp * e.med.Nlambda * d.Nchan, ncclDouble, ncclSum, comm, e.stream));
            if (d.chanSed) SKG_NCCL(nccl.allReduce(d.chanSed, d.chanSed, (size_t)e.med.Nlambda * d.Nchan, ncclDouble, ncclSum, comm, e.stream));
        }
        SKG_NCCL(nccl.groupEnd());
        e.sync();
        return 0;
    }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
}
