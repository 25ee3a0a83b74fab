// Host-side engine object behind the C ABI of include/skirtgpu.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/skirtgpu.h"
#include "tables.h"

namespace skg
{

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
void setLastError(const std::string& msg);   // message returned by skg_last_error()

#define SKG_CART_SMEM_MAX (64 * 1024)      // shared-memory budget of the staged Cartesian borders (3 arrays of N+1 doubles)

#define SKG_CUDA(call) do { cudaError_t err__ = (call); if (err__ != cudaSuccess) \
    throw skg::Error(std::string(#call) + ": " + cudaGetErrorString(err__)); } while (0)

// owning device allocation
struct DevBuf
{
    void* p = nullptr; size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    void ensure(size_t n) { if (n > bytes) { release(); SKG_CUDA(cudaMalloc(&p, n ? n : 1)); bytes = n; } }
    void upload(const void* host, size_t n, cudaStream_t st)
    { ensure(n); if (n) SKG_CUDA(cudaMemcpyAsync(p, host, n, cudaMemcpyHostToDevice, st)); }
    template<class T> T* as() const { return static_cast<T*>(p); }
};

// sampler tables for one stellar component on the device
struct SourceDev
{
    int geometry; double p[8];
    int spiral_arms, spiral_index; double spiral_pitch, spiral_radius, spiral_phase, spiral_weight;
    double spiral_c, spiral_tanp, spiral_cn;    // derived constants (SpiralStructureGeometryDecorator.cpp:24-45)
    int ntab; const double* rv; const double* Xv; const double* Sv;
    double rho0;                                // density normalisation of the bare geometry
};

// one InstrumentFrame of a MultiFrameInstrument (InstrumentFrame.cpp:29-44): its pixel grid and where its pixels start in a slab
struct FrameDev { int Nxp, Nyp; double xpmin, ypmin, xpsiz, ypsiz; long long offset; };
struct InstrDev
{
    int kind;
    double costheta, sintheta, cosphi, sinphi, cospa, sinpa;
    double kobsx, kobsy, kobsz;
    int Nxp, Nyp; double xpmin, ypmin, xpsiz, ypsiz;
    double* frame; double* sed;     // device accumulators (frame: Nxp*Nyp*Nlambda, sed: Nlambda)
    // FullInstrument: Nchan = 5 + Nscatt channels, channel-major: chanFrame[(c*Nlambda + ell)*Nxp*Nyp + l], chanSed[c*Nlambda + ell]
    int Nchan, Nscatt; double* chanFrame; double* chanSed;
    int pol;                        // FullInstrument of a simulation with polarisation: channels Nchan-3 .. Nchan-1 are Stokes Q, U, V
    double kyx, kyy, kyz;           // DistantInstrument::bfky (DistantInstrument.cpp:47-49): the frame's y axis in model coordinates
    long long frameCount;           // doubles in `frame`
    // MultiFrameInstrument: `frame` holds slabs of mfPixels doubles (the frames of all wavelengths one after the other): the total
    // flux (slab mfTotal, -1: not recorded) and one slab per stellar component (from slab mfComp0, -1: not recorded)
    const FrameDev* frames; long long mfPixels; int mfTotal, mfComp0, mfNcomp;
};

struct Engine
{
    int device = 0;
    cudaStream_t stream = nullptr;
    int smCount = 148;

    int gridKind = GRID_NONE;
    CartGrid cart{}; TreeGrid tree{}; AMeshGrid amesh{}; VoroGrid voro{}; SymGrid sym{};
    int Ncells = 0;
    std::vector<DevBuf*> gridBufs;
    // device arrays of replaced tables are kept for the next upload of that size: a series of simulations re-uploads its
    // tables without cudaFree (which would wait for the result transfers of skg_fetch_snapshot_async)
    // a few words from the device: stored by a one-warp kernel into mapped page-locked memory rather than copied, so that
    // the read does not queue on the device-to-host copy engine behind a result transfer in flight (<= 512 bytes, 8-byte words)
    void readSmall(void* dst, const void* devSrc, size_t bytes); void* smallHost = nullptr;
    std::vector<DevBuf*> spareBufs; DevBuf* takeBuf(size_t bytes); void recycle(std::vector<DevBuf*>& list);
    Medium med{};
    DevBuf rho, kext, ksca, gasym;
    DevBuf mueller[4], thetaX, pfnorm;     // polarisation tables (skg_medium_polarization)
    DevBuf counters;                    // Counters
    DevBuf scratchR, scratchK, scratchEll, scratchDist, scratchCounts, scratchOffsets, scratchCub, scratchOut[5], scratchTau, scratchM, scratchWork;

    // Monte Carlo state
    int Nsources = 0; int NlambdaSrc = 0; double emissionBias = 0.5;
    std::vector<SourceDev> sources; DevBuf sourcesDev, lumDev, lumCdfDev, lumTotDev; std::vector<DevBuf*> sourceBufs;
    std::vector<double> lumHost, lumTotHost;
    DevBuf perspDev; int Npersp = 0;        // PerspectiveInstruments (PerspDev), outside the observer groups
    std::vector<InstrDev> instr; DevBuf instrDev; std::vector<DevBuf*> instrBufs;
    DevBuf labs; int64_t labsCount = 0;    // absorbed luminosity, wavelength-major on the device: labs[ell*Ncells+m]
    DevBuf labsDust;                        // absorbed dust emission (self-absorption cycles), same layout
    DevBuf libVol, libKabs, libLambda, libDlambda, libTv, libPlanckabs; bool haveDustLib = false;     // DustLib tables
    DevBuf dustLvOut;                       // output of skg_dust_cell_luminosities
    DevBuf dustLv, dustCdf, dustLtot;       // per-wavelength cell luminosities of a dust phase, their CDFs and totals
    DevBuf labsT;                           // scratch for the (m,ell) row-major copy handed to the host
    int instrNlambda = 0;                   // number of wavelengths the detector arrays were allocated for
    bool instrPol = false;                  // whether they were allocated for a polarised medium (FullInstrument Q, U, V)
    DevBuf instrGroupedDev, groupsDev; int Ngroups = 0, maxGroupCount = 0;    // instruments ordered by line of sight + the groups
    DevBuf mcPool, mcPolPool, mcLists, mcCounts, mcEllList; int* mcHostCounts = nullptr;   // packet pool of the wavefront shooter
    void* nccl = nullptr; int rank = 0, nranks = 1;
    // what each accumulator holds with respect to the other processes (skg_allreduce): nothing since the last reset,
    // rank-local additions only, the sum over all ranks, or the sum over all ranks plus later rank-local additions
    enum AccState { ACC_ZERO = 0, ACC_LOCAL = 1, ACC_GLOBAL = 2, ACC_MIXED = 3 };
    int accLabs = ACC_ZERO, accLabsDust = ACC_ZERO, accInstr = ACC_ZERO;
    void touched(int& st) const { st = (nranks > 1 && (st == ACC_GLOBAL || st == ACC_MIXED)) ? ACC_MIXED : ACC_LOCAL; }
    // opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE function attribute: set once per engine
    bool attrPath = false, attrFill = false, attrStages = false;
    DevBuf scalarDev;                       // one double: device-side totals (skg_labs_dust_total)
    // skg_results_snapshot: copies of the accumulators that a second stream hands to the host while the next phase runs
    struct Shadow { int which, part; int64_t count; DevBuf* buf; };
    std::vector<Shadow> shadows; cudaStream_t copyStream = nullptr; cudaEvent_t snapEvent = nullptr;
    double stageMs[4] = {0, 0, 0, 0}; uint64_t mcIterations = 0;     // launch, peel, absorb, propagate device time of the last phase
    cudaEvent_t mcEvents[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t launches = 0;              // kernels launched by this engine (skg_launch_count)

    explicit Engine(int dev);
    ~Engine();
    void freeGrid();
    Counters* ctr() { return counters.as<Counters>(); }
    Counters readCounters();
    void sync() { SKG_CUDA(cudaStreamSynchronize(stream)); }
};

// kernels launchers (path_kernels.cu)
void launchPathCount(Engine& e, int64_t n, const double* d_r, const double* d_k, int* d_counts);
void launchPathFill(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                    const int64_t* d_offsets, skg_segment* d_segments, int* d_lengths = nullptr);
void launchPathCapacity(Engine& e, int64_t n, const double* d_r, const double* d_k, int* d_cap);
void launchOpticalDepth(Engine& e, int64_t n, const double* d_r, const double* d_k, const int* d_ell, int ellStride,
                        const double* d_dist, double* d_tau, bool mcWalker = false);
void launchWhichCell(Engine& e, int64_t n, const double* d_r, int* d_m);
unsigned long long runDivisionSelfTest(Engine& e, unsigned long long n, unsigned long long seed);
void exclusiveScan(Engine& e, int64_t n, const int* d_counts, int64_t* d_offsets);   // writes n+1 offsets

// Monte Carlo (mc_kernels.cu)
void mcSetSources(Engine& e, int Ncomp, const skg_source* comps, int Nlambda, const double* L, double emissionBias);
void mcSetInstruments(Engine& e, int n, const skg_instrument* instr);
void mcSetPolarization(Engine& e, int Ntheta, const double* S11, const double* S12, const double* S33, const double* S34);
void mcRunStellar(Engine& e, const skg_mc_params& p, skg_mc_stats* stats);
void mcResetResults(Engine& e);
double mcLabsTotal(Engine& e, int which);     // sum over the whole (stellar: 0, dust: 1) absorption table, on the device
void destroyComm(Engine& e);                  // comm.cu
void mcFetchLabs(Engine& e, double* host, int add, int which);
void mcTransposeLabs(Engine& e, const double* src, double* dst);      // wavelength-major device table -> (m, ell) row-major, on the engine's stream
void mcLabsBolometric(Engine& e, double* host);
void mcDustLibrary(Engine& e, const double* volumes, const double* kappaabs, const double* lambda, const double* dlambda);
double* mcDustCellLuminosities(Engine& e);
void mcSampleDensity(Engine& e, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* rho);
void mcSampleBoxes(Engine& e, int64_t n, const double* box, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* mass, double* dispersion = nullptr);
double mcAtomicRate(Engine& e, uint64_t n, int cells);
void mcSampleLaunch(Engine& e, int ell, int n, uint64_t seed, double* r, double* k, double* L);
void mcRunDust(Engine& e, const skg_mc_params& p, int phase, double emissionBias, int mem, const double* Lcell, skg_mc_stats* stats);

}   // namespace skg
