/* skirtgpu.h -- C ABI of the B200-native photon-packet engine (libskirtgpu.so).
 *
 * This is the drop-in boundary for SKIRT's propagation hot path.  The
 * reference has no FFI; its seam is C++ virtual dispatch on SimulationItem subclasses.  Each entry
 * point below names the reference interface it replaces (file:line relative to the reference
 * tree); INTEGRATION.md shows the adapter a SKIRT maintainer would add on the reference side.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a non-zero
 * status, and skg_last_error() then returns a message (the reference-side adapter turns that into
 * `throw FATALERROR(msg)`, FatalError.hpp:47).  All floating point data is IEEE binary64, all
 * indices int32 unless stated, exactly as in the reference.  Pointers are HOST
 * pointers unless the parameter name starts with `d_` (device pointers on the engine's GPU).
 * Engine calls must come from one host thread at a time per engine (same rule as Parallel::call,
 * Parallel.cpp:79-80).  There is no CPU fallback: creation fails without a CUDA device.
 */
#ifndef SKIRTGPU_H
#define SKIRTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skg_engine skg_engine;

/* ---- engine life cycle ------------------------------------------------------------------------ */
int skg_engine_create(int device, skg_engine** out);
void skg_engine_destroy(skg_engine* e);
const char* skg_last_error(void);
int skg_version(void);
/* the engine's CUDA stream (cudaStream_t) -- every kernel and copy of the engine is issued on it, so host code can
 * record its own events on it or order other work after it; and the number of kernels launched so far */
int skg_stream(skg_engine* e, void** stream);
int skg_launch_count(skg_engine* e, uint64_t* launches);

/* page-locked host memory for result arrays: skg_fetch_* into such a buffer is a single DMA transfer (pageable
 * destinations are staged by the driver at a fraction of the PCIe rate) */
int skg_host_alloc(size_t bytes, void** ptr);
int skg_host_free(void* ptr);
/* copy of an engine-owned device array (e.g. the result of skg_dust_cell_luminosities) to the host, on the engine's stream */
int skg_copy_to_host(skg_engine* e, const void* d_src, void* host, size_t bytes);

/* ---- dust grids: replace DustGrid::path / whichcell / randomPositionInCell (DustGrid.hpp:89-106) -- */

/* CartesianDustGrid (CartesianDustGrid.cpp:28-43): bin borders _xv[Nx+1], _yv[Ny+1], _zv[Nz+1];
 * cell number m = k + Nz*j + Nz*Ny*i (CartesianDustGrid.cpp:326-329). */
int skg_grid_cartesian(skg_engine* e, const double* xv, int Nx, const double* yv, int Ny, const double* zv, int Nz);

/* The grids with symmetries.  Borders as the reference's setupSelfAfter leaves them:
 *   Sphere1DDustGrid (Sphere1DDustGrid.cpp:24-33, path :111-186): rv[Nr+1] = mesh * maxR; cell m = i
 *   Sphere2DDustGrid (Sphere2DDustGrid.cpp:27-75, path :230-345): rv[Nr+1], thetav[Ntheta+1] = mesh * pi and their cosines cv with
 *       cv[0] = 1, cv[Ntheta] = -1 and exactly one border in the xy-plane (cv == 0; the reference inserts it when the mesh has
 *       none); cell m = k + Ntheta*i
 *   Cylinder2DDustGrid (Cylinder2DDustGrid.cpp:26-41, path :135-374): Rv[NR+1] = mesh * maxR, zv[Nz+1]; cell m = k + Nz*i */
int skg_grid_sphere1d(skg_engine* e, int Nr, const double* rv);
int skg_grid_sphere2d(skg_engine* e, int Nr, const double* rv, int Ntheta, const double* thetav, const double* cv);
int skg_grid_cylinder2d(skg_engine* e, int NR, const double* Rv, int Nz, const double* zv);

/* TreeDustGrid (TreeDustGrid.cpp:50-164): the node vector _tree flattened in id order.
 *   kind   : 0 OctTreeDustGrid, 1 BinTreeDustGrid
 *   search : 0 TopDown, 1 Neighbor, 2 Bookkeeping (TreeDustGrid.hpp:155; Bookkeeping is octree-only);
 *            3 the traversal of ParticleTreeDustGrid (ParticleTreeDustGrid.cpp:262-320: its own wall selection, top-down search)
 *   box[6*l..] = xmin,ymin,zmin,xmax,ymax,zmax of node l (TreeNode : Box)
 *   child0[l]  = id of the first child (children have consecutive ids, OctTreeNode.cpp:38-49), -1 for leaves
 *   parent[l]  = id of the father, -1 for the root;  cell[l] = _cellnumberv[l] (m for leaves, else -1)
 *   dir[l]     = BinTreeNode::_dir (0 x, 1 y, 2 z), ignored for octrees
 *   nbrStart[6*l+w] .. nbrStart[6*l+w+1] index nbrIds: TreeNode::_neighbors[w] IN STORED ORDER
 *                (walls BACK,FRONT,LEFT,RIGHT,BOTTOM,TOP = 0..5, TreeNode.hpp:100); may be NULL unless search==1 */
int skg_grid_tree(skg_engine* e, int kind, int search, int Nnodes, const double* box, const int* child0,
                  const int* parent, const int* cell, const int* dir, const int* nbrStart, const int* nbrIds);

/* AdaptiveMesh (AdaptiveMesh.cpp:21-57, AdaptiveMeshNode.cpp:14-80): nodes numbered so that the
 * children of a node are consecutive in local Morton order (k*Ny+j)*Nx+i starting at child0[l].
 *   nxyz[3*l..] = _Nx,_Ny,_Nz (0 for leaves); cell[l] = _m; wallNbr[6*l+w] = neighbour node beyond wall w or -1 */
int skg_grid_amesh(skg_engine* e, int Nnodes, const double* box, const int* nxyz, const int* child0,
                   const int* cell, const int* wallNbr);

/* VoronoiMesh (VoronoiMesh.cpp:310-393): particles[3*m..], neighbour lists in Voro++ order (ids >= 0
 * cells, -1..-6 domain walls), block lists indexed i*nb*nb+j*nb+k, per-block kd-trees (blkTree[b] =
 * root kd node or -1; node arrays kdM (cell), kdAxis, kdUp, kdLeft, kdRight with -1 = none) and the
 * enclosing box of every cell (for randomPosition, VoronoiMesh.cpp:591-606). */
int skg_grid_voronoi(skg_engine* e, int Ncells, const double* particles, const int* nbrStart, const int* nbrIds,
                     const double* extent, int nb, const int* blkStart, const int* blkIds, const int* blkTree,
                     int Nkd, const int* kdM, const int* kdAxis, const int* kdUp, const int* kdLeft,
                     const int* kdRight, const double* cellBox);

int skg_num_cells(skg_engine* e);

/* ---- medium: DustSystem::_rhovv (DustSystem.hpp:434) + DustMix kappa tables (DustMix.cpp:55-90) ---- */
/* rho[m*Ncomp+h]; kext/ksca/g[h*Nlambda+ell] */
int skg_medium(skg_engine* e, int Ncells, int Ncomp, int Nlambda, const double* rho, const double* kext,
               const double* ksca, const double* g);

/* Polarisation: the Mueller matrix coefficients of every dust component, DustMix::addpolarization
 * (DustMix.cpp:325-361; e.g. ElectronDustMix.cpp:41-59): S11, S12, S33, S34 [(h*Nlambda + ell)*Ntheta + t] on the scattering
 * angles theta_t = t*pi/(Ntheta-1).  Call after skg_medium (a new medium is unpolarised) and before skg_instruments.
 * The engine derives the cumulative distribution of theta and the phase-function normalisation (DustMix.cpp:96-123);
 * the shooting phases then carry a Stokes vector per packet (StokesVector.cpp), sample the scattering angles from the
 * Mueller matrix (DustMix::scatteringDirectionAndPolarization, DustMix.cpp:584-605, sampleTheta / samplePhi :716-731),
 * weight the peel-off by the polarised phase function (:648-662) and record Stokes Q, U, V in FullInstruments
 * (scatteringPeelOffPolarization :619-644, FullInstrument.cpp:136-141,165-170).  Like the reference (DustSystem.cpp:74)
 * either every component supports polarisation or none. */
int skg_medium_polarization(skg_engine* e, int Ntheta, const double* S11, const double* S12, const double* S33, const double* S34);

/* ---- deterministic geometry: batched DustGrid::path() + DustGridPath::fillOpticalDepth() ------------ */
/* Replaces DustSystem::fillOpticalDepth (DustSystem.cpp:959-980) for n rays at once.
 * Step 1 counts the segments of every ray and returns CSR offsets (offsets[n+1], int64) and the total;
 * step 2 fills segments[offsets[i] .. offsets[i+1]) for the same rays (device pointers must be 32-byte aligned).
 * ell[n] may be NULL (geometry only: dtau = tau = 0) or point to ONE value when ell_stride == 0. */
#define SKG_HOST 0
#define SKG_DEVICE 1
/* one path segment: the reference's DustGridPath::Segment { int m; double ds, s, dtau, tau; } (DustGridPath.hpp:161-167),
 * same 40-byte layout (4 bytes of padding after m), so that a batch of paths can be copied straight into
 * DustGridPath::_v vectors */
typedef struct skg_segment { int32_t m; int32_t reserved; double ds, s, dtau, tau; } skg_segment;
int skg_path_count(skg_engine* e, int mem, int64_t n, const double* r, const double* k, int64_t* offsets, int64_t* total);
int skg_path_fill(skg_engine* e, int mem, int64_t n, const double* r, const double* k, const int* ell, int ell_stride,
                  const int64_t* offsets, skg_segment* segments);
/* The same in ONE traversal per ray (what the reference does: DustGrid::path fills the caller's DustGridPath in a single
 * pass).  Ray i receives the slab segments[starts[i] .. starts[i+1]) of which the first lengths[i] records are its path;
 * slabs are sized without walking -- on Cartesian grids from the closed form of the number of crossings plus a few spare
 * records per ray (about 5 % of the total), on the other grids by the counting pass (then lengths[i] == slab size).
 * *needed receives the number of records the segment array must hold; when capacity is smaller (or segments is NULL)
 * nothing is written and the call succeeds only if segments is NULL (a size query).  Device segment arrays must be
 * 32-byte aligned.  A path that outgrows its slab (never observed) makes the call fail. */
int skg_path_batch(skg_engine* e, int mem, int64_t n, const double* r, const double* k, const int* ell, int ell_stride,
                   int64_t* starts /*[n+1]*/, int32_t* lengths /*[n]*/, skg_segment* segments, int64_t capacity, int64_t* needed);
/* DustSystem::opticaldepth(pp, distance) (DustSystem.cpp:984-1000); distance may be NULL (= DBL_MAX) */
int skg_opticaldepth(skg_engine* e, int mem, int64_t n, const double* r, const double* k, const int* ell, int ell_stride,
                     const double* distance, double* tau);
/* the same quantity computed with the walker of the photon SHOOTING stages (skg_run_stellar / skg_run_dust), whose results
 * are Monte Carlo estimates gated at 3 sigma rather than bit-exact paths: on Cartesian grids that walker carries the
 * crossing in the path-length parameter (no divisions; csrc/geom.cuh CartFastWalker) and may differ from the exact one by
 * rounding; on the other grids the two are the same.  Exposed so that the deviation can be measured (tests: <= 1e-10). */
int skg_opticaldepth_mc(skg_engine* e, int mem, int64_t n, const double* r, const double* k, const int* ell, int ell_stride,
                        const double* distance, double* tau);
/* DustGrid::whichcell (DustGrid.hpp:89) */
int skg_whichcell(skg_engine* e, int mem, int64_t n, const double* r, int* m);
/* number of "stuck packet" escapes / terminations since engine creation (the reference logs warnings,
 * TreeDustGrid.cpp:437-454, AdaptiveMesh.cpp:348-365) */
int skg_stuck_counts(skg_engine* e, int64_t* escaped, int64_t* terminated);
/* self test of the engine's invariant-divisor fp64 division against the IEEE division on n pseudo-random operand
 * pairs (the walkers divide by the direction cosines, which are constant along a path); *mismatches must be 0 */
int skg_selftest_division(skg_engine* e, uint64_t n, uint64_t seed, uint64_t* mismatches);

/* measurement aid: the rate (per second) at which this GPU retires n fp64 atomicAdds to pseudo-random cells of a table of
 * `cells` doubles -- the access pattern of the escape + absorption stage (LockFree::add into _Labsvv), whose ceiling it is */
int skg_selftest_atomics(skg_engine* e, uint64_t n, int cells, double* atomicsPerSecond);

/* ---- sources: StellarSystem::launch (StellarSystem.cpp:116-158) -------------------------------------- */
enum { SKG_GEOM_EXPDISK = 1, SKG_GEOM_SERSIC = 2 };
typedef struct skg_source
{
    int geometry;          /* SKG_GEOM_* */
    /* ExpDiskGeometry (ExpDiskGeometry.cpp:134-161): p[0]=hR p[1]=hz p[2]=Rmax p[3]=zmax p[4]=Rmin
     * SersicGeometry + SpheroidalGeometryDecorator: p[0]=Reff p[1]=flattening q; table = inverse-CDF grid */
    double p[8];
    /* SpiralStructureGeometryDecorator (SpiralStructureGeometryDecorator.cpp:177-192); arms == 0: none */
    int spiral_arms, spiral_index;
    double spiral_pitch, spiral_radius, spiral_phase, spiral_weight;
    /* tabulated radial CDF for Sersic (SersicGeometry.cpp:44-62): rv[ntab], Xv[ntab] (host pointers) */
    int ntab;
    const double* rv;
    const double* Xv;
    /* tabulated profile S(s) on the same radii (SersicFunction::_Sv), only needed by skg_sample_density */
    const double* Sv;
} skg_source;
/* L[h*Nlambda+ell] = StellarComp::luminosity(ell) of component h */
int skg_sources(skg_engine* e, int Ncomp, const skg_source* comps, int Nlambda, const double* L, double emissionBias);

/* ---- set-up side: DustSystem::setSampleDensityBody (DustSystem.cpp:152-177) on the device ----------
 * rho[m*Ncomp + h] = norm[h] * mean over sampleCount random positions in cell m (DustGrid::randomPositionInCell, the
 * same position for all components) of the geometry density of component h (ExpDiskGeometry.cpp:117-129,
 * SersicGeometry.cpp:66-70 + SpheroidalGeometryDecorator.cpp:47-52, SpiralStructureGeometryDecorator.cpp:49-58);
 * norm[h] is the mass normalisation of the component (e.g. FaceOnDustCompNormalization.cpp:67-74).  The grid must have
 * been set; geometries are described like the sources of skg_sources.  rho is a host array. */
int skg_sample_density(skg_engine* e, int Ncomp, const skg_source* geometries, const double* norm, int sampleCount,
                       uint64_t seed, double* rho);

/* Tree subdivision sampling (TreeDustGrid::subdivide, TreeDustGrid.cpp:168-233 -> TreeNodeSampleDensityCalculator.cpp:25-45)
 * for a batch of node boxes: mass[q] = volume(q) * mean over sampleCount random positions in box q of sum_h norm[h] *
 * density_h.  box[6*q..] = xmin,ymin,zmin,xmax,ymax,zmax; geometries and norm as for skg_sample_density.  Host arrays.
 * No grid needs to be set: this is what grows one (include/skirthost.h, skh_tree_*). */
int skg_sample_boxes(skg_engine* e, int64_t n, const double* box, int Ncomp, const skg_source* geometries, const double* norm,
                     int sampleCount, uint64_t seed, double* mass);

/* The same samples, and besides the mass (bit-identical to skg_sample_boxes) the density dispersion of every box,
 * dispersion[q] = (max - min) / max of the sampled total densities, 0 when max == 0
 * (TreeNodeSampleDensityCalculator::densityDispersion, TreeNodeSampleDensityCalculator.cpp:62-67): what
 * TreeDustGrid::subdivide compares with maxDensDispFraction (TreeDustGrid.cpp:215-221). */
int skg_sample_boxes_dispersion(skg_engine* e, int64_t n, const double* box, int Ncomp, const skg_source* geometries,
                                const double* norm, int sampleCount, uint64_t seed, double* mass, double* dispersion);

/* n launches of StellarSystem::launch(pp, ell, 1.0) with the engine's samplers (the kernel the shooting phase uses):
 * positions r[3n], directions k[3n] and bias-weighted luminosities L[n]; for distribution-level checks */
int skg_sample_launch(skg_engine* e, int ell, int n, uint64_t seed, double* r, double* k, double* L);

/* ---- instruments: DistantInstrument / SingleFrameInstrument / Frame-, SED-, SimpleInstrument ---------- */
enum { SKG_INSTR_FRAME = 1, SKG_INSTR_SED = 2, SKG_INSTR_SIMPLE = 3, SKG_INSTR_FULL = 4, SKG_INSTR_MULTIFRAME = 5, SKG_INSTR_PERSPECTIVE = 6 };
/* FullInstrument (FullInstrument.cpp:107-172, unpolarised part): one data cube + SED per channel, in this order */
enum { SKG_CHAN_TRANSPARENT = 0, SKG_CHAN_STELLAR_DIRECT = 1, SKG_CHAN_STELLAR_SCATTERED = 2, SKG_CHAN_DUST_DIRECT = 3,
       SKG_CHAN_DUST_SCATTERED = 4, SKG_CHAN_SCATTERING_LEVEL1 = 5 /* + (level - 1), level = 1..scatteringLevels */ };
/* with a polarised medium three more channels follow the scattering levels: Stokes Q, U, V of the total flux
 * (FullInstrument::_ftotQv/_ftotUv/_ftotVv): channel 5 + scatteringLevels + {0, 1, 2} */
/* MultiFrameInstrument (MultiFrameInstrument.cpp:85-99) + InstrumentFrame (InstrumentFrame.cpp:153-187): one frame per
 * wavelength, each with its own pixel grid; a packet of wavelength ell is recorded in frame ell only, in the total array
 * (writeTotal) and -- stellar packets -- in the array of the stellar component that emitted it (writeStellarComps) */
typedef struct skg_instrument_frame
{
    int Nxp, Nyp;
    double fovxp, fovyp, xpc, ypc;
} skg_instrument_frame;
typedef struct skg_instrument
{
    int kind;
    double distance, inclination, azimuth, positionAngle;       /* DistantInstrument.hpp */
    int Nxp, Nyp;                                                /* SingleFrameInstrument */
    double fovxp, fovyp, xpc, ypc;
    int scatteringLevels;                                        /* FullInstrument::setScatteringLevels (0 for the other kinds) */
    int writeTotal, writeStellarComps;                           /* MultiFrameInstrument only */
    const skg_instrument_frame* frames;                          /* MultiFrameInstrument only: one per wavelength of the medium; else NULL */
    /* PerspectiveInstrument (PerspectiveInstrument.cpp:36-108; distance and angles unused): Nxp x Nyp square pixels over a viewport
     * of width fovxp centred on `view`, looking at `cross`hair with `up` upwards, the eye `focal` behind the viewport.  The peel-off
     * direction is from the packet towards the eye (bfkobs(bfr), :290-304); the frame is fetched with skg_fetch_frame */
    double viewX, viewY, viewZ, crossX, crossY, crossZ, upX, upY, upZ, focal;
} skg_instrument;
int skg_instruments(skg_engine* e, int n, const skg_instrument* instr);
/* one frame of a MultiFrameInstrument: which = -1 the total flux, k >= 0 the flux of stellar component k (InstrumentFrame's
 * _ftotv / _fcompvv[k]); frame[Nxp*Nyp] of frame ell; add as for skg_fetch_frame */
int skg_fetch_multiframe(skg_engine* e, int instrument, int which, int ell, double* frame, int add);
/* detector arrays of one FullInstrument channel (replace FullInstrument's private _f*v / _F*v arrays); add as for skg_fetch_frame */
int skg_fetch_frame_channel(skg_engine* e, int instrument, int channel, double* frame, int add);
int skg_fetch_sed_channel(skg_engine* e, int instrument, int channel, double* sed, int add);

/* ---- photon shooting: MonteCarloSimulation::runstellaremission (MonteCarloSimulation.cpp:251-301) ------ */
typedef struct skg_mc_params
{
    double packages;            /* packets per wavelength for THIS engine (the caller splits the budget over GPUs) */
    double luminosityScale;     /* L per packet = luminosity(ell) / Npp_total ; pass Npp_total here */
    double minWeightReduction;  /* MonteCarloSimulation.cpp:32, default 1e4 */
    double minScattEvents;      /* default 0 */
    double scattBias;           /* xi, default 0.5 */
    int storeAbsorption;        /* DustSystem::storeabsorptionrates() */
    uint64_t seed;              /* Philox key; replaces Random's seed (Random.cpp:21) */
    uint64_t streamOffset;      /* first global packet index of this engine (disjoint Philox counters per GPU) */
    int ellBegin, ellEnd;       /* wavelength range [ellBegin, ellEnd) to shoot */
    int poolPackets;            /* packets in flight on the device at a time (0: default 2^22) */
    int continuousScattering;   /* MonteCarloSimulation::continuousScattering (MonteCarloSimulation.cpp:287,291,367-434), default 0 */
} skg_mc_params;
typedef struct skg_mc_stats
{
    uint64_t packets;           /* launched */
    uint64_t pathSegments;      /* packet-steps (addSegment calls) over all traversals */
    uint64_t paths;             /* traversals (full, propagation and peel-off) */
    uint64_t scatterings;
    double kernel_ms;           /* device time of the shooting kernels (CUDA events) */
    uint64_t absorbSegments;    /* segments that added into the absorption table (one fp64 atomic each) */
    uint64_t detections;        /* detector updates: frame pixel or SED bin (one fp64 atomic each) */
    double launch_ms, peel_ms, absorb_ms, propagate_ms;     /* device time per stage kernel family (CUDA events) */
    uint64_t iterations;        /* wavefront iterations (one launch of every stage each) */
    uint64_t peelSegments, propagateSegments;   /* packet-steps of the peel-off and the propagation stage (the rest: escape + absorption) */
} skg_mc_stats;
int skg_run_stellar(skg_engine* e, const skg_mc_params* p, skg_mc_stats* stats);

/* ---- dust emission phases: PanMonteCarloSimulation::dodustselfabsorptionchunk / dodustemissionchunk
 *      (PanMonteCarloSimulation.cpp:187-238, 269-342) for all wavelengths at once ------------------------------- */
enum { SKG_PHASE_STELLAR = 0, SKG_PHASE_DUST_SELFABS = 1, SKG_PHASE_DUST_EMISSION = 2 };
/* Lcell[ell*Ncells + m] = Labsbol[m] * dustluminosity(m, ell)  (the vector Lv of :193-198 / :275-280 for every ell;
 * host memory when mem == SKG_HOST, device memory when SKG_DEVICE).  The engine builds the per-wavelength CDFs
 * (NR::cdf, NR.hpp:388-394), picks the emitting cell (self-absorption: natural distribution, :217-218; emission:
 * composite biasing with emissionBias, :296-312), draws the position with DustGrid::randomPositionInCell and an
 * isotropic direction, and runs the packet life cycle:
 *   SKG_PHASE_DUST_SELFABS   no peel-off; absorbed luminosity is added to the DUST absorption table
 *                            (PanDustSystem::absorb(..., ynstellar=false), PanDustSystem.cpp:304-316)
 *   SKG_PHASE_DUST_EMISSION  peel-off of emission and scattering towards every instrument; nothing is absorbed
 * p->storeAbsorption is ignored; p->packages is the number of packets per wavelength for this engine
 * (the caller applies the stage factor / emissionBoost, :142, :258). */
int skg_run_dust(skg_engine* e, const skg_mc_params* p, int phase, double emissionBias, int mem, const double* Lcell,
                 skg_mc_stats* stats);
/* ---- dust emission spectra between the phases: AllCellsDustLib + GreyBodyDustEmissivity ---------
 * skg_dust_library sets the tables DustLib::calculate needs (DustLib.cpp:59-193): cell volumes [Ncells], kappa_abs
 * [Ncomp*Nlambda], wavelengths and bin widths [Nlambda].  skg_dust_cell_luminosities then computes, on the device and
 * from the device-resident absorption tables,
 *     Lcell[ell*Ncells + m] = Labs(m) * DustLib::luminosity(m, ell)
 * i.e. mean intensity (DustSystem::meanintensityv, DustSystem.cpp:935-955) -> equilibrium temperature per component
 * (DustMix::equilibrium / invplanckabs, DustMix.cpp:689-711, table of :238-262) -> kappa_abs * B(T) (GreyBodyDustEmissivity.cpp:22-45)
 * -> normalised cell SED (DustLib.cpp:126-158) times the bolometric absorbed luminosity (PanMonteCarloSimulation.cpp:193-198),
 * and returns the device pointer, ready for skg_run_dust(..., SKG_DEVICE, *d_Lcell, ...). */
int skg_dust_library(skg_engine* e, const double* volumes, const double* kappaabs, const double* lambda, const double* dlambda);
int skg_dust_cell_luminosities(skg_engine* e, double** d_Lcell);

/* PanDustSystem::rebootLabsdust (PanDustSystem.cpp:330-333), the dust absorption table, and
 * PanDustSystem::Labs(m) (PanDustSystem.cpp:337-348): bolometric absorbed luminosity per cell, stellar + dust */
int skg_reset_labs_dust(skg_engine* e);
int skg_fetch_labs_dust(skg_engine* e, double* labs /* [Ncells*Nlambda] */, int add);
int skg_labs_bolometric(skg_engine* e, double* Labsbol /* [Ncells] */);

/* accumulators (replace LockFree::add targets: Instrument _ftotv/_Ftotv, DustSystem _Labsvv) */
int skg_reset_results(skg_engine* e);
/* frame cube [Nxp*Nyp*Nlambda] (index l + ell*Nframep, FrameInstrument.cpp:38-39), sed [Nlambda];
 * add != 0: add into the host array (so that Instrument::write() runs unchanged afterwards) */
int skg_fetch_frame(skg_engine* e, int instrument, double* frame, int add);
int skg_fetch_sed(skg_engine* e, int instrument, double* sed, int add);
int skg_fetch_labs(skg_engine* e, double* labs /* [Ncells*Nlambda] */, int add);
/* Results to the host WHILE the engine goes on (the next phase, the next simulation of a series): skg_results_snapshot
 * copies every accumulator into a shadow array on the engine's stream (device to device; the absorption tables already in
 * the (m, ell) layout of the host interface), skg_fetch_snapshot_async moves one shadow to host memory -- page-locked, see
 * skg_host_alloc -- on a second stream, and skg_fetch_snapshot_wait returns when every such transfer has landed.
 * which / part as for skg_device_accumulators (0 stellar Labs, -1 dust Labs, i+1 instrument i; part 0 frame(s), 1 SED(s));
 * *count (may be NULL) receives the number of doubles; host == NULL only queries it. */
int skg_results_snapshot(skg_engine* e);
int skg_fetch_snapshot_async(skg_engine* e, int which, int part, double* host, int64_t* count);
int skg_fetch_snapshot_wait(skg_engine* e);
/* device views of the accumulators, for collectives issued by the host (NCCL through torch.distributed);
 * note that the absorption table is wavelength-major on the device: labs[ell*Ncells + m] */
int skg_device_accumulators(skg_engine* e, int which /*0 labs, -1 dust labs, 1.. instruments*/, int part /*0 frame,1 sed*/,
                            double** d_ptr, int64_t* count);

/* ---- multi-GPU: replaces ProcessManager::sum / sum_all (ProcessManager.cpp:122-140) ------------------- */
/* NCCL communicator over one engine per GPU/process.  unique_id is the 128-byte ncclUniqueId that rank 0
 * obtained from skg_comm_unique_id and distributed by any means (MPI_Bcast, torch.distributed, a file).
 * A second skg_comm_init on the same engine destroys the previous communicator first. */
int skg_comm_unique_id(void* unique_id_128);
int skg_comm_init(skg_engine* e, int rank, int nranks, const void* unique_id_128);

/* The reference sums each accumulator over the processes at its own moment, and exactly once:
 *   SKG_REDUCE_LABS_STELLAR  the stellar absorption table, before the first dust emission spectra are made
 *                            (PanDustSystem::calculatedustemission(true) -> sumResults(true), PanDustSystem.cpp:383-404)
 *   SKG_REDUCE_LABS_DUST     the dust absorption table, after every self-absorption cycle (sumResults(false), called at
 *                            the start of the next cycle or of the emission phase, PanMonteCarloSimulation.cpp:131,249)
 *   SKG_REDUCE_INSTRUMENTS   every detector array, once, in Instrument::write() (Instrument.cpp:57-65)
 * skg_allreduce sums the selected accumulators in place over all ranks (one grouped ncclAllReduce(double, sum) on the
 * engine's stream).  The engine tracks per accumulator whether it holds rank-local additions only, the global sum, or
 * nothing since the last reset: an accumulator that already holds the global sum (or nothing) is skipped, so the call
 * is idempotent; one that received rank-local additions AFTER it was summed cannot be summed again in place and is an
 * error ("... already summed over the processes").  elapsed_ms (may be NULL) receives the device time of the collective.
 * Every rank of the communicator must make the same calls in the same order (as with sum_all). */
enum { SKG_REDUCE_LABS_STELLAR = 1, SKG_REDUCE_LABS_DUST = 2, SKG_REDUCE_INSTRUMENTS = 4, SKG_REDUCE_ALL = 7 };
int skg_allreduce(skg_engine* e, int which, double* elapsed_ms);
/* = skg_allreduce(e, SKG_REDUCE_ALL, NULL) */
int skg_allreduce_results(skg_engine* e);
/* PanDustSystem::Labsdusttot() (PanDustSystem.cpp:363-379): the total of the dust absorption table over all cells,
 * wavelengths AND processes, identical on every rank (the convergence test of the self-absorption cycles,
 * PanMonteCarloSimulation.cpp:152-167, must take the same decision everywhere).  Summed on the device; a rank-local
 * table adds a scalar all-reduce, a table that already holds the global sum is totalled as it is (rank 0's value is
 * broadcast).  skg_labs_stellar_total is PanDustSystem::Labsstellartot() (:351-359) with the same conventions. */
int skg_labs_dust_total(skg_engine* e, double* total);
int skg_labs_stellar_total(skg_engine* e, double* total);

#ifdef __cplusplus
}
#endif
#endif
