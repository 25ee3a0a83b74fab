// oracle/_ref harness -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product path).
//
// C-ABI front end over the reference's OWN classes (compiled in place from /root/reference by
// oracle/Makefile).  A simulation hierarchy is assembled with the reference's public setters from a
// small line-oriented spec, exactly like Discover/XmlHierarchyCreator would do from a ski file, and
// is then set up and run by the reference's own Simulation::setup()/run().  The harness adds
//   * getters that flatten the reference's internal state (grid, densities, optical properties,
//     luminosities, detector arrays, absorbed luminosities) into the POD tables of include/skirtgpu.h,
//   * skr_path_batch(): DustGrid::path() + DustGridPath::fillOpticalDepth() for a batch of fixed rays.
// Private members are read through the `#define private public` idiom in THIS translation unit only.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <valarray>
#include <unordered_map>
#include <condition_variable>

#define private public
#define protected public
#include "VoronoiMesh.cpp"      // compiled here (not as its own object) to reach VoronoiMesh_Private
#include "AdaptiveMesh.hpp"
#include "AdaptiveMeshNode.hpp"
#include "AdaptiveMeshFile.hpp"
#include "AdaptiveMeshDustDistribution.hpp"
#include "AdaptiveMeshDustGrid.hpp"
#include "MeshDustComponent.hpp"
#include "BinTreeDustGrid.hpp"
#include "BinTreeNode.hpp"
#include "BlackBodySED.hpp"
#include "BolLuminosityStellarCompNormalization.hpp"
#include "CartesianDustGrid.hpp"
#include "CompDustDistribution.hpp"
#include "Console.hpp"
#include "DustComp.hpp"
#include "DustMix.hpp"
#include "DustSystem.hpp"
#include "ExpDiskGeometry.hpp"
#include "FaceOnDustCompNormalization.hpp"
#include "DustMassDustCompNormalization.hpp"
#include "FatalError.hpp"
#include "FrameInstrument.hpp"
#include "GeometricStellarComp.hpp"
#include "GreyBodyDustEmissivity.hpp"
#include "AllCellsDustLib.hpp"
#include "InstrumentSystem.hpp"
#include "InterstellarDustMix.hpp"
#include "LinMesh.hpp"
#include "LogWavelengthGrid.hpp"
#include "NestedLogWavelengthGrid.hpp"
#include "MultiFrameInstrument.hpp"
#include "PerspectiveInstrument.hpp"
#include "InstrumentFrame.hpp"
#include "OctTreeDustGrid.hpp"
#include "Sphere1DDustGrid.hpp"
#include "Sphere2DDustGrid.hpp"
#include "Cylinder2DDustGrid.hpp"
#include "ParticleTreeDustGrid.hpp"
#include "DustParticleInterface.hpp"
#include "OctTreeNode.hpp"
#include "OligoDustSystem.hpp"
#include "OligoMonteCarloSimulation.hpp"
#include "OligoWavelengthGrid.hpp"
#include "PanDustSystem.hpp"
#include "PanMonteCarloSimulation.hpp"
#include "PanStellarComp.hpp"
#include "Parallel.hpp"
#include "ParallelFactory.hpp"
#include "PhotonPackage.hpp"
#include "PowMesh.hpp"
#include "Random.hpp"
#include "SEDInstrument.hpp"
#include "SersicGeometry.hpp"
#include "SIUnits.hpp"
#include "StellarUnits.hpp"
#include "ExtragalacticUnits.hpp"
#include "SimpleInstrument.hpp"
#include "FullInstrument.hpp"
#include "SpheroidalGeometryDecorator.hpp"
#include "SpiralStructureGeometryDecorator.hpp"
#include "StellarSystem.hpp"
#include "SymPowMesh.hpp"
#include "TreeDustGrid.hpp"
#include "TreeNode.hpp"
#include "VoronoiDustGrid.hpp"
#include "VoronoiMeshFile.hpp"
#undef private
#undef protected

extern long skr_warning_count;
extern int skr_verbose;

namespace
{
    thread_local std::string lastError;

    // ---- harness-side subclasses of the reference's abstract items (inputs from memory) ----

    // a dust mix whose per-wavelength properties are given on the simulation's wavelength grid;
    // everything downstream (kappa tables, albedo, HG sampling) is the reference's DustMix.cpp
    class TableDustMix : public DustMix
    {
    public:
        std::vector<double> kabs, ksca, g;
        int Ntheta = 0; std::vector<double> S11, S12, S33, S34;     // optional Mueller matrix coefficients [Nlambda*Ntheta]
        void setupSelfBefore()
        {
            DustMix::setupSelfBefore();
            int n = kabs.size();
            Array a(n), s(n), gg(n);
            for (int i = 0; i < n; i++) { a[i] = kabs[i]; s[i] = ksca[i]; gg[i] = g[i]; }
            addpopulation(1.0, a, s, gg);   // mu = 1 so that sigma == kappa, as InterstellarDustMix.cpp:58
            if (Ntheta > 0)
            {
                // a polarised mix, like ElectronDustMix.cpp:59 / MultiGrainDustMix: DustMix::addpolarization (DustMix.cpp:325-361)
                Table<2> t11(n, Ntheta), t12(n, Ntheta), t33(n, Ntheta), t34(n, Ntheta);
                for (int i = 0; i < n; i++) for (int t = 0; t < Ntheta; t++)
                { t11(i,t) = S11[(size_t)i*Ntheta+t]; t12(i,t) = S12[(size_t)i*Ntheta+t]; t33(i,t) = S33[(size_t)i*Ntheta+t]; t34(i,t) = S34[(size_t)i*Ntheta+t]; }
                addpolarization(t11, t12, t33, t34);
            }
        }
    };

    // a stellar component with given luminosities per wavelength; launch() is GeometricStellarComp's
    class TableStellarComp : public GeometricStellarComp
    {
    public:
        std::vector<double> L;
        void setupSelfBefore()
        {
            GeometricStellarComp::setupSelfBefore();
            _Lv.resize(L.size());
            for (size_t i = 0; i < L.size(); i++) _Lv[i] = L[i];
        }
    };

    // the dust distribution of the geometry components that also hands out a list of particle positions: what
    // ParticleTreeDustGrid asks its distribution for (DustParticleInterface; SPHDustDistribution in a real ski file)
    class ParticleCompDustDistribution : public CompDustDistribution, public DustParticleInterface
    {
    public:
        std::vector<double> xyz;
        int numParticles() const { return (int)(xyz.size() / 3); }
        Vec particleCenter(int i) const { return Vec(xyz[3*(size_t)i], xyz[3*(size_t)i+1], xyz[3*(size_t)i+2]); }
    };

    class MemVoronoiMeshFile : public VoronoiMeshFile
    {
    public:
        std::vector<double> xyz; long pos = -1;
        void open() { pos = -1; }
        void close() {}
        bool read() { pos++; return (size_t)(3*pos) < xyz.size(); }
        Vec particle() const { return Vec(xyz[3*pos], xyz[3*pos+1], xyz[3*pos+2]); }
        double value(int) const { return 0; }
    };

    // adaptive mesh description: one record per node in the reference's file order (depth-first);
    // nonleaf: nx,ny,nz > 0 ; leaf: nx = 0 and value = density
    class MemAdaptiveMeshFile : public AdaptiveMeshFile
    {
    public:
        std::vector<int> nxyz; std::vector<double> val; long pos = -1;
        void open() { pos = -1; }
        void close() {}
        bool read() { pos++; return (size_t)pos < val.size(); }
        bool isNonLeaf() const { return nxyz[3*pos] > 0; }
        void numChildNodes(int& nx, int& ny, int& nz) const { nx = nxyz[3*pos]; ny = nxyz[3*pos+1]; nz = nxyz[3*pos+2]; }
        double value(int) const { return val[pos]; }
    };

    struct Sim
    {
        MonteCarloSimulation* mc = 0;
        bool pan = false;
        OligoWavelengthGrid* olg = 0; PanWavelengthGrid* plg = 0;
        StellarSystem* ss = 0;
        DustSystem* ds = 0;
        DustDistribution* dd = 0; CompDustDistribution* cdd = 0; AdaptiveMeshDustDistribution* amdd = 0;
        InstrumentSystem* is = 0;
        DustGrid* grid = 0;
        double box[6] = {0,0,0,0,0,0};
        std::vector<TableStellarComp*> stars;
        std::vector<TableDustMix*> mixes;
        Geometry* lastGeom = 0;         // most recently created geometry (target for decorators)
        std::function<void(Geometry*)> lastGeomSetter;
        MemVoronoiMeshFile* vfile = 0; ParticleCompDustDistribution* pdd = 0;
        MemAdaptiveMeshFile* afile = 0;
        int gridKind = -1;              // 0 cartesian, 1 octtree, 2 bintree, 3 voronoi, 4 adaptive mesh, 5 particle tree (oct), 6 particle tree (bin)
        // flattened adaptive mesh numbering
        std::vector<const AdaptiveMeshNode*> amNodes; std::unordered_map<const AdaptiveMeshNode*, int> amIndex;
        // flattened voronoi kd nodes
        std::vector<VoronoiMesh_Private::Node*> kdNodes; std::unordered_map<VoronoiMesh_Private::Node*, int> kdIndex;
    };

    MoveableMesh* makeMesh(std::istringstream& in, int n)
    {
        std::string kind; in >> kind;
        MoveableMesh* mesh = 0;
        if (kind == "lin") mesh = new LinMesh();
        else if (kind == "sympow") { SymPowMesh* m = new SymPowMesh(); double r; in >> r; m->setRatio(r); mesh = m; }
        else if (kind == "pow") { PowMesh* m = new PowMesh(); double r; in >> r; m->setRatio(r); mesh = m; }
        else throw std::runtime_error("unknown mesh kind " + kind);
        mesh->setNumBins(n);
        return mesh;
    }

    void setBox(BoxDustGrid* g, const double* b)
    { g->setMinX(b[0]); g->setMaxX(b[1]); g->setMinY(b[2]); g->setMaxY(b[3]); g->setMinZ(b[4]); g->setMaxZ(b[5]); }

    Geometry* makeGeometry(Sim* S, std::istringstream& in)
    {
        std::string kind; in >> kind;
        if (kind == "expdisk")
        {
            double hR, hz, Rmax, zmax; in >> hR >> hz >> Rmax >> zmax;
            ExpDiskGeometry* g = new ExpDiskGeometry();
            g->setRadialScale(hR); g->setAxialScale(hz); g->setRadialTrunc(Rmax); g->setAxialTrunc(zmax);
            return g;
        }
        if (kind == "sersic")
        {
            double n, Re, q; in >> n >> Re >> q;
            SersicGeometry* g = new SersicGeometry();
            g->setIndex(n); g->setRadius(Re);
            if (q == 1.0) return g;
            SpheroidalGeometryDecorator* d = new SpheroidalGeometryDecorator();
            d->setGeometry(g); d->setFlattening(q);
            return d;
        }
        throw std::runtime_error("unknown geometry " + kind);
    }

    Geometry* maybeSpiral(Geometry* g, std::istringstream& in)
    {
        std::string word;
        if (in >> word)
        {
            if (word != "spiral") throw std::runtime_error("unexpected token " + word);
            int arms, index; double pitch, radius, phase, weight;
            in >> arms >> pitch >> radius >> phase >> weight >> index;
            SpiralStructureGeometryDecorator* d = new SpiralStructureGeometryDecorator();
            AxGeometry* ax = dynamic_cast<AxGeometry*>(g);
            if (!ax) throw std::runtime_error("spiral decorator needs an axisymmetric geometry");
            d->setGeometry(ax); d->setArms(arms); d->setPitch(pitch); d->setRadius(radius);
            d->setPhase(phase); d->setPerturbWeight(weight); d->setIndex(index);
            return d;
        }
        return g;
    }

    void build(Sim* S, const char* spec)
    {
        std::istringstream all(spec);
        std::string line;
        double packages = 1e6, mwr = 1e4, minscatt = 0, xi = 0.5, ebias = 0.5; int contscatt = 0;
        int threads = 1, seed = 4357, dustsamples = 100, storeabs = 0, selfabs = 0, unitsys = 0, fluxstyle = 0;
        std::vector<std::string> lines;
        while (std::getline(all, line)) if (!line.empty() && line[0] != '#') lines.push_back(line);

        // pass 1: the simulation type
        for (auto& l : lines) { std::istringstream in(l); std::string key; in >> key;
            if (key == "sim") { std::string t; in >> t; S->pan = (t == "pan"); } }
        if (S->pan) S->mc = new PanMonteCarloSimulation(); else S->mc = new OligoMonteCarloSimulation();
        S->ss = new StellarSystem();
        S->is = new InstrumentSystem();
        S->cdd = 0;

        // pass 2: everything else in order
        for (auto& l : lines)
        {
            std::istringstream in(l); std::string key; in >> key;
            if (key == "sim") continue;
            else if (key == "threads") in >> threads;
            else if (key == "seed") in >> seed;
            else if (key == "packages") in >> packages;
            else if (key == "minweightreduction") in >> mwr;
            else if (key == "minscatt") in >> minscatt;
            else if (key == "scattbias") in >> xi;
            else if (key == "continuousscattering") in >> contscatt;
            else if (key == "emissionbias") in >> ebias;
            else if (key == "dustsamples") in >> dustsamples;
            else if (key == "storeabs") in >> storeabs;
            else if (key == "units") in >> unitsys >> fluxstyle;      // 0 SI / 1 stellar / 2 extragalactic; 0 neutral / 1 wavelength / 2 frequency
            else if (key == "selfabs") in >> selfabs;
            else if (key == "wavelengths")
            {
                QList<double> lv; double v; while (in >> v) lv << v;
                S->olg = new OligoWavelengthGrid(); S->olg->setWavelengths(lv);
            }
            else if (key == "loggrid")
            {
                double a, b; int n; in >> a >> b >> n;
                LogWavelengthGrid* g = new LogWavelengthGrid(); g->setMinWavelength(a); g->setMaxWavelength(b); g->setPoints(n);
                S->plg = g;
            }
            else if (key == "nestedloggrid")      // min max points  zoom-min zoom-max zoom-points
            {
                double a, b, za, zb; int n, zn; in >> a >> b >> n >> za >> zb >> zn;
                NestedLogWavelengthGrid* g = new NestedLogWavelengthGrid(); g->setMinWavelength(a); g->setMaxWavelength(b); g->setPoints(n);
                g->setMinWavelengthSubGrid(za); g->setMaxWavelengthSubGrid(zb); g->setPointsSubGrid(zn);
                S->plg = g;
            }
            else if (key == "box") { for (int i = 0; i < 6; i++) in >> S->box[i]; }
            else if (key == "grid")
            {
                std::string kind; in >> kind;
                if (kind == "cartesian")
                {
                    int nx, ny, nz; in >> nx >> ny >> nz;
                    CartesianDustGrid* g = new CartesianDustGrid(); setBox(g, S->box);
                    g->setMeshX(makeMesh(in, nx)); g->setMeshY(makeMesh(in, ny)); g->setMeshZ(makeMesh(in, nz));
                    S->grid = g; S->gridKind = 0;
                }
                else if (kind == "octtree" || kind == "bintree")
                {
                    int minl, maxl, search, bary, samples; double mf; in >> minl >> maxl >> search >> mf >> bary >> samples;
                    TreeDustGrid* g;
                    if (kind == "octtree") { OctTreeDustGrid* o = new OctTreeDustGrid(); o->setBarycentric(bary != 0); g = o; S->gridKind = 1; }
                    else { BinTreeDustGrid* b = new BinTreeDustGrid();
                           b->setDirectionMethod(bary ? BinTreeDustGrid::Barycenter : BinTreeDustGrid::Alternating); g = b; S->gridKind = 2; }
                    setBox(g, S->box);
                    g->setMinLevel(minl); g->setMaxLevel(maxl); g->setSearchMethod((TreeDustGrid::SearchMethod)search);
                    g->setMaxMassFraction(mf); g->setSampleCount(samples);
                    S->grid = g;
                }
                else if (kind == "voronoi")
                {
                    std::string how; in >> how;
                    VoronoiDustGrid* g = new VoronoiDustGrid(); setBox(g, S->box);
                    if (how == "file") { S->vfile = new MemVoronoiMeshFile(); g->setVoronoiMeshFile(S->vfile); g->setDistribution(VoronoiDustGrid::File); }
                    else { int n; in >> n; g->setNumParticles(n);
                           g->setDistribution(how == "uniform" ? VoronoiDustGrid::Uniform : how == "peak" ? VoronoiDustGrid::CentralPeak : VoronoiDustGrid::DustDensity); }
                    S->grid = g; S->gridKind = 3;
                }
                else if (kind == "amesh") { S->grid = new AdaptiveMeshDustGrid(); S->gridKind = 4; }
                else if (kind == "sphere1d")
                {
                    double rmax; int n; in >> rmax >> n;            // grid sphere1d <maxR> <n> <mesh ...>
                    Sphere1DDustGrid* g = new Sphere1DDustGrid(); g->setMaxR(rmax); g->setMeshR(makeMesh(in, n));
                    S->grid = g; S->gridKind = 7;
                }
                else if (kind == "sphere2d")
                {
                    double rmax; int nr, nt; in >> rmax >> nr;      // grid sphere2d <maxR> <nr> <mesh ...> <ntheta> <mesh ...>
                    Sphere2DDustGrid* g = new Sphere2DDustGrid(); g->setMaxR(rmax); g->setMeshR(makeMesh(in, nr));
                    in >> nt; g->setMeshTheta(makeMesh(in, nt));
                    S->grid = g; S->gridKind = 8;
                }
                else if (kind == "cylinder2d")
                {
                    double Rmax, zmin, zmax; int nR, nz; in >> Rmax >> zmin >> zmax >> nR;     // grid cylinder2d <maxR> <minZ> <maxZ> <nR> <mesh ...> <nz> <mesh ...>
                    Cylinder2DDustGrid* g = new Cylinder2DDustGrid(); g->setMaxR(Rmax); g->setMinZ(zmin); g->setMaxZ(zmax); g->setMeshR(makeMesh(in, nR));
                    in >> nz; g->setMeshZ(makeMesh(in, nz));
                    S->grid = g; S->gridKind = 9;
                }
                else if (kind == "particletree")
                {
                    // grid particletree <oct|bin> <extraLevels>; the particles come through skr_set_particles
                    std::string tt; int extra; in >> tt >> extra;
                    ParticleTreeDustGrid* g = new ParticleTreeDustGrid(); setBox(g, S->box);
                    g->setTreeType(tt == "bin" ? ParticleTreeDustGrid::BinTree : ParticleTreeDustGrid::OctTree); g->setExtraLevels(extra);
                    S->grid = g; S->gridKind = tt == "bin" ? 6 : 5;
                    S->pdd = new ParticleCompDustDistribution(); S->cdd = S->pdd;
                }
                else throw std::runtime_error("unknown grid " + kind);
                S->grid->setWriteGrid(false);
            }
            else if (key == "stellar")
            {
                // stellar <L per wavelength given later through skr_set_luminosities> geometry...
                Geometry* g = makeGeometry(S, in); g = maybeSpiral(g, in);
                TableStellarComp* sc = new TableStellarComp(); sc->setGeometry(g);
                S->ss->insertComponent(S->stars.size(), sc); S->stars.push_back(sc);
            }
            else if (key == "dust")
            {
                // dust <tau_faceon> <lambda_norm> geometry...
                double tau, lam; in >> tau >> lam;
                Geometry* g = makeGeometry(S, in); g = maybeSpiral(g, in);
                if (!S->cdd) S->cdd = new CompDustDistribution();
                DustComp* dc = new DustComp(); dc->setGeometry(g);
                TableDustMix* mix = new TableDustMix(); dc->setMix(mix); S->mixes.push_back(mix);
                FaceOnDustCompNormalization* nrm = new FaceOnDustCompNormalization(); nrm->setWavelength(lam); nrm->setOpticalDepth(tau);
                dc->setNormalization(nrm);
                S->cdd->insertComponent(S->cdd->components().size(), dc);
            }
            else if (key == "dustmass")
            {
                // dustmass <total dust mass> geometry...   (DustMassDustCompNormalization: works for any geometry, also spherical ones)
                double mass; in >> mass;
                Geometry* g = makeGeometry(S, in); g = maybeSpiral(g, in);
                if (!S->cdd) S->cdd = new CompDustDistribution();
                DustComp* dc = new DustComp(); dc->setGeometry(g);
                TableDustMix* mix = new TableDustMix(); dc->setMix(mix); S->mixes.push_back(mix);
                DustMassDustCompNormalization* nrm = new DustMassDustCompNormalization(); nrm->setDustMass(mass);
                dc->setNormalization(nrm);
                S->cdd->insertComponent(S->cdd->components().size(), dc);
            }
            else if (key == "ameshdust")
            {
                // dust taken from the adaptive mesh itself: ameshdust <densityUnits>
                double units; in >> units;
                S->amdd = new AdaptiveMeshDustDistribution();
                S->amdd->setMinX(S->box[0]); S->amdd->setMaxX(S->box[1]); S->amdd->setMinY(S->box[2]);
                S->amdd->setMaxY(S->box[3]); S->amdd->setMinZ(S->box[4]); S->amdd->setMaxZ(S->box[5]);
                S->amdd->setDensityUnits(units);
                S->afile = new MemAdaptiveMeshFile(); S->amdd->setAdaptiveMeshFile(S->afile);
                MeshDustComponent* mdc = new MeshDustComponent(); mdc->setDensityIndex(0); mdc->setMultiplierIndex(-1); mdc->setDensityFraction(1.0);
                TableDustMix* mix = new TableDustMix(); mdc->setMix(mix); S->mixes.push_back(mix);
                S->amdd->insertComponent(0, mdc);
            }
            else if (key == "perspective")
            {
                // perspective <name> Nx Ny width Vx Vy Vz Cx Cy Cz Ux Uy Uz focal
                std::string name; int nx, ny; double w, v[10]; in >> name >> nx >> ny >> w; for (double& q : v) in >> q;
                PerspectiveInstrument* pi = new PerspectiveInstrument(); pi->setInstrumentName(QString(name));
                pi->setPixelsX(nx); pi->setPixelsY(ny); pi->setWidth(w); pi->setViewX(v[0]); pi->setViewY(v[1]); pi->setViewZ(v[2]);
                pi->setCrossX(v[3]); pi->setCrossY(v[4]); pi->setCrossZ(v[5]); pi->setUpX(v[6]); pi->setUpY(v[7]); pi->setUpZ(v[8]); pi->setFocal(v[9]);
                S->is->insertInstrument(S->is->instruments().size(), pi);
            }
            else if (key == "instrument")
            {
                std::string kind, name; double d, inc, az, pa; in >> kind >> name >> d >> inc >> az >> pa;
                DistantInstrument* di = 0;
                if (kind == "sed") di = new SEDInstrument();
                else if (kind == "multiframe")
                {
                    // instrument multiframe <name> d inc az pa <writeTotal> <writeStellarComps> <N>  then N x (nx fovx ny fovy xc yc)
                    int wt, wc, n; in >> wt >> wc >> n;
                    MultiFrameInstrument* mf = new MultiFrameInstrument(); mf->setWriteTotal(wt != 0); mf->setWriteStellarComps(wc != 0);
                    for (int q = 0; q < n; q++)
                    {
                        int nx, ny; double fx, fy, xc, yc; in >> nx >> fx >> ny >> fy >> xc >> yc;
                        InstrumentFrame* fr = new InstrumentFrame(); fr->setPixelsX(nx); fr->setFieldOfViewX(fx); fr->setPixelsY(ny); fr->setFieldOfViewY(fy);
                        fr->setCenterX(xc); fr->setCenterY(yc);
                        mf->insertFrame(q, fr);
                    }
                    di = mf;
                }
                else
                {
                    int nx, ny; double fx, fy; in >> nx >> fx >> ny >> fy;
                    SingleFrameInstrument* fi = (kind == "frame") ? (SingleFrameInstrument*)new FrameInstrument()
                                              : (kind == "full") ? (SingleFrameInstrument*)new FullInstrument() : (SingleFrameInstrument*)new SimpleInstrument();
                    if (kind == "full") { int nscatt; in >> nscatt; ((FullInstrument*)fi)->setScatteringLevels(nscatt); }
                    fi->setPixelsX(nx); fi->setFieldOfViewX(fx); fi->setPixelsY(ny); fi->setFieldOfViewY(fy); fi->setCenterX(0); fi->setCenterY(0);
                    di = fi;
                }
                di->setInstrumentName(QString(name)); di->setDistance(d); di->setInclination(inc); di->setAzimuth(az); di->setPositionAngle(pa);
                S->is->insertInstrument(S->is->instruments().size(), di);
            }
            else throw std::runtime_error("unknown spec key " + key);
        }

        // assemble the hierarchy the way the ski file would
        S->mc->parallelFactory()->setMaxThreadCount(threads);
        S->mc->random()->setSeed(seed);
        S->mc->setPackages(packages); S->mc->setMinWeightReduction(mwr); S->mc->setMinScattEvents(minscatt); S->mc->setScattBias(xi); S->mc->setContinuousScattering(contscatt != 0);
        S->ss->setEmissionBias(ebias);
        S->mc->setInstrumentSystem(S->is);
        Units* units = unitsys == 1 ? (Units*)new StellarUnits() : unitsys == 2 ? (Units*)new ExtragalacticUnits() : (Units*)new SIUnits();
        units->setFluxOutputStyle(fluxstyle == 1 ? Units::Wavelength : fluxstyle == 2 ? Units::Frequency : Units::Neutral);
        S->mc->setUnits(units);
        S->dd = S->amdd ? (DustDistribution*)S->amdd : (DustDistribution*)S->cdd;
        if (S->pan)
        {
            PanMonteCarloSimulation* p = (PanMonteCarloSimulation*)S->mc;
            if (!S->plg) throw std::runtime_error("pan simulation needs loggrid");
            p->setWavelengthGrid(S->plg); p->setStellarSystem(S->ss);
            if (S->dd)
            {
                PanDustSystem* ds = new PanDustSystem();
                ds->setDustDistribution(S->dd); ds->setDustGrid(S->grid); ds->setSampleCount(dustsamples);
                ds->setWriteConvergence(false); ds->setWriteDensity(false); ds->setWriteDepthMap(false);
                ds->setWriteQuality(false); ds->setWriteCellProperties(false); ds->setWriteCellsCrossed(false);
                // storeabs: PanDustSystem::storeabsorptionrates() == dustemission() (PanDustSystem.cpp:290-299), so the
                // absorption tables only exist with a dust emissivity + library attached
                if (storeabs) { ds->setDustEmissivity(new GreyBodyDustEmissivity()); ds->setDustLib(new AllCellsDustLib()); }
                ds->setSelfAbsorption(storeabs && selfabs); ds->setWriteEmissivity(false); ds->setWriteTemperature(false); ds->setWriteISRF(false);
                S->ds = ds; p->setDustSystem(ds);
            }
        }
        else
        {
            OligoMonteCarloSimulation* o = (OligoMonteCarloSimulation*)S->mc;
            if (!S->olg) throw std::runtime_error("oligo simulation needs wavelengths");
            o->setWavelengthGrid(S->olg); o->setStellarSystem(S->ss);
            if (S->dd)
            {
                OligoDustSystem* ds = new OligoDustSystem();
                ds->setDustDistribution(S->dd); ds->setDustGrid(S->grid); ds->setSampleCount(dustsamples);
                ds->setWriteConvergence(false); ds->setWriteDensity(false); ds->setWriteDepthMap(false);
                ds->setWriteQuality(false); ds->setWriteCellProperties(false); ds->setWriteCellsCrossed(false);
                ds->setWriteMeanIntensity(storeabs != 0);
                S->ds = ds; o->setDustSystem(ds);
            }
        }
    }

    template<class F> int guarded(F f)
    {
        try { f(); return 0; }
        catch (FatalError& e) { lastError = e.message().join(" | ").toStdString(); }
        catch (std::exception& e) { lastError = e.what(); }
        catch (...) { lastError = "unknown exception"; }
        return 1;
    }

    int numLambda(Sim* S) { return S->pan ? S->plg->Nlambda() : S->olg->Nlambda(); }
}

// captured by the Image / TextOutFile stand-ins (ref_stubs/services.cpp)
std::map<std::string, std::vector<double>>& skr_saved_images();
std::map<std::string, std::vector<std::vector<double>>>& skr_saved_rows();

extern "C"
{
const char* skr_error() { return lastError.c_str(); }
void skr_set_verbose(int v) { skr_verbose = v; }
long skr_warnings() { return skr_warning_count; }

void* skr_create(const char* spec)
{
    Sim* S = new Sim();
    if (guarded([&]{ build(S, spec); })) { return 0; }
    return S;
}
void skr_destroy(void* h) { Sim* S = (Sim*)h; if (S) { delete S->mc; delete S; } }

// inputs that are arrays ----------------------------------------------------------------------
int skr_set_luminosities(void* h, int comp, const double* L, int n)
{ Sim* S = (Sim*)h; if (comp < 0 || comp >= (int)S->stars.size()) return 1; S->stars[comp]->L.assign(L, L+n); return 0; }
int skr_set_mix(void* h, int comp, const double* kabs, const double* ksca, const double* g, int n)
{ Sim* S = (Sim*)h; if (comp < 0 || comp >= (int)S->mixes.size()) return 1;
  S->mixes[comp]->kabs.assign(kabs, kabs+n); S->mixes[comp]->ksca.assign(ksca, ksca+n); S->mixes[comp]->g.assign(g, g+n); return 0; }
// Mueller matrix coefficients of dust component `comp`: S11, S12, S33, S34 [Nlambda*Ntheta] (theta = t*pi/(Ntheta-1))
int skr_set_mueller(void* h, int comp, int Ntheta, int Nlambda, const double* S11, const double* S12, const double* S33, const double* S34)
{ Sim* S = (Sim*)h; if (comp < 0 || comp >= (int)S->mixes.size() || Ntheta < 2) return 1;
  size_t n = (size_t)Ntheta * Nlambda; TableDustMix* m = S->mixes[comp]; m->Ntheta = Ntheta;
  m->S11.assign(S11, S11+n); m->S12.assign(S12, S12+n); m->S33.assign(S33, S33+n); m->S34.assign(S34, S34+n); return 0; }
int skr_set_particles(void* h, const double* xyz, int n)
{ Sim* S = (Sim*)h; if (S->pdd) { S->pdd->xyz.assign(xyz, xyz+3*(size_t)n); return 0; }
  if (!S->vfile) return 1; S->vfile->xyz.assign(xyz, xyz+3*(size_t)n); return 0; }
int skr_set_amesh(void* h, const int* nxyz, const double* val, int n)
{ Sim* S = (Sim*)h; if (!S->afile) return 1; S->afile->nxyz.assign(nxyz, nxyz+3*(size_t)n); S->afile->val.assign(val, val+n); return 0; }

int skr_setup(void* h)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        S->mc->setup();
        // The reference never assigns AdaptiveMeshDustGrid::_random (initialised to 0 in AdaptiveMeshDustGrid.cpp:19 and used
        // by randomPositionInCell, :86-89), so its dust emission phases crash on adaptive mesh grids.  The harness supplies the
        // simulation's Random instance -- what the class evidently intends -- so that those phases can serve as a reference.
        if (AdaptiveMeshDustGrid* ag = dynamic_cast<AdaptiveMeshDustGrid*>(S->grid)) if (!ag->_random) ag->_random = S->mc->find<Random>();
    });
}

// the reference's InterstellarDustMix evaluated on this simulation's wavelength grid (needs dat/)
int skr_interstellar_mix(void* h, double* kabs, double* ksca, double* g)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        InterstellarDustMix* mix = new InterstellarDustMix(); mix->setParent(S->ds ? (QObject*)S->ds : (QObject*)S->mc); mix->setup();
        int n = numLambda(S);
        for (int ell = 0; ell < n; ell++) { kabs[ell] = mix->kappaabs(ell); ksca[ell] = mix->kappasca(ell); g[ell] = mix->_asymmparv[ell]; }
        delete mix;
    });
}

// scalar queries ------------------------------------------------------------------------------
int skr_num_lambda(void* h) { return numLambda((Sim*)h); }
int skr_num_cells(void* h) { Sim* S = (Sim*)h; return S->ds ? S->ds->Ncells() : 0; }
int skr_num_comp(void* h) { Sim* S = (Sim*)h; return S->ds ? S->ds->Ncomp() : 0; }
int skr_num_stellar(void* h) { return ((Sim*)h)->stars.size(); }
int skr_grid_kind(void* h) { return ((Sim*)h)->gridKind; }
double skr_packages_per_lambda(void* h) { Sim* S = (Sim*)h; S->mc->setChunkParams(S->mc->packages()); return S->mc->_Npp; }

void skr_get_lambda(void* h, double* lambda, double* dlambda)
{ Sim* S = (Sim*)h; WavelengthGrid* g = S->pan ? (WavelengthGrid*)S->plg : (WavelengthGrid*)S->olg;
  for (int i = 0; i < g->Nlambda(); i++) { lambda[i] = g->lambda(i); if (dlambda) dlambda[i] = g->dlambda(i); } }
void skr_get_rho(void* h, double* rho)
{ Sim* S = (Sim*)h; int N = S->ds->Ncells(), C = S->ds->Ncomp();
  for (int m = 0; m < N; m++) for (int c = 0; c < C; c++) rho[(size_t)m*C+c] = S->ds->density(m, c); }
void skr_get_volumes(void* h, double* vol) { Sim* S = (Sim*)h; for (int m = 0; m < S->ds->Ncells(); m++) vol[m] = S->ds->volume(m); }
void skr_get_opt(void* h, double* kext, double* ksca, double* g)
{ Sim* S = (Sim*)h; int C = S->ds->Ncomp(), L = numLambda(S);
  for (int c = 0; c < C; c++) for (int l = 0; l < L; l++)
  { DustMix* mix = S->ds->mix(c); kext[c*L+l] = mix->kappaext(l); ksca[c*L+l] = mix->kappasca(l); g[c*L+l] = mix->_asymmparv[l]; } }
void skr_get_albedo(void* h, double* alb)
{ Sim* S = (Sim*)h; int C = S->ds->Ncomp(), L = numLambda(S);
  for (int c = 0; c < C; c++) for (int l = 0; l < L; l++) alb[c*L+l] = S->ds->mix(c)->albedo(l); }
void skr_get_luminosities(void* h, double* L)   // [Nstellar*Nlambda]
{ Sim* S = (Sim*)h; int n = numLambda(S);
  for (size_t c = 0; c < S->stars.size(); c++) for (int l = 0; l < n; l++) L[c*n+l] = S->stars[c]->luminosity(l); }

// ---- cartesian --------------------------------------------------------------------------------
void skr_cart_dims(void* h, int* n) { CartesianDustGrid* g = (CartesianDustGrid*)((Sim*)h)->grid; n[0] = g->_Nx; n[1] = g->_Ny; n[2] = g->_Nz; }
void skr_cart_axes(void* h, double* xv, double* yv, double* zv)
{ CartesianDustGrid* g = (CartesianDustGrid*)((Sim*)h)->grid;
  for (int i = 0; i <= g->_Nx; i++) xv[i] = g->_xv[i]; for (int i = 0; i <= g->_Ny; i++) yv[i] = g->_yv[i]; for (int i = 0; i <= g->_Nz; i++) zv[i] = g->_zv[i]; }

// ---- grids with symmetries: the border arrays as setupSelfAfter left them ------------------------------
// sizes[0] = bins along r / R, sizes[1] = bins along theta / z (0 for the 1D grid)
void skr_sym_sizes(void* h, int* sizes)
{
    Sim* S = (Sim*)h; sizes[0] = sizes[1] = 0;
    if (S->gridKind == 7) sizes[0] = ((Sphere1DDustGrid*)S->grid)->_Nr;
    if (S->gridKind == 8) { Sphere2DDustGrid* g = (Sphere2DDustGrid*)S->grid; sizes[0] = g->_Nr; sizes[1] = g->_Ntheta; }
    if (S->gridKind == 9) { Cylinder2DDustGrid* g = (Cylinder2DDustGrid*)S->grid; sizes[0] = g->_NR; sizes[1] = g->_Nz; }
}
void skr_sym_tables(void* h, double* v1, double* v2, double* cv)
{
    Sim* S = (Sim*)h;
    if (S->gridKind == 7) { Sphere1DDustGrid* g = (Sphere1DDustGrid*)S->grid; for (int i = 0; i <= g->_Nr; i++) v1[i] = g->_rv[i]; }
    if (S->gridKind == 8)
    { Sphere2DDustGrid* g = (Sphere2DDustGrid*)S->grid; for (int i = 0; i <= g->_Nr; i++) v1[i] = g->_rv[i];
      for (int k = 0; k <= g->_Ntheta; k++) { v2[k] = g->_thetav[k]; cv[k] = g->_cv[k]; } }
    if (S->gridKind == 9)
    { Cylinder2DDustGrid* g = (Cylinder2DDustGrid*)S->grid; for (int i = 0; i <= g->_NR; i++) v1[i] = g->_Rv[i]; for (int k = 0; k <= g->_Nz; k++) v2[k] = g->_zv[k]; }
}

// ---- tree -------------------------------------------------------------------------------------
// the node vector, the cell numbers and eps of a TreeDustGrid or a ParticleTreeDustGrid
static void treeParts(Sim* S, std::vector<TreeNode*>*& tree, std::vector<int>*& cellnumber, double& eps)
{
    if (S->gridKind >= 5) { ParticleTreeDustGrid* g = (ParticleTreeDustGrid*)S->grid; tree = &g->_tree; cellnumber = &g->_cellnumberv; eps = g->_eps; }
    else { TreeDustGrid* g = (TreeDustGrid*)S->grid; tree = &g->_tree; cellnumber = &g->_cellnumberv; eps = g->_eps; }
}
void skr_tree_sizes(void* h, int* nnodes, int* nnbr, double* eps)
{
    std::vector<TreeNode*>* tree; std::vector<int>* cn; treeParts((Sim*)h, tree, cn, *eps);
    *nnodes = (int)tree->size(); long total = 0;
    for (TreeNode* n : *tree) for (auto& v : n->_neighbors) total += v.size();
    *nnbr = (int)total;
}
// box[6N] = xmin,ymin,zmin,xmax,ymax,zmax; child0[N]; parent[N]; cell[N]; dir[N]; nbrStart[6N+1]; nbrIds[]
int skr_tree_tables(void* h, double* box, int* child0, int* parent, int* cell, int* dir, int* nbrStart, int* nbrIds)
{
    Sim* S = (Sim*)h; std::vector<TreeNode*>* tree; std::vector<int>* cn; double eps; treeParts(S, tree, cn, eps);
    int N = (int)tree->size(); int pos = 0;
    for (int l = 0; l < N; l++)
    {
        TreeNode* n = (*tree)[l];
        if (n->id() != l) { lastError = "node id mismatch"; return 1; }
        box[6*l+0] = n->xmin(); box[6*l+1] = n->ymin(); box[6*l+2] = n->zmin();
        box[6*l+3] = n->xmax(); box[6*l+4] = n->ymax(); box[6*l+5] = n->zmax();
        parent[l] = n->father() ? n->father()->id() : -1;
        cell[l] = (*cn)[l];
        child0[l] = n->ynchildless() ? -1 : n->child(0)->id();
        if (!n->ynchildless())
            for (size_t c = 0; c < n->children().size(); c++)
                if (n->child(c)->id() != child0[l] + (int)c) { lastError = "children ids not consecutive"; return 1; }
        BinTreeNode* b = dynamic_cast<BinTreeNode*>(n);
        dir[l] = b ? b->_dir : 0;
        for (int w = 0; w < 6; w++)
        {
            nbrStart[6*l+w] = pos;
            if ((int)n->_neighbors.size() == 6) for (TreeNode* q : n->_neighbors[w]) nbrIds[pos++] = q->id();
        }
    }
    nbrStart[6*N] = pos;
    return 0;
}

// ---- adaptive mesh ----------------------------------------------------------------------------
static void amFlatten(Sim* S)
{
    if (!S->amNodes.empty()) return;
    AdaptiveMesh* am = ((AdaptiveMeshDustGrid*)S->grid)->_amesh;
    // breadth-first numbering so that the children of a node get consecutive indices in local Morton order
    S->amNodes.push_back(am->_root); S->amIndex[am->_root] = 0;
    for (size_t q = 0; q < S->amNodes.size(); q++)
    {
        const AdaptiveMeshNode* n = S->amNodes[q];
        if (!n->isLeaf()) for (const AdaptiveMeshNode* c : n->_nodes) { S->amIndex[c] = S->amNodes.size(); S->amNodes.push_back(c); }
    }
}
void skr_amesh_sizes(void* h, int* nnodes, double* eps)
{ Sim* S = (Sim*)h; amFlatten(S); *nnodes = S->amNodes.size(); *eps = ((AdaptiveMeshDustGrid*)S->grid)->_amesh->_eps; }
// box[6N]; nxyz[3N] (0 for leaves); child0[N] (-1 for leaves); cell[N]; wallNbr[6N] (-1: none)
void skr_amesh_tables(void* h, double* box, int* nxyz, int* child0, int* cell, int* wallNbr)
{
    Sim* S = (Sim*)h; amFlatten(S);
    for (size_t l = 0; l < S->amNodes.size(); l++)
    {
        const AdaptiveMeshNode* n = S->amNodes[l];
        box[6*l+0] = n->xmin(); box[6*l+1] = n->ymin(); box[6*l+2] = n->zmin();
        box[6*l+3] = n->xmax(); box[6*l+4] = n->ymax(); box[6*l+5] = n->zmax();
        nxyz[3*l] = n->_Nx; nxyz[3*l+1] = n->_Ny; nxyz[3*l+2] = n->_Nz;
        cell[l] = n->_m;
        child0[l] = n->isLeaf() ? -1 : S->amIndex[n->_nodes[0]];
        for (int w = 0; w < 6; w++)
            wallNbr[6*l+w] = (n->isLeaf() && n->_nodes.size() == 6 && n->_nodes[w]) ? S->amIndex[n->_nodes[w]] : -1;
    }
}

// ---- voronoi ----------------------------------------------------------------------------------
static VoronoiMesh* vmesh(Sim* S) { return ((VoronoiDustGrid*)S->grid)->_mesh; }
static void kdFlatten(Sim* S)
{
    if (!S->kdNodes.empty()) return;
    VoronoiMesh* vm = vmesh(S);
    std::function<void(VoronoiMesh_Private::Node*)> visit = [&](VoronoiMesh_Private::Node* n)
    { if (!n) return; S->kdIndex[n] = S->kdNodes.size(); S->kdNodes.push_back(n); visit(n->_left); visit(n->_right); };
    for (auto* t : vm->_blocktrees) visit(t);
}
// sizes[0]=Ncells [1]=total neighbours [2]=nb [3]=total block refs [4]=kd nodes
void skr_voro_sizes(void* h, long* sizes, double* eps)
{
    Sim* S = (Sim*)h; VoronoiMesh* vm = vmesh(S); kdFlatten(S);
    long nn = 0; for (auto* c : vm->_cells) nn += c->_neighbors.size();
    long nr = 0; for (auto& b : vm->_blocklists) nr += b.size();
    sizes[0] = vm->_Ncells; sizes[1] = nn; sizes[2] = vm->_nb; sizes[3] = nr; sizes[4] = S->kdNodes.size(); *eps = vm->_eps;
}
void skr_voro_tables(void* h, double* particles, double* cellBox, int* nbrStart, int* nbrIds, int* blkStart, int* blkIds,
                     int* blkTree, int* kdM, int* kdAxis, int* kdUp, int* kdLeft, int* kdRight)
{
    Sim* S = (Sim*)h; VoronoiMesh* vm = vmesh(S); kdFlatten(S);
    int pos = 0;
    for (int m = 0; m < vm->_Ncells; m++)
    {
        auto* c = vm->_cells[m];
        particles[3*m] = c->_r.x(); particles[3*m+1] = c->_r.y(); particles[3*m+2] = c->_r.z();
        cellBox[6*m] = c->xmin(); cellBox[6*m+1] = c->ymin(); cellBox[6*m+2] = c->zmin();
        cellBox[6*m+3] = c->xmax(); cellBox[6*m+4] = c->ymax(); cellBox[6*m+5] = c->zmax();
        nbrStart[m] = pos; for (int id : c->_neighbors) nbrIds[pos++] = id;
    }
    nbrStart[vm->_Ncells] = pos;
    pos = 0;
    for (int b = 0; b < vm->_nb3; b++)
    {
        blkStart[b] = pos; for (int id : vm->_blocklists[b]) blkIds[pos++] = id;
        blkTree[b] = vm->_blocktrees[b] ? S->kdIndex[vm->_blocktrees[b]] : -1;
    }
    blkStart[vm->_nb3] = pos;
    for (size_t i = 0; i < S->kdNodes.size(); i++)
    {
        auto* n = S->kdNodes[i];
        kdM[i] = n->_m; kdAxis[i] = n->_axis;
        kdUp[i] = n->_up ? S->kdIndex[n->_up] : -1; kdLeft[i] = n->_left ? S->kdIndex[n->_left] : -1; kdRight[i] = n->_right ? S->kdIndex[n->_right] : -1;
    }
}

// ---- deterministic geometry: DustGrid::path + fillOpticalDepth for fixed rays --------------------
// ell < 0: geometry only (dtau/tau left zero).  Returns the total number of segments; arrays are
// filled only while they fit in `cap` (call once with cap = 0 to size them).  Multi-threaded over rays.
long skr_path_batch(void* h, const double* r, const double* k, long n, int ell, long cap,
                    long* offsets, int* m, double* ds, double* s, double* dtau, double* tau, int nthreads)
{
    Sim* S = (Sim*)h;
    // pass 1: counts (also used for the CSR offsets)
    std::vector<int> counts(n);
    auto work = [&](int pass, int tid, int nt)
    {
        DustGridPath path;
        for (long i = tid; i < n; i += nt)
        {
            path.setPosition(Position(r[3*i], r[3*i+1], r[3*i+2]));
            path.setDirection(Direction(k[3*i], k[3*i+1], k[3*i+2]));
            S->grid->path(&path);
            if (pass == 0) { counts[i] = path.size(); continue; }
            if (ell >= 0 && S->ds)
            {
                // KappaRho functor of DustSystem.cpp:465-491 through the public accessors
                DustSystem* dsys = S->ds; int C = dsys->Ncomp();
                path.fillOpticalDepth([dsys, C, ell](int mm)
                { double res = 0; for (int c = 0; c < C; c++) res += dsys->mix(c)->kappaext(ell) * dsys->density(mm, c); return res; });
            }
            long o = offsets[i];
            if (o + path.size() > cap) continue;
            for (int j = 0; j < path.size(); j++)
            { m[o+j] = path.m(j); ds[o+j] = path.ds(j); s[o+j] = path.s(j); dtau[o+j] = path.dtau(j); tau[o+j] = path.tau(j); }
        }
    };
    auto run = [&](int pass)
    {
        int nt = std::max(1, nthreads);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, pass, t, nt);
        work(pass, 0, nt);
        for (auto& t : th) t.join();
    };
    int rc = guarded([&]{
        run(0);
        long total = 0; for (long i = 0; i < n; i++) { offsets[i] = total; total += counts[i]; } offsets[n] = total;
        if (cap > 0) run(1);
    });
    return rc ? -1 : offsets[n];
}

int skr_whichcell(void* h, const double* r, long n, int* m)
{ Sim* S = (Sim*)h; return guarded([&]{ for (long i = 0; i < n; i++) m[i] = S->grid->whichcell(Position(r[3*i], r[3*i+1], r[3*i+2])); }); }

// DustSystem::opticaldepth(pp, distance) for fixed rays (Instrument::opticalDepth, Instrument.cpp:69-72)
int skr_opticaldepth_batch(void* h, const double* r, const double* k, long n, int ell, const double* distance, double* tau)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        PhotonPackage pp;
        for (long i = 0; i < n; i++)
        {
            pp.launch(1.0, ell, Position(r[3*i], r[3*i+1], r[3*i+2]), Direction(k[3*i], k[3*i+1], k[3*i+2]));
            tau[i] = S->ds->opticaldepth(&pp, distance ? distance[i] : DBL_MAX);
        }
    });
}

// ---- Monte Carlo -------------------------------------------------------------------------------
// zero the detector arrays / absorption tables and reseed, so that independent batches can be run
int skr_reset(void* h, int seed)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        for (Instrument* ins : S->is->instruments())
        {
            if (FrameInstrument* f = dynamic_cast<FrameInstrument*>(ins)) f->_ftotv = 0.0;
            if (PerspectiveInstrument* f = dynamic_cast<PerspectiveInstrument*>(ins)) f->_ftotv = 0.0;
            if (SEDInstrument* f = dynamic_cast<SEDInstrument*>(ins)) f->_Ftotv = 0.0;
            if (SimpleInstrument* f = dynamic_cast<SimpleInstrument*>(ins)) { f->_ftotv = 0.0; f->_Ftotv = 0.0; }
            if (MultiFrameInstrument* mf = dynamic_cast<MultiFrameInstrument*>(ins))
                for (InstrumentFrame* fr : mf->_frames) { if (fr->_ftotv.size()) fr->_ftotv = 0.0; for (size_t k = 0; k < fr->_fcompvv.size(0); k++) fr->_fcompvv[k] = 0.0; }
            if (FullInstrument* f = dynamic_cast<FullInstrument*>(ins))
            {
                for (Array* a : {&f->_ftrav, &f->_Ftrav, &f->_fstrdirv, &f->_Fstrdirv, &f->_fstrscav, &f->_Fstrscav, &f->_fdusdirv, &f->_Fdusdirv,
                                 &f->_fdusscav, &f->_Fdusscav}) if (a->size()) *a = 0.0;
                for (int n = 0; n < f->_Nscatt; n++) { f->_fstrscavv[n] = 0.0; f->_Fstrscavv[n] = 0.0; }
                for (Array* a : {&f->_ftotQv, &f->_FtotQv, &f->_ftotUv, &f->_FtotUv, &f->_ftotVv, &f->_FtotVv}) if (a->size()) *a = 0.0;
            }
        }
        if (S->ds)
        {
            if (OligoDustSystem* o = dynamic_cast<OligoDustSystem*>(S->ds)) { if (o->_Labsvv.size(0)) o->_Labsvv.clear(); }
            if (PanDustSystem* p = dynamic_cast<PanDustSystem*>(S->ds)) { if (p->_Labsstelvv.size(0)) p->_Labsstelvv.clear(); if (p->_Labsdustvv.size(0)) p->_Labsdustvv.clear(); }
        }
        S->mc->random()->setSeed(seed);
        S->mc->random()->initialize(S->mc->parallelFactory()->maxThreadCount());
    });
}
void skr_set_packages(void* h, double packages) { ((Sim*)h)->mc->setPackages(packages); }

// the stellar emission phase exactly as the reference drives it; returns wall-clock seconds
int skr_run_stellar(void* h, double* seconds)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        auto t0 = std::chrono::steady_clock::now();
        S->mc->runstellaremission();
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}

int skr_num_instruments(void* h) { return ((Sim*)h)->is->instruments().size(); }
// sizes: frame pixels*Nlambda (0 if none), sed Nlambda (0 if none)
void skr_instrument_sizes(void* h, int i, long* nframe, long* nsed)
{
    Instrument* ins = ((Sim*)h)->is->instruments()[i]; *nframe = 0; *nsed = 0;
    if (FrameInstrument* f = dynamic_cast<FrameInstrument*>(ins)) *nframe = f->_ftotv.size();
    if (PerspectiveInstrument* f = dynamic_cast<PerspectiveInstrument*>(ins)) *nframe = f->_ftotv.size();
    if (SEDInstrument* f = dynamic_cast<SEDInstrument*>(ins)) *nsed = f->_Ftotv.size();
    if (SimpleInstrument* f = dynamic_cast<SimpleInstrument*>(ins)) { *nframe = f->_ftotv.size(); *nsed = f->_Ftotv.size(); }
}
void skr_get_instrument(void* h, int i, double* frame, double* sed)
{
    Instrument* ins = ((Sim*)h)->is->instruments()[i];
    const Array* fa = 0; const Array* sa = 0;
    if (FrameInstrument* f = dynamic_cast<FrameInstrument*>(ins)) fa = &f->_ftotv;
    if (PerspectiveInstrument* f = dynamic_cast<PerspectiveInstrument*>(ins)) fa = &f->_ftotv;
    if (SEDInstrument* f = dynamic_cast<SEDInstrument*>(ins)) sa = &f->_Ftotv;
    if (SimpleInstrument* f = dynamic_cast<SimpleInstrument*>(ins)) { fa = &f->_ftotv; sa = &f->_Ftotv; }
    if (fa && frame) for (size_t j = 0; j < fa->size(); j++) frame[j] = (*fa)[j];
    if (sa && sed) for (size_t j = 0; j < sa->size(); j++) sed[j] = (*sa)[j];
}
// raw arrays of one FullInstrument channel (0 transparent, 1 direct, 2 scattered, 3 dust direct, 4 dust scattered, 5+n level n+1)
int skr_get_full_channel(void* h, int i, int c, double* frame, double* sed)
{
    FullInstrument* f = dynamic_cast<FullInstrument*>(((Sim*)h)->is->instruments()[i]);
    if (!f) return 1;
    const Array* fa = 0; const Array* sa = 0;
    switch (c)
    {
    case 0: fa = &f->_ftrav; sa = &f->_Ftrav; break;
    case 1: fa = &f->_fstrdirv; sa = &f->_Fstrdirv; break;
    case 2: fa = &f->_fstrscav; sa = &f->_Fstrscav; break;
    case 3: fa = &f->_fdusdirv; sa = &f->_Fdusdirv; break;
    case 4: fa = &f->_fdusscav; sa = &f->_Fdusscav; break;
    default:
        if (c - 5 < f->_Nscatt) { fa = &f->_fstrscavv[c - 5]; sa = &f->_Fstrscavv[c - 5]; break; }
        // the Stokes Q, U, V arrays of a simulation with polarisation follow the scattering levels (FullInstrument.cpp:79-87)
        if (!f->_polarization || c - 5 - f->_Nscatt > 2) return 1;
        if (c - 5 - f->_Nscatt == 0) { fa = &f->_ftotQv; sa = &f->_FtotQv; }
        else if (c - 5 - f->_Nscatt == 1) { fa = &f->_ftotUv; sa = &f->_FtotUv; }
        else { fa = &f->_ftotVv; sa = &f->_FtotVv; }
    }
    if (frame) for (size_t j = 0; j < fa->size(); j++) frame[j] = (*fa)[j];
    if (sed) for (size_t j = 0; j < sa->size(); j++) sed[j] = (*sa)[j];
    return 0;
}
// raw array of one frame of a MultiFrameInstrument: which = -1 total, k >= 0 stellar component k; returns the number of pixels (out may be null)
long skr_get_multiframe(void* h, int i, int which, int ell, double* out)
{
    MultiFrameInstrument* mf = dynamic_cast<MultiFrameInstrument*>(((Sim*)h)->is->instruments()[i]);
    if (!mf || ell < 0 || ell >= mf->_frames.size()) return -1;
    InstrumentFrame* fr = mf->_frames[ell];
    const Array* a = which < 0 ? &fr->_ftotv : (which < (int)fr->_fcompvv.size(0) ? &fr->_fcompvv[which] : 0);
    if (!a) return -1;
    if (out) for (size_t j = 0; j < a->size(); j++) out[j] = (*a)[j];
    return (long)a->size();
}
// instrument geometry as the reference derived it (DistantInstrument.cpp:27-50, SingleFrameInstrument.cpp)
void skr_get_instrument_geometry(void* h, int i, double* out /*[16]*/)
{
    Instrument* ins = ((Sim*)h)->is->instruments()[i];
    DistantInstrument* d = dynamic_cast<DistantInstrument*>(ins);
    if (!d) { for (int j = 0; j < 16; j++) out[j] = 0; return; }
    out[0] = d->_costheta; out[1] = d->_sintheta; out[2] = d->_cosphi; out[3] = d->_sinphi; out[4] = d->_cospa; out[5] = d->_sinpa;
    out[6] = d->_bfkobs.kx(); out[7] = d->_bfkobs.ky(); out[8] = d->_bfkobs.kz();
    SingleFrameInstrument* f = dynamic_cast<SingleFrameInstrument*>(ins);
    if (f) { out[9] = f->_Nxp; out[10] = f->_Nyp; out[11] = f->_xpmin; out[12] = f->_ypmin; out[13] = f->_xpsiz; out[14] = f->_ypsiz; }
    else for (int j = 9; j < 15; j++) out[j] = 0;
    out[15] = d->_distance;
}
int skr_get_labs(void* h, double* labs)   // [Ncells*Nlambda], stellar absorption
{
    Sim* S = (Sim*)h; if (!S->ds || !S->ds->storeabsorptionrates()) return 1;
    int N = S->ds->Ncells(), L = numLambda(S);
    for (int m = 0; m < N; m++) for (int l = 0; l < L; l++) labs[(size_t)m*L+l] = S->ds->Labs(m, l);
    return 0;
}

// ---- dust emission phases (PanMonteCarloSimulation.cpp:105-342) ---------------------------------------------
// PanDustSystem::calculatedustemission (sumResults + DustLib::calculate) and the bolometric absorbed luminosities,
// exactly as rundustselfabsorption :131-136 / rundustemission :250-255 prepare a shooting phase; returns the
// vectors Lv of dodust*chunk (:193-198, :275-280) for all wavelengths: Lv[ell*Ncells + m]
int skr_prepare_dust(void* h, int ynstellar, double* Lv)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        PanMonteCarloSimulation* p = dynamic_cast<PanMonteCarloSimulation*>(S->mc);
        if (!p || !p->_pds || !p->_pds->dustemission()) throw std::runtime_error("not a panchromatic simulation with dust emission");
        p->_pds->calculatedustemission(ynstellar != 0);
        int N = p->_Ncells, L = numLambda(S);
        for (int m = 0; m < N; m++) p->_Labsbolv[m] = p->_pds->Labs(m);
        if (Lv) for (int ell = 0; ell < L; ell++) for (int m = 0; m < N; m++)
        {
            double Labsbol = p->_Labsbolv[m];
            Lv[(size_t)ell*N + m] = Labsbol > 0.0 ? Labsbol * p->_pds->dustluminosity(m, ell) : 0.0;
        }
    });
}
// one dust self-absorption cycle (selfabs != 0; :139-148) or the dust emission phase (:258-264) with
// packages*factor packets per wavelength; skr_prepare_dust must have been called
int skr_run_dust(void* h, int selfabs, double factor, double* seconds)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        PanMonteCarloSimulation* p = dynamic_cast<PanMonteCarloSimulation*>(S->mc);
        if (!p || !p->_pds || !p->_pds->dustemission()) throw std::runtime_error("not a panchromatic simulation with dust emission");
        auto t0 = std::chrono::steady_clock::now();
        if (selfabs) p->_pds->rebootLabsdust();
        p->setChunkParams(p->packages() * factor);
        p->initprogress(selfabs ? "dust self-absorption cycle" : "dust emission");
        Parallel* parallel = p->find<ParallelFactory>()->parallel();
        if (selfabs) parallel->call(p, &PanMonteCarloSimulation::dodustselfabsorptionchunk, p->assigner());
        else parallel->call(p, &PanMonteCarloSimulation::dodustemissionchunk, p->assigner());
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
}
int skr_get_labs_dust(void* h, double* labs)   // [Ncells*Nlambda], absorbed dust emission
{
    Sim* S = (Sim*)h; PanDustSystem* p = dynamic_cast<PanDustSystem*>(S->ds);
    if (!p || !p->_haveLabsdust) return 1;
    int N = p->Ncells(), L = numLambda(S);
    for (int m = 0; m < N; m++) for (int l = 0; l < L; l++) labs[(size_t)m*L+l] = p->_Labsdustvv(m, l);
    return 0;
}
int skr_get_labs_bol(void* h, double* out)     // PanDustSystem::Labs(m)
{
    Sim* S = (Sim*)h; PanDustSystem* p = dynamic_cast<PanDustSystem*>(S->ds);
    if (!p) return 1;
    for (int m = 0; m < p->Ncells(); m++) out[m] = p->Labs(m);
    return 0;
}
// DustGrid::randomPositionInCell driven by the simulation's Random (thread 0)
int skr_random_positions(void* h, int m, long n, double* xyz)
{
    Sim* S = (Sim*)h;
    return guarded([&]{ for (long i = 0; i < n; i++) { Position r = S->grid->randomPositionInCell(m); xyz[3*i] = r.x(); xyz[3*i+1] = r.y(); xyz[3*i+2] = r.z(); } });
}

// ---- output: Instrument::write() = sumResults + calibration (SingleFrameInstrument.cpp:151-226, DistantInstrument.cpp:131-183);
// the calibrated data cubes / SED rows are captured by the Image / TextOutFile stand-ins of ref_stubs
int skr_write_instruments(void* h)
{
    Sim* S = (Sim*)h;
    return guarded([&]{ for (Instrument* ins : S->is->instruments()) ins->write(); });
}
long skr_saved_image(const char* name, double* out, long cap)
{
    auto it = skr_saved_images().find(name);
    if (it == skr_saved_images().end()) return -1;
    if (out) for (long i = 0; i < cap && i < (long)it->second.size(); i++) out[i] = it->second[i];
    return (long)it->second.size();
}
long skr_saved_table(const char* name, double* out, long cap, int* ncols)
{
    auto it = skr_saved_rows().find(name);
    if (it == skr_saved_rows().end()) return -1;
    long n = 0; *ncols = it->second.empty() ? 0 : (int)it->second[0].size();
    for (auto& row : it->second) for (double v : row) { if (out && n < cap) out[n] = v; n++; }
    return n;
}

// ---- samplers exposed for distribution-level checks ------------------------------------------------
// draws n launches from the stellar system at wavelength ell with unit luminosity: r[3n], k[3n], L[n]
int skr_sample_launch(void* h, int ell, long n, double* r, double* k, double* L)
{
    Sim* S = (Sim*)h;
    return guarded([&]{
        PhotonPackage pp;
        for (long i = 0; i < n; i++)
        {
            S->ss->launch(&pp, ell, 1.0);
            r[3*i] = pp.position().x(); r[3*i+1] = pp.position().y(); r[3*i+2] = pp.position().z();
            k[3*i] = pp.direction().kx(); k[3*i+1] = pp.direction().ky(); k[3*i+2] = pp.direction().kz();
            L[i] = pp.luminosity();
        }
    });
}
// raw uniform deviates of thread 0 (MT19937 stream, Random.cpp:89-126)
void skr_uniforms(void* h, long n, double* u) { Sim* S = (Sim*)h; for (long i = 0; i < n; i++) u[i] = S->mc->random()->uniform(); }
}
