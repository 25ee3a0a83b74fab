// Command-line driver of the C++ host layer: builds a simulation hierarchy from a small parameter file (one item
// per line, the same vocabulary as the ski properties it stands for), runs the stellar emission phase on the GPU
// and writes the raw detector arrays plus what Instrument::write() produces in the reference: the calibrated FITS data
// cubes and SED text files (Output.hpp).  ski/XML parsing stays in the reference (out of scope).
// With --write-only the shooting is skipped: the raw arrays <prefix>_<name>_frame.f64 / _sed.f64 of an earlier run are
// read back and only the calibration + output step runs (no GPU needed).
//
//   sim oligo|pan ; packages N ; seed S ; minweightreduction f ; minscatt n ; scattbias xi ; emissionbias xi
//   wavelengths l1 l2 ...          | loggrid min max points
//   box xmin xmax ymin ymax zmin zmax
//   grid cartesian nx ny nz lin|pow r|sympow r  (x3) | grid octtree|bintree minLevel maxLevel search maxMassFraction [samples [maxTau]]
//        | grid amesh <file> densityUnits [index] (+ meshdust) | grid voronoi <particle file> | grid particletree oct|bin <particle file> [extraLevels]
//   dustmix interstellar <file> | dustmix table kabs ksca g      (one wavelength)
//   dust tau lambda expdisk hR hz Rmax zmax
//   stellar L1[,L2,...]|bb:T:Lbol expdisk hR hz Rmax zmax | sersic n Reff q
//   instrument frame|sed|simple name distance inclination azimuth pa [nx fovx ny fovy]
//   storeabs 0|1 ; device d ; lattice n ; dustemission 0|1 ; selfabs 0|1 ; cycles n (0: until convergence)
//   units si|stellar|extragalactic neutral|wavelength|frequency
#include <cstdio>
#include <cstring>
#include <iostream>
#include "SimulationItems.hpp"
#include "Output.hpp"

using namespace skirt;

static Mesh* makeMesh(std::istringstream& in, int n)
{
    std::string kind; in >> kind; Mesh* m = nullptr;
    if (kind == "lin") m = new LinMesh();
    else if (kind == "pow") { auto* p = new PowMesh(); double r; in >> r; p->setRatio(r); m = p; }
    else if (kind == "sympow") { auto* p = new SymPowMesh(); double r; in >> r; p->setRatio(r); m = p; }
    else if (kind == "log") { auto* p = new LogMesh(); double tc; in >> tc; p->setCentralBinFraction(tc); m = p; }
    else SKIRT_FATAL("unknown mesh " + kind);
    m->setNumBins(n);
    return m;
}

static Geometry* makeGeometry(std::istringstream& in)
{
    std::string kind; in >> kind;
    if (kind == "expdisk")
    {
        double hR, hz, Rmax, zmax; in >> hR >> hz >> Rmax >> zmax;
        auto* g = new ExpDiskGeometry(); g->setRadialScale(hR); g->setAxialScale(hz); g->setRadialTrunc(Rmax); g->setAxialTrunc(zmax);
        return g;
    }
    if (kind == "sersic")
    {
        double n, Re, q; in >> n >> Re >> q;
        auto* g = new SersicGeometry(); g->setIndex(n); g->setRadius(Re); g->setFlattening(q);
        return g;
    }
    SKIRT_FATAL("unknown geometry " + kind);
}

int main(int argc, char** argv)
{
    bool writeOnly = argc > 1 && std::string(argv[1]) == "--write-only";
    if (writeOnly) { argv++; argc--; }
    if (argc < 3) { std::fprintf(stderr, "usage: %s [--write-only] <parameter file> <output prefix>\n", argv[0]); return 2; }
    try
    {
        std::ifstream file(argv[1]);
        if (!file) SKIRT_FATAL(std::string("cannot open ") + argv[1]);
        std::unique_ptr<UnitSystem> units(new SIUnits());
        MonteCarloSimulation sim;
        auto* ss = new StellarSystem(); auto* ds = new DustSystem(); auto* is = new InstrumentSystem();
        double box[6] = {0, 0, 0, 0, 0, 0};
        std::string mixKind = "table", mixFile; double mixv[3] = {0, 0, 0};
        bool pan = false;
        std::string line;
        while (std::getline(file, line))
        {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream in(line); std::string key; in >> key;
            if (key == "sim") { std::string t; in >> t; pan = t == "pan"; }
            else if (key == "packages") { double v; in >> v; sim.setPackages(v); }
            else if (key == "seed") { int v; in >> v; sim.setSeed(v); }
            else if (key == "device") { int v; in >> v; sim.setDevice(v); }
            else if (key == "minweightreduction") { double v; in >> v; sim.setMinWeightReduction(v); }
            else if (key == "minscatt") { double v; in >> v; sim.setMinScattEvents(v); }
            else if (key == "scattbias") { double v; in >> v; sim.setScattBias(v); }
            else if (key == "emissionbias") { double v; in >> v; ss->setEmissionBias(v); }
            else if (key == "storeabs") { int v; in >> v; ds->setStoreAbsorptionRates(v != 0); }
            else if (key == "lattice") { int v; in >> v; ds->setSampleLattice(v); }
            else if (key == "continuousscattering") { int v; in >> v; sim.setContinuousScattering(v != 0); }
            else if (key == "dustemission") { int v; in >> v; sim.setDustEmission(v != 0); }
            else if (key == "selfabs") { int v; in >> v; sim.setSelfAbsorption(v != 0); }
            else if (key == "cycles") { int v; in >> v; sim.setCycles(v); }
            else if (key == "units")
            {
                std::string sys, style; in >> sys >> style;
                if (sys == "si") units.reset(new SIUnits()); else if (sys == "stellar") units.reset(new StellarUnits());
                else if (sys == "extragalactic") units.reset(new ExtragalacticUnits()); else SKIRT_FATAL("unknown unit system " + sys);
                if (style == "wavelength") units->setFluxOutputStyle(UnitSystem::Wavelength); else if (style == "frequency") units->setFluxOutputStyle(UnitSystem::Frequency);
                else if (style != "neutral" && !style.empty()) SKIRT_FATAL("Unknown flux output style " + style);
            }
            else if (key == "wavelengths") { std::vector<double> lv; double v; while (in >> v) lv.push_back(v); auto* g = new OligoWavelengthGrid(); g->setWavelengths(lv); sim.setWavelengthGrid(g); }
            else if (key == "loggrid") { double a, b; int n; in >> a >> b >> n; auto* g = new LogWavelengthGrid(); g->setMinWavelength(a); g->setMaxWavelength(b); g->setPoints(n); sim.setWavelengthGrid(g); }
            else if (key == "filegrid") { std::string f; in >> f; auto* g = new FileWavelengthGrid(); g->setFilename(f); sim.setWavelengthGrid(g); }
            else if (key == "nestedloggrid")
            {
                double a, b, za, zb; int n, zn; in >> a >> b >> n >> za >> zb >> zn;
                auto* g = new NestedLogWavelengthGrid(); g->setMinWavelength(a); g->setMaxWavelength(b); g->setPoints(n);
                g->setMinWavelengthSubGrid(za); g->setMaxWavelengthSubGrid(zb); g->setPointsSubGrid(zn); sim.setWavelengthGrid(g);
            }
            else if (key == "box") { for (double& v : box) in >> v; }
            else if (key == "grid")
            {
                std::string kind; in >> kind;
                auto setBox = [&](BoxDustGrid* g) { g->setMinX(box[0]); g->setMaxX(box[1]); g->setMinY(box[2]); g->setMaxY(box[3]); g->setMinZ(box[4]); g->setMaxZ(box[5]); };
                if (kind == "octtree" || kind == "bintree")
                {
                    // grid octtree|bintree minLevel maxLevel search(0 TopDown, 1 Neighbor, 2 Bookkeeping) maxMassFraction [sampleCount [maxOpticalDepth [maxDensDispFraction]]]
                    int minl, maxl, search; double mf; int samples = 100; double maxtau = 0, maxdisp = 0;
                    in >> minl >> maxl >> search >> mf; in >> samples >> maxtau >> maxdisp;
                    TreeDustGrid* g = kind == "octtree" ? (TreeDustGrid*)new OctTreeDustGrid() : (TreeDustGrid*)new BinTreeDustGrid();
                    setBox(g); g->setMinLevel(minl); g->setMaxLevel(maxl); g->setSearchMethod((TreeDustGrid::SearchMethod)search);
                    g->setMaxMassFraction(mf); g->setSampleCount(samples); g->setMaxOpticalDepth(maxtau); g->setMaxDensDispFraction(maxdisp);
                    ds->setDustGrid(g); continue;
                }
                if (kind == "amesh")
                {
                    // grid amesh <adaptive mesh data file> densityUnits [densityIndex]; the dust is the mesh's own density field
                    std::string file; double units; int index = 0; in >> file >> units; in >> index;
                    auto* g = new AdaptiveMeshDustGrid(); setBox(g); g->setAdaptiveMeshFile(file); g->setDensityUnits(units); g->setDensityIndex(index);
                    ds->setDustGrid(g); continue;
                }
                if (kind == "voronoi")
                {
                    std::string file; in >> file;           // grid voronoi <particle file: x y z per line>
                    auto* g = new VoronoiDustGrid(); setBox(g); g->setParticleFile(file);
                    ds->setDustGrid(g); continue;
                }
                if (kind == "sphere1d" || kind == "sphere2d" || kind == "cylinder2d")
                {
                    // grid sphere1d <maxR> <n> <mesh> | grid sphere2d <maxR> <nr> <mesh> <ntheta> <mesh> | grid cylinder2d <maxR> <minZ> <maxZ> <nR> <mesh> <nz> <mesh>
                    if (kind == "sphere1d") { double r; int n; in >> r >> n; auto* g = new Sphere1DDustGrid(); g->setMaxR(r); g->setMeshR(makeMesh(in, n)); ds->setDustGrid(g); }
                    else if (kind == "sphere2d")
                    { double r; int nr, nt; in >> r >> nr; auto* g = new Sphere2DDustGrid(); g->setMaxR(r); g->setMeshR(makeMesh(in, nr)); in >> nt; g->setMeshTheta(makeMesh(in, nt)); ds->setDustGrid(g); }
                    else
                    { double r, z0, z1; int nR, nz; in >> r >> z0 >> z1 >> nR; auto* g = new Cylinder2DDustGrid(); g->setMaxR(r); g->setMinZ(z0); g->setMaxZ(z1);
                      g->setMeshR(makeMesh(in, nR)); in >> nz; g->setMeshZ(makeMesh(in, nz)); ds->setDustGrid(g); }
                    continue;
                }
                if (kind == "particletree")
                {
                    std::string tt, file; int extra = 0; in >> tt >> file; in >> extra;      // grid particletree oct|bin <particle file> [extraLevels]
                    auto* g = new ParticleTreeDustGrid(); setBox(g); g->setParticleFile(file); g->setExtraLevels(extra);
                    g->setTreeType(tt == "bin" ? ParticleTreeDustGrid::BinTree : ParticleTreeDustGrid::OctTree);
                    ds->setDustGrid(g); continue;
                }
                if (kind != "cartesian" && kind != "twophase") SKIRT_FATAL("unknown dust grid " + kind);
                int nx, ny, nz; in >> nx >> ny >> nz;
                CartesianDustGrid* g;
                if (kind == "twophase") { double ff, contrast; in >> ff >> contrast; auto* t = new TwoPhaseDustGrid(); t->setFillingFactor(ff); t->setContrast(contrast); g = t; }   // grid twophase nx ny nz ff contrast <meshes>
                else g = new CartesianDustGrid();
                g->setMinX(box[0]); g->setMaxX(box[1]); g->setMinY(box[2]); g->setMaxY(box[3]); g->setMinZ(box[4]); g->setMaxZ(box[5]);
                g->setMeshX(makeMesh(in, nx)); g->setMeshY(makeMesh(in, ny)); g->setMeshZ(makeMesh(in, nz));
                ds->setDustGrid(g);
            }
            else if (key == "dustmix") { in >> mixKind; if (mixKind == "interstellar") in >> mixFile; else in >> mixv[0] >> mixv[1] >> mixv[2]; }
            else if (key == "meshdust")
            {
                // meshdust: the dust component of an adaptive mesh's own density field (MeshDustComponent), with the current mix
                auto* c = new DustComp();
                if (mixKind == "interstellar") c->setMix(new InterstellarDustMix(mixFile));
                else { auto* m = new TableDustMix(); m->setTable({0.55e-6}, {mixv[0]}, {mixv[1]}, {mixv[2]}); c->setMix(m); }
                ds->addComponent(c);
            }
            else if (key == "dust")
            {
                double tau, lam; in >> tau >> lam;
                auto* c = new DustComp(); c->setGeometry(makeGeometry(in));
                if (mixKind == "interstellar") c->setMix(new InterstellarDustMix(mixFile));
                else { auto* m = new TableDustMix(); m->setTable({lam}, {mixv[0]}, {mixv[1]}, {mixv[2]}); c->setMix(m); }
                auto* n = new FaceOnDustCompNormalization(); n->setWavelength(lam); n->setOpticalDepth(tau); c->setNormalization(n);
                ds->addComponent(c);
            }
            else if (key == "stellar")
            {
                std::string lum; in >> lum;
                auto* c = new StellarComp();
                if (lum.rfind("bb:", 0) == 0) { double T, Lbol; if (std::sscanf(lum.c_str(), "bb:%lf:%lf", &T, &Lbol) != 2) SKIRT_FATAL("bad black body " + lum); c->setBlackBody(T, Lbol); }
                else { std::vector<double> L; std::istringstream ls(lum); std::string tok; while (std::getline(ls, tok, ',')) L.push_back(std::stod(tok)); c->setLuminosities(L); }
                c->setGeometry(makeGeometry(in));
                ss->addComponent(c);
            }
            else if (key == "perspective")
            {
                // perspective <name> Nx Ny width Vx Vy Vz Cx Cy Cz Ux Uy Uz focal
                std::string name; int nx, ny; double w, v[10]; in >> name >> nx >> ny >> w; for (double& q : v) in >> q;
                auto* pi = new PerspectiveInstrument(); pi->setInstrumentName(name); pi->setPixelsX(nx); pi->setPixelsY(ny); pi->setWidth(w);
                pi->setViewX(v[0]); pi->setViewY(v[1]); pi->setViewZ(v[2]); pi->setCrossX(v[3]); pi->setCrossY(v[4]); pi->setCrossZ(v[5]);
                pi->setUpX(v[6]); pi->setUpY(v[7]); pi->setUpZ(v[8]); pi->setFocal(v[9]);
                is->addInstrument(pi);
            }
            else if (key == "instrument")
            {
                std::string kind, name; double d, inc, az, pa; in >> kind >> name >> d >> inc >> az >> pa;
                if (kind == "multiframe")
                {
                    // instrument multiframe <name> d inc az pa <writeTotal> <writeStellarComps> <N>  then N x (nx fovx ny fovy xc yc)
                    int wt, wc, n; in >> wt >> wc >> n;
                    auto* mf = new MultiFrameInstrument(); mf->setWriteTotal(wt != 0); mf->setWriteStellarComps(wc != 0);
                    mf->setInstrumentName(name); mf->setDistance(d); mf->setInclination(inc); mf->setAzimuth(az); mf->setPositionAngle(pa);
                    for (int q = 0; q < n; q++)
                    {
                        InstrumentFrame fr; int nx, ny; double fx, fy, xc, yc; in >> nx >> fx >> ny >> fy >> xc >> yc;
                        fr.setPixelsX(nx); fr.setFieldOfViewX(fx); fr.setPixelsY(ny); fr.setFieldOfViewY(fy); fr.setCenterX(xc); fr.setCenterY(yc);
                        mf->addFrame(fr);
                    }
                    is->addInstrument(mf); continue;
                }
                Instrument* i = kind == "sed" ? (Instrument*)new SEDInstrument() : kind == "frame" ? (Instrument*)new FrameInstrument()
                              : kind == "full" ? (Instrument*)new FullInstrument() : (Instrument*)new SimpleInstrument();
                i->setInstrumentName(name); i->setDistance(d); i->setInclination(inc); i->setAzimuth(az); i->setPositionAngle(pa);
                if (kind != "sed") { int nx, ny; double fx, fy; in >> nx >> fx >> ny >> fy; i->setPixelsX(nx); i->setFieldOfViewX(fx); i->setPixelsY(ny); i->setFieldOfViewY(fy); }
                if (kind == "full") { int nscatt; in >> nscatt; static_cast<FullInstrument*>(i)->setScatteringLevels(nscatt); }
                is->addInstrument(i);
            }
            else SKIRT_FATAL("unknown key " + key);
        }
        sim.setStellarSystem(ss); sim.setDustSystem(ds); sim.setInstrumentSystem(is);
        std::string prefix = argv[2];
        if (writeOnly)
        {
            if (!sim.wavelengthGrid()) SKIRT_FATAL("no wavelength grid");
            sim.wavelengthGrid()->setup();
            const int Nl = sim.wavelengthGrid()->Nlambda();
            auto load = [&](const std::string& name, size_t n)
            {
                std::vector<double> v(n); std::ifstream f(prefix + "_" + name + ".f64", std::ios::binary);
                if (!f.read(reinterpret_cast<char*>(v.data()), sizeof(double) * n)) SKIRT_FATAL("cannot read " + prefix + "_" + name + ".f64");
                return v;
            };
            for (auto& i : sim.instrumentSystem()->instruments())
            {
                skg_instrument d = i->descriptor();
                if (FullInstrument* f = dynamic_cast<FullInstrument*>(i.get()))
                {
                    for (int c = 0; c < f->channels(); c++)
                    {
                        f->fchanv.push_back(load(i->name + "_frame" + std::to_string(c), (size_t)d.Nxp * d.Nyp * Nl));
                        f->Fchanv.push_back(load(i->name + "_sed" + std::to_string(c), Nl));
                    }
                    writeInstrument(*i, *sim.wavelengthGrid(), *units, prefix, "", true, sim.dustemission());
                    continue;
                }
                if (d.kind != SKG_INSTR_SED) i->ftotv = load(i->name + "_frame", (size_t)d.Nxp * d.Nyp * Nl);
                if (d.kind != SKG_INSTR_FRAME) i->Ftotv = load(i->name + "_sed", Nl);
                writeInstrument(*i, *sim.wavelengthGrid(), *units, prefix);
            }
            return 0;
        }
        sim.setup();
        skg_mc_stats st = sim.runstellaremission();
        int cycles = 0;
        if (sim.dustemission()) { if (pan) { cycles = sim.rundustselfabsorptionIfEnabled(); sim.rundustemission(); } else SKIRT_FATAL("dust emission needs a panchromatic simulation"); }
        sim.fetchResults();
        auto dump = [&](const std::string& name, const std::vector<double>& v)
        { std::ofstream out(prefix + "_" + name + ".f64", std::ios::binary); out.write(reinterpret_cast<const char*>(v.data()), sizeof(double) * v.size()); };
        for (auto& i : sim.instrumentSystem()->instruments())
        {
            if (!i->ftotv.empty()) dump(i->name + "_frame", i->ftotv);
            if (!i->Ftotv.empty()) dump(i->name + "_sed", i->Ftotv);
            if (FullInstrument* f = dynamic_cast<FullInstrument*>(i.get()))
                for (int c = 0; c < f->channels(); c++) { dump(i->name + "_frame" + std::to_string(c), f->fchanv[c]); dump(i->name + "_sed" + std::to_string(c), f->Fchanv[c]); }
            writeInstrument(*i, *sim.wavelengthGrid(), *units, prefix, "", true, sim.dustemission());        // Instrument::write(): calibration + FITS / SED files
        }
        if (!sim.Labs().empty()) dump("Labs", sim.Labs());
        dump("rho", sim.dustSystem()->rho());
        std::printf("{\"packets\": %llu, \"pathSegments\": %llu, \"scatterings\": %llu, \"kernel_ms\": %.3f, \"cells\": %d, \"wavelengths\": %d, \"selfabs_cycles\": %d}\n",
                    (unsigned long long)st.packets, (unsigned long long)st.pathSegments, (unsigned long long)st.scatterings, st.kernel_ms,
                    sim.dustSystem()->Ncells(), sim.wavelengthGrid()->Nlambda(), cycles);
        return 0;
    }
    catch (std::exception& ex) { std::fprintf(stderr, "*** Error: %s\n", ex.what()); return 1; }
}
