// Host adapters -> C ABI (see SimulationItems.hpp).  Every engine status != 0 becomes a FatalError carrying
// skg_last_error(), the reference's FATALERROR convention.
#include "SimulationItems.hpp"

namespace skirt
{

static void check(int rc) { if (rc) SKIRT_FATAL(skg_last_error()); }

void CartesianDustGrid::upload(skg_engine* e) const
{ check(skg_grid_cartesian(e, _xv.data(), (int)_xv.size() - 1, _yv.data(), (int)_yv.size() - 1, _zv.data(), (int)_zv.size() - 1)); }

// ---- tree grids ------------------------------------------------------------------------------------------------------
void TreeDustGrid::build(skg_engine* e, const std::vector<skg_source>& geoms, const std::vector<double>& norms, uint64_t seed)
{
    BoxDustGrid::setup();
    if (_Nrandom < 1) SKIRT_FATAL("Number of random samples must be at least 1");
    if (_maxOpticalDepth < 0.0) SKIRT_FATAL("The maximum mean optical depth should be positive");
    if (_maxMassFraction < 0.0) SKIRT_FATAL("The maximum mass fraction should be positive");
    if (_maxDensDispFraction < 0.0) SKIRT_FATAL("The maximum density dispersion fraction should be positive");      // TreeDustGrid.cpp:63
    try
    {
        skirt::TreeBuilder tb(_kind, _ext, _minlevel, _maxlevel);
        double total = 0; for (double v : norms) total += v;                        // CompDustDistribution::mass
        const bool always = _maxOpticalDepth == 0 && _maxMassFraction == 0 && _maxDensDispFraction == 0;      // TreeDustGrid.cpp:192
        const double kappaV = 2600.0;                                               // Units::kappaV(), Units.cpp:30
        std::vector<double> box, mass, disp; std::vector<unsigned char> flags;
        while (!tb.done())
        {
            const size_t n = tb.frontierSize();
            if (!tb.frontierNeedsDecision()) { tb.subdivide(nullptr); continue; }
            flags.assign(n, always ? 1 : 0);
            if (!always)
            {
                box.resize(6 * n); mass.resize(n); tb.frontierBoxes(box.data());
                const uint64_t levelSeed = seed + 7919ull * (uint64_t)tb.frontierLevel();
                if (_maxDensDispFraction > 0)
                {
                    disp.resize(n);
                    check(skg_sample_boxes_dispersion(e, (int64_t)n, box.data(), (int)geoms.size(), geoms.data(), norms.data(), _Nrandom,
                                                      levelSeed, mass.data(), disp.data()));
                }
                else check(skg_sample_boxes(e, (int64_t)n, box.data(), (int)geoms.size(), geoms.data(), norms.data(), _Nrandom, levelSeed, mass.data()));
                for (size_t q = 0; q < n; q++)
                {
                    const double* b = &box[6 * q]; const double vol = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2]);
                    if (_maxMassFraction > 0 && mass[q] / total >= _maxMassFraction) flags[q] = 1;                      // :197-201
                    if (_maxOpticalDepth > 0 && kappaV * mass[q] / std::pow(vol, 2. / 3.) >= _maxOpticalDepth) flags[q] = 1;   // :204-208
                    if (_maxDensDispFraction > 0 && disp[q] >= _maxDensDispFraction) flags[q] = 1;                      // :215-221
                }
            }
            tb.subdivide(flags.data());
        }
        tb.finish((int)_search);
        _t = tb.tables();
    }
    catch (std::runtime_error& ex) { SKIRT_FATAL(ex.what()); }
    _cellNode.assign(_t.Ncells, 0);
    for (int l = 0; l < _t.Nnodes; l++) if (_t.cell[l] >= 0) _cellNode[_t.cell[l]] = l;
}

void TreeDustGrid::upload(skg_engine* e) const
{
    if (_t.Nnodes == 0) SKIRT_FATAL("the tree has not been built");
    check(skg_grid_tree(e, _kind, (int)_search, _t.Nnodes, _t.box.data(), _t.child0.data(), _t.parent.data(), _t.cell.data(), _t.dir.data(),
                        _t.nbrStart.empty() ? nullptr : _t.nbrStart.data(), _t.nbrIds.empty() ? nullptr : _t.nbrIds.data()));
}

// ---- grids with symmetries ---------------------------------------------------------------------------------------------------
void Sphere1DDustGrid::setup()
{
    if (_rmax <= 0) SKIRT_FATAL("The outer radius of the grid should be positive");           // SphereDustGrid.cpp:21-26
    if (!_meshr) SKIRT_FATAL("the radial mesh was not set");
    _N1 = _meshr->numBins(); _N2 = 0;
    _v1 = _meshr->mesh(); for (double& v : _v1) v *= _rmax;
    _volumes.resize(_N1);
    for (int i = 0; i < _N1; i++) { const double rL = _v1[i], rR = _v1[i + 1]; _volumes[i] = 4.0 * M_PI / 3.0 * (rR - rL) * (rR * rR + rR * rL + rL * rL); }
}
void Sphere1DDustGrid::upload(skg_engine* e) const { check(skg_grid_sphere1d(e, _N1, _v1.data())); }

void Sphere2DDustGrid::setup()
{
    if (_rmax <= 0) SKIRT_FATAL("The outer radius of the grid should be positive");
    if (!_meshr || !_mesht) SKIRT_FATAL("the radial or the polar mesh was not set");
    _N1 = _meshr->numBins();
    _v1 = _meshr->mesh(); for (double& v : _v1) v *= _rmax;
    _v2 = _mesht->mesh(); for (double& v : _v2) v *= M_PI;
    _cv.resize(_v2.size()); for (size_t k = 0; k < _v2.size(); k++) _cv[k] = std::cos(_v2[k]);
    _cv.front() = 1.; _cv.back() = -1.;
    // Sphere2DDustGrid.cpp:39-72: path() needs a border in the xy-plane -- snap one that is there, insert one otherwise
    int zeros = 0, at = -1;
    for (size_t k = 1; k + 1 < _cv.size(); k++) if (std::fabs(_cv[k]) < 1e-9) { zeros++; at = (int)k; }
    if (zeros > 1) SKIRT_FATAL("There are multiple grid points very close to pi/2");
    if (zeros == 1) _cv[at] = 0.;
    else
    {
        size_t pos = 0; while (pos < _cv.size() && _cv[pos] > 0) pos++;
        _v2.insert(_v2.begin() + pos, M_PI_2); _cv.insert(_cv.begin() + pos, 0.);
    }
    _N2 = (int)_v2.size() - 1;
    _volumes.resize((size_t)_N1 * _N2);
    for (int i = 0; i < _N1; i++) for (int k = 0; k < _N2; k++)
        _volumes[k + (size_t)_N2 * i] = (2.0 / 3.0) * M_PI * (std::pow(_v1[i + 1], 3) - std::pow(_v1[i], 3)) * (std::cos(_v2[k]) - std::cos(_v2[k + 1]));
}
void Sphere2DDustGrid::upload(skg_engine* e) const { check(skg_grid_sphere2d(e, _N1, _v1.data(), _N2, _v2.data(), _cv.data())); }

void Cylinder2DDustGrid::setup()
{
    if (_Rmax <= 0) SKIRT_FATAL("The outer radius of the grid should be positive");
    if (_zmax <= _zmin) SKIRT_FATAL("The extent of the cylinder should be positive in the Z direction");
    if (!_meshR || !_meshz) SKIRT_FATAL("the radial or the vertical mesh was not set");
    _N1 = _meshR->numBins(); _N2 = _meshz->numBins();
    _v1 = _meshR->mesh(); for (double& v : _v1) v *= _Rmax;
    _v2 = _meshz->mesh(); for (double& v : _v2) v = v * (_zmax - _zmin) + _zmin;
    _volumes.resize((size_t)_N1 * _N2);
    for (int i = 0; i < _N1; i++) for (int k = 0; k < _N2; k++)
        _volumes[k + (size_t)_N2 * i] = M_PI * (_v2[k + 1] - _v2[k]) * (_v1[i + 1] - _v1[i]) * (_v1[i + 1] + _v1[i]);
}
void Cylinder2DDustGrid::upload(skg_engine* e) const { check(skg_grid_cylinder2d(e, _N1, _v1.data(), _N2, _v2.data())); }

// ---- particle tree ---------------------------------------------------------------------------------------------------------
static void readParticleFile(const std::string& file, std::vector<double>& xyz, const char* what)
{
    std::ifstream in(file);
    if (!in) SKIRT_FATAL(std::string("Could not open the ") + what + " particle file " + file);
    xyz.clear();
    std::string line; double x, y, z;
    while (std::getline(in, line)) { if (line.empty() || line[0] == '#') continue; std::istringstream is(line); if (is >> x >> y >> z) { xyz.push_back(x); xyz.push_back(y); xyz.push_back(z); } }
}

void ParticleTreeDustGrid::setup()
{
    BoxDustGrid::setup();
    if (_extra < 0) SKIRT_FATAL("The number of extra levels should not be negative");
    if (!_file.empty()) readParticleFile(_file, _particles, "tree");
    try
    {
        skirt::TreeBuilder tb((int)_kind, _ext, 0, 2);          // (the level limits of TreeDustGrid play no role here)
        tb.addParticles(_particles.data(), _particles.size() / 3, _extra);
        _t = tb.tables();
    }
    catch (std::runtime_error& ex) { SKIRT_FATAL(ex.what()); }
    _cellNode.assign(_t.Ncells, 0);
    for (int l = 0; l < _t.Nnodes; l++) if (_t.cell[l] >= 0) _cellNode[_t.cell[l]] = l;
}

void ParticleTreeDustGrid::upload(skg_engine* e) const
{
    if (_t.Nnodes == 0) SKIRT_FATAL("the tree has not been built");
    check(skg_grid_tree(e, (int)_kind, 3, _t.Nnodes, _t.box.data(), _t.child0.data(), _t.parent.data(), _t.cell.data(), _t.dir.data(), nullptr, nullptr));
}

// ---- adaptive mesh -------------------------------------------------------------------------------------------------------
void AdaptiveMeshDustGrid::setup()
{
    BoxDustGrid::setup();
    if (!_file.empty())
    {
        // AdaptiveMeshAsciiFile::read, AdaptiveMeshAsciiFile.cpp:43-100
        std::ifstream in(_file);
        if (!in) SKIRT_FATAL("Could not open the adaptive mesh data file " + _file);
        _nxyz.clear(); _values.clear();
        std::string line;
        while (std::getline(in, line))
        {
            std::istringstream is(line); std::string first;
            if (!(is >> first) || first[0] == '#') continue;
            if (first[0] == '!')
            {
                int n[3] = {0, 0, 0}; int have = 0;
                if (first.size() > 1) { n[have++] = std::atoi(first.c_str() + 1); }
                while (have < 3 && (is >> n[have])) have++;
                if (n[0] < 1 || n[1] < 1 || n[2] < 1) SKIRT_FATAL("Invalid nonleaf line in mesh data");
                _nxyz.insert(_nxyz.end(), n, n + 3); _values.push_back(0.0);
            }
            else
            {
                std::vector<std::string> cols{first}; std::string w; while (is >> w) cols.push_back(w);
                if (_densityIndex < 0) SKIRT_FATAL("Field index out of range");
                if (_densityIndex >= (int)cols.size()) SKIRT_FATAL("Insufficient number of field values in mesh data");
                char* end = nullptr; const double v = std::strtod(cols[_densityIndex].c_str(), &end);
                if (!end || *end) SKIRT_FATAL("Invalid leaf line in mesh data");
                _nxyz.insert(_nxyz.end(), {0, 0, 0}); _values.push_back(v);
            }
        }
    }
    if (_nxyz.empty() || _nxyz.size() != 3 * _values.size()) SKIRT_FATAL("Reached end of file in mesh data before all nodes were read");
    try { _t = skirt::buildAdaptiveMesh(_ext, _nxyz.data(), _values.size()); }
    catch (std::runtime_error& ex) { SKIRT_FATAL(ex.what()); }
    _cellNode.assign(_t.Ncells, 0);
    for (int l = 0; l < _t.Nnodes; l++) if (_t.cell[l] >= 0) _cellNode[_t.cell[l]] = l;
}

void AdaptiveMeshDustGrid::upload(skg_engine* e) const
{ check(skg_grid_amesh(e, _t.Nnodes, _t.box.data(), _t.nxyz.data(), _t.child0.data(), _t.cell.data(), _t.wallNbr.data())); }

// ---- Voronoi ---------------------------------------------------------------------------------------------------------------
void VoronoiDustGrid::setup()
{
    BoxDustGrid::setup();
    if (!_file.empty())
    {
        std::ifstream in(_file);
        if (!in) SKIRT_FATAL("Could not open the Voronoi particle file " + _file);
        _particles.clear();
        std::string line; double x, y, z;
        while (std::getline(in, line)) { if (line.empty() || line[0] == '#') continue; std::istringstream is(line); if (is >> x >> y >> z) { _particles.push_back(x); _particles.push_back(y); _particles.push_back(z); } }
    }
    // VoronoiMesh.cpp:262-263: particles outside of the domain are ignored
    std::vector<double> inside;
    for (size_t q = 0; q + 2 < _particles.size(); q += 3)
    {
        const double* p = &_particles[q];
        if (p[0] >= _ext[0] && p[0] <= _ext[1] && p[1] >= _ext[2] && p[1] <= _ext[3] && p[2] >= _ext[4] && p[2] <= _ext[5]) inside.insert(inside.end(), p, p + 3);
    }
    if (inside.empty()) SKIRT_FATAL("a Voronoi grid needs at least one particle inside its extent");
    try { _t = skirt::buildVoronoiMesh(_ext, inside.data(), inside.size() / 3); }
    catch (std::runtime_error& ex) { SKIRT_FATAL(ex.what()); }
}

void VoronoiDustGrid::upload(skg_engine* e) const
{
    check(skg_grid_voronoi(e, _t.Ncells, _t.particles.data(), _t.nbrStart.data(), _t.nbrIds.data(), _ext, _t.nb, _t.blkStart.data(), _t.blkIds.data(),
                           _t.blkTree.data(), (int)_t.kdM.size(), _t.kdM.data(), _t.kdAxis.data(), _t.kdUp.data(), _t.kdLeft.data(), _t.kdRight.data(),
                           _t.cellBox.data()));
}

// DustSystem::setupSelfAfter: density table rho(m,h) (DustSystem.cpp:93-177) and the kappa tables per component.
// The reference averages 100 random samples per cell; a deterministic nsub^3 lattice is used here (set-up only).
// the part of the set-up that needs no device: property validation, meshes, the tessellation / mesh of file-based grids
void DustSystem::presetup(const WavelengthGrid& lg)
{
    if (_presetup) return;
    if (!_grid) SKIRT_FATAL("Dust grid was not set");
    if (_comps.empty()) SKIRT_FATAL("There are no dust components");
    _grid->setup();
    for (auto& c : _comps) { if (c->geometry) c->geometry->setup(); if (c->mix) c->mix->setup(lg); }
    _presetup = true;
}

void DustSystem::setup(const WavelengthGrid& lg, skg_engine* e, uint64_t seed)
{
    presetup(lg);
    _Nlambda = lg.Nlambda();
    const int C = (int)_comps.size();
    _kabs.resize((size_t)C * _Nlambda);
    _kext.resize((size_t)C * _Nlambda); _ksca.resize((size_t)C * _Nlambda); _g.resize((size_t)C * _Nlambda);
    std::vector<double> own;
    const bool meshDust = _grid->ownDensity(own);       // AdaptiveMeshDustDistribution: the mesh's own density field, one component
    if (meshDust)
    {
        if (C != 1 || !_comps[0]->mix) SKIRT_FATAL("an adaptive mesh dust distribution has exactly one dust component with a mix");
        for (int ell = 0; ell < _Nlambda; ell++)
        { _kabs[ell] = _comps[0]->mix->kappaabsv[ell]; _kext[ell] = _comps[0]->mix->kappaext(ell); _ksca[ell] = _comps[0]->mix->kappascav[ell]; _g[ell] = _comps[0]->mix->asymmparv[ell]; }
        _rho = own;
        return;
    }
    // mass normalisation of every component (FaceOnDustCompNormalization.cpp:67-74) and its optical properties
    std::vector<double> norms(C); std::vector<skg_source> geoms;
    for (int h = 0; h < C; h++)
    {
        DustComp& c = *_comps[h];
        if (!c.geometry || !c.mix || !c.norm) SKIRT_FATAL("dust component is incomplete");
        for (int ell = 0; ell < _Nlambda; ell++)
        { _kabs[(size_t)h * _Nlambda + ell] = c.mix->kappaabsv[ell]; _kext[(size_t)h * _Nlambda + ell] = c.mix->kappaext(ell); _ksca[(size_t)h * _Nlambda + ell] = c.mix->kappascav[ell]; _g[(size_t)h * _Nlambda + ell] = c.mix->asymmparv[ell]; }
        double kv;
        if (_Nlambda == 1) kv = c.mix->kappaext(0);
        else
        {
            const std::vector<double>& lv = lg.lambdav(); double lam = c.norm->wavelength();
            size_t i = std::upper_bound(lv.begin(), lv.end(), lam) - lv.begin();
            i = std::max<size_t>(1, std::min(lv.size() - 1, i)) - 1;
            double t = (std::log10(lam) - std::log10(lv[i])) / (std::log10(lv[i + 1]) - std::log10(lv[i]));
            kv = std::pow(10.0, std::log10(c.mix->kappaext((int)i)) + t * (std::log10(c.mix->kappaext((int)i + 1)) - std::log10(c.mix->kappaext((int)i))));
        }
        norms[h] = c.norm->opticalDepth() / (c.geometry->SigmaZ() * kv);
        geoms.push_back(c.geometry->sampler());
    }
    if (_grid->densityOnDevice())
    {
        // tree / Voronoi grids: the grid is grown (tree) and uploaded, then DustSystem::setSampleDensityBody (DustSystem.cpp:152-177)
        // runs on the device: the mean of _Nrandom random positions per cell
        _grid->build(e, geoms, norms, seed);
        _grid->upload(e);
        _rho.assign((size_t)_grid->numCells() * C, 0.0);
        check(skg_sample_density(e, C, geoms.data(), norms.data(), _Nrandom, seed, _rho.data()));
        return;
    }
    const int N = _grid->numCells();
    _rho.assign((size_t)N * C, 0.0);
    for (int h = 0; h < C; h++)
    {
        DustComp& c = *_comps[h];
        const double scale = norms[h];
        const int ns = _nsub;
        for (int m = 0; m < N; m++)
        {
            double b[6]; _grid->cellBox(m, b);
            double sum = 0;
            for (int a = 0; a < ns; a++) for (int bb = 0; bb < ns; bb++) for (int cc = 0; cc < ns; cc++)
                sum += c.geometry->density(b[0] + (a + 0.5) / ns * (b[3] - b[0]), b[1] + (bb + 0.5) / ns * (b[4] - b[1]), b[2] + (cc + 0.5) / ns * (b[5] - b[2]));
            _rho[(size_t)m * C + h] = _grid->weight(m) * scale * sum / (ns * ns * ns);       // DustSystem.cpp:165-176
        }
    }
}

void DustSystem::upload(skg_engine* e) const
{
    _grid->upload(e);
    check(skg_medium(e, _grid->numCells(), (int)_comps.size(), _Nlambda, _rho.data(), _kext.data(), _ksca.data(), _g.data()));
}

void StellarSystem::upload(skg_engine* e) const
{
    std::vector<skg_source> src; std::vector<double> L;
    for (auto& c : _comps) { src.push_back(c->geometry->sampler()); L.insert(L.end(), c->Lv.begin(), c->Lv.end()); }
    check(skg_sources(e, (int)src.size(), src.data(), _Nlambda, L.data(), _emissionBias));
}

void InstrumentSystem::upload(skg_engine* e) const
{
    std::vector<skg_instrument> d;
    for (auto& i : _instruments)
    {
        const MultiFrameInstrument* mf = dynamic_cast<const MultiFrameInstrument*>(i.get());
        d.push_back(mf ? mf->descriptorWithFrames() : i->descriptor());
    }
    check(skg_instruments(e, (int)d.size(), d.data()));
}

void MonteCarloSimulation::setup()
{
    // MonteCarloSimulation::setupSelfBefore, MonteCarloSimulation.cpp:55-67
    if (!_lambdagrid) SKIRT_FATAL("Wavelength grid was not set");
    if (!_ss) SKIRT_FATAL("Stellar system was not set");
    if (!_is) SKIRT_FATAL("Instrument system was not set");
    _lambdagrid->setup();
    _ss->setup(*_lambdagrid);
    if (_ds) _ds->presetup(*_lambdagrid);            // validation and host-only set-up before a device is asked for
    check(skg_engine_create(_device, &_engine));
    if (_ds) _ds->setup(*_lambdagrid, _engine, (uint64_t)_seed);
    if (_ds) _ds->upload(_engine);
    _ss->upload(_engine);
    if (!_ds)
    {
        // instruments need the number of wavelengths, which the engine learns from the medium: an empty medium stands in
        SKIRT_FATAL("a simulation without a dust system is not supported by the engine front end");
    }
    _is->upload(_engine);
    if (_dustemission)
    {
        if (!_lambdagrid->issampledrange()) SKIRT_FATAL("dust emission needs a panchromatic wavelength grid");
        _ds->setStoreAbsorptionRates(true);          // PanDustSystem::storeabsorptionrates() == dustemission()
        std::vector<double> vol = _ds->volumes(), lam = _lambdagrid->lambdav(), dlam(lam.size());
        for (size_t i = 0; i < lam.size(); i++) dlam[i] = _lambdagrid->dlambda((int)i);
        check(skg_dust_library(_engine, vol.data(), _ds->kappaabs().data(), lam.data(), dlam.data()));
    }
    if (_nranks > 1)
    {
        if (!_uid) SKIRT_FATAL("the NCCL unique id was not set for a multi-process run");
        check(skg_comm_init(_engine, _rank, _nranks, _uid));
    }
}

skg_mc_stats MonteCarloSimulation::runstellaremission()
{
    if (!_engine) SKIRT_FATAL("Simulation has not been setup before being run");
    skg_mc_params p{};
    // every process shoots ceil(packages/nranks) packets per wavelength with its own block of Philox streams
    double npr = std::ceil(_packages / _nranks);
    p.packages = npr; p.luminosityScale = npr * _nranks;
    p.minWeightReduction = _minWeightReduction; p.minScattEvents = _minfs; p.scattBias = _xi;
    p.storeAbsorption = _ds && _ds->storeabsorptionrates(); p.continuousScattering = _continuousScattering;
    p.seed = (uint64_t)_seed; p.streamOffset = (uint64_t)(_rank * npr);
    p.ellBegin = 0; p.ellEnd = _lambdagrid->Nlambda();
    skg_mc_stats st{};
    check(skg_run_stellar(_engine, &p, &st));
    // the stellar absorption table is summed over the processes once (PanDustSystem::sumResults(true), PanDustSystem.cpp:394-404);
    // the detector arrays once, when they are read (fetchResults; Instrument::sumResults at write(), Instrument.cpp:57-65)
    if (_nranks > 1 && p.storeAbsorption) check(skg_allreduce(_engine, SKG_REDUCE_LABS_STELLAR, nullptr));
    return st;
}

skg_mc_stats MonteCarloSimulation::shootDust(int phase, double packages)
{
    skg_mc_params p{};
    double npr = std::ceil(packages / _nranks);
    p.packages = npr; p.luminosityScale = npr * _nranks;
    p.minWeightReduction = _minWeightReduction; p.minScattEvents = _minfs; p.scattBias = _xi;
    p.seed = (uint64_t)_seed + 7919ull * (uint64_t)(++_phaseCounter); p.streamOffset = (uint64_t)(_rank * npr);
    p.ellBegin = 0; p.ellEnd = _lambdagrid->Nlambda();
    double* dL = nullptr;
    check(skg_dust_cell_luminosities(_engine, &dL));            // PanDustSystem::calculatedustemission + the vectors Lv
    if (phase == SKG_PHASE_DUST_SELFABS) check(skg_reset_labs_dust(_engine));   // rebootLabsdust, after the spectra were made
    skg_mc_stats st{};
    check(skg_run_dust(_engine, &p, phase, _dustBias, SKG_DEVICE, dL, &st));
    // PanDustSystem::sumResults(false): the dust absorption of this cycle, summed before the next spectra are made from it
    if (_nranks > 1 && phase == SKG_PHASE_DUST_SELFABS) check(skg_allreduce(_engine, SKG_REDUCE_LABS_DUST, nullptr));
    return st;
}

// PanMonteCarloSimulation::rundustselfabsorption, PanMonteCarloSimulation.cpp:105-185
int MonteCarloSimulation::rundustselfabsorption()
{
    if (!_engine || !_dustemission) SKIRT_FATAL("dust self-absorption needs a set-up simulation with dust emission");
    const double stage_factor[] = {1. / 10., 1. / 3., 1.}; const double stage_epsmax[] = {0.010, 0.007, 0.005};
    double prev = 0.; int total = 0;
    for (int stage = 0; stage < 3; stage++)
    {
        bool fixed = _cycles > 0; const int Ncyclesmax = fixed ? _cycles : 100;
        bool convergence = false; int cycle = 1;
        while (cycle <= Ncyclesmax && (!convergence || fixed))
        {
            shootDust(SKG_PHASE_DUST_SELFABS, _packages * stage_factor[stage]);
            double Labsdusttot = 0;                 // PanDustSystem::Labsdusttot(): the same number on every process
            check(skg_labs_dust_total(_engine, &Labsdusttot));
            double eps = std::fabs((Labsdusttot - prev) / Labsdusttot);
            prev = Labsdusttot;
            if ((stage < 2 || cycle > 1) && eps < stage_epsmax[stage]) convergence = true;
            cycle++; total++;
        }
    }
    return total;
}

// PanMonteCarloSimulation::rundustemission, PanMonteCarloSimulation.cpp:242-264
skg_mc_stats MonteCarloSimulation::rundustemission()
{
    if (!_engine || !_dustemission) SKIRT_FATAL("dust emission needs a set-up simulation with dust emission");
    return shootDust(SKG_PHASE_DUST_EMISSION, _packages * _dustBoost);
}

void MonteCarloSimulation::fetchResults()
{
    const int Nl = _lambdagrid->Nlambda();
    if (_nranks > 1) check(skg_allreduce(_engine, SKG_REDUCE_INSTRUMENTS, nullptr));     // Instrument::sumResults; once
    int idx = 0;
    for (auto& i : _is->instruments())
    {
        skg_instrument d = i->descriptor();
        if (FullInstrument* f = dynamic_cast<FullInstrument*>(i.get()))
        {
            f->fchanv.assign(f->channels(), std::vector<double>((size_t)d.Nxp * d.Nyp * Nl, 0.0));
            f->Fchanv.assign(f->channels(), std::vector<double>(Nl, 0.0));
            for (int c = 0; c < f->channels(); c++)
            {
                check(skg_fetch_frame_channel(_engine, idx, c, f->fchanv[c].data(), 0));
                check(skg_fetch_sed_channel(_engine, idx, c, f->Fchanv[c].data(), 0));
            }
            idx++; continue;
        }
        if (MultiFrameInstrument* mf = dynamic_cast<MultiFrameInstrument*>(i.get()))
        {
            // the arrays InstrumentFrame::calibrateAndWriteData lists (InstrumentFrame.cpp:193-207): total, then stellar_k
            std::vector<int> which; mf->arrayNames.clear();
            if (d.writeTotal) { which.push_back(-1); mf->arrayNames.push_back("total"); }
            if (d.writeStellarComps) for (int k = 0; k < _ss->Ncomp(); k++) { which.push_back(k); mf->arrayNames.push_back("stellar_" + std::to_string(k)); }
            mf->arrays.assign(Nl, std::vector<std::vector<double>>());
            for (int ell = 0; ell < Nl; ell++)
                for (int w : which)
                {
                    const skg_instrument_frame& f = mf->frames()[ell];
                    mf->arrays[ell].emplace_back((size_t)f.Nxp * f.Nyp, 0.0);
                    check(skg_fetch_multiframe(_engine, idx, w, ell, mf->arrays[ell].back().data(), 0));
                }
            idx++; continue;
        }
        if (d.kind != SKG_INSTR_SED) { i->ftotv.assign((size_t)d.Nxp * d.Nyp * Nl, 0.0); check(skg_fetch_frame(_engine, idx, i->ftotv.data(), 0)); }
        if (d.kind != SKG_INSTR_FRAME && d.kind != SKG_INSTR_PERSPECTIVE) { i->Ftotv.assign(Nl, 0.0); check(skg_fetch_sed(_engine, idx, i->Ftotv.data(), 0)); }
        idx++;
    }
    if (_ds && _ds->storeabsorptionrates()) { _Labs.assign((size_t)_ds->Ncells() * Nl, 0.0); check(skg_fetch_labs(_engine, _Labs.data(), 0)); }
}

}   // namespace skirt
