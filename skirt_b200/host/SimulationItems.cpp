// Host adapters -> C ABI (see SimulationItems.hpp).  Every engine status != 0 becomes a FatalError carrying
// skg_last_error(), the reference's FATALERROR convention.
#include "SimulationItems.hpp"

namespace skirt
{

static void check(int rc) { if (rc) SKIRT_FATAL(skg_last_error()); }

void CartesianDustGrid::upload(skg_engine* e) const
{ check(skg_grid_cartesian(e, _xv.data(), (int)_xv.size() - 1, _yv.data(), (int)_yv.size() - 1, _zv.data(), (int)_zv.size() - 1)); }

// DustSystem::setupSelfAfter: density table rho(m,h) (DustSystem.cpp:93-177) and the kappa tables per component.
// The reference averages 100 random samples per cell; a deterministic nsub^3 lattice is used here (set-up only).
void DustSystem::setup(const WavelengthGrid& lg)
{
    if (!_grid) SKIRT_FATAL("Dust grid was not set");
    if (_comps.empty()) SKIRT_FATAL("There are no dust components");
    _grid->setup();
    _Nlambda = lg.Nlambda();
    const int N = _grid->numCells(), C = (int)_comps.size();
    _kabs.resize((size_t)C * _Nlambda);
    _rho.assign((size_t)N * C, 0.0); _kext.resize((size_t)C * _Nlambda); _ksca.resize((size_t)C * _Nlambda); _g.resize((size_t)C * _Nlambda);
    for (int h = 0; h < C; h++)
    {
        DustComp& c = *_comps[h];
        if (!c.geometry || !c.mix || !c.norm) SKIRT_FATAL("dust component is incomplete");
        c.geometry->setup(); c.mix->setup(lg);
        for (int ell = 0; ell < _Nlambda; ell++)
        { _kabs[(size_t)h * _Nlambda + ell] = c.mix->kappaabsv[ell]; _kext[(size_t)h * _Nlambda + ell] = c.mix->kappaext(ell); _ksca[(size_t)h * _Nlambda + ell] = c.mix->kappascav[ell]; _g[(size_t)h * _Nlambda + ell] = c.mix->asymmparv[ell]; }
        // FaceOnDustCompNormalization.cpp:67-74: rho scale = tau / (SigmaZ * kappaext(lambda)); kappaext at lambda by
        // log-log interpolation on the simulation grid for panchromatic grids, the grid value itself for oligochromatic ones
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
        const double scale = c.norm->opticalDepth() / (c.geometry->SigmaZ() * kv);
        const int ns = _nsub;
        for (int m = 0; m < N; m++)
        {
            double b[6]; _grid->cellBox(m, b);
            double sum = 0;
            for (int a = 0; a < ns; a++) for (int bb = 0; bb < ns; bb++) for (int cc = 0; cc < ns; cc++)
                sum += c.geometry->density(b[0] + (a + 0.5) / ns * (b[3] - b[0]), b[1] + (bb + 0.5) / ns * (b[4] - b[1]), b[2] + (cc + 0.5) / ns * (b[5] - b[2]));
            _rho[(size_t)m * C + h] = scale * sum / (ns * ns * ns);
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
    for (auto& i : _instruments) d.push_back(i->descriptor());
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
    if (_ds) _ds->setup(*_lambdagrid);
    check(skg_engine_create(_device, &_engine));
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
    p.storeAbsorption = _ds && _ds->storeabsorptionrates();
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
        if (d.kind != SKG_INSTR_SED) { i->ftotv.assign((size_t)d.Nxp * d.Nyp * Nl, 0.0); check(skg_fetch_frame(_engine, idx, i->ftotv.data(), 0)); }
        if (d.kind != SKG_INSTR_FRAME) { i->Ftotv.assign(Nl, 0.0); check(skg_fetch_sed(_engine, idx, i->Ftotv.data(), 0)); }
        idx++;
    }
    if (_ds && _ds->storeabsorptionrates()) { _Labs.assign((size_t)_ds->Ncells() * Nl, 0.0); check(skg_fetch_labs(_engine, _Labs.data(), 0)); }
}

}   // namespace skirt
