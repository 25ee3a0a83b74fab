// C++ host side of the engine: simulation items with the reference's class and property names
// (MonteCarloSimulation, DustSystem, DustGrid, StellarSystem, Instrument, ... -- the names a ski file uses),
// each of which only keeps what the propagation hot path needs, flattens it into the POD tables of
// include/skirtgpu.h and hands it to libskirtgpu.so.  Nothing here walks a grid or shoots a packet: that is all
// on the GPU behind the C ABI; an engine error becomes a FatalError like the reference's FATALERROR
// (FatalError.hpp:47).  Set-up arithmetic follows the reference files cited at each class.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/skirtgpu.h"
#include "GridBuilders.hpp"

namespace skirt
{

struct FatalError : std::runtime_error { using std::runtime_error::runtime_error; };
#define SKIRT_FATAL(msg) throw ::skirt::FatalError(msg)

namespace Units      // Units.cpp:17-30
{
    constexpr double pc = 3.08567758e16, Lsun = 3.839e26, Msun = 1.9891e30, lambdaV = 550e-9, kappaV = 2600.0;
    constexpr double h = 6.62606957e-34, c = 2.99792458e8, k = 1.3806488e-23;
}

// ---- wavelength grids ----------------------------------------------------------------------------------
class WavelengthGrid
{
public:
    virtual ~WavelengthGrid() {}
    virtual void setup() = 0;
    int Nlambda() const { return (int)_lambdav.size(); }
    double lambda(int ell) const { return _lambdav[ell]; }
    double dlambda(int ell) const { return _dlambdav[ell]; }
    const std::vector<double>& lambdav() const { return _lambdav; }
    virtual bool issampledrange() const = 0;
protected:
    std::vector<double> _lambdav, _dlambdav;
};

class OligoWavelengthGrid : public WavelengthGrid
{
public:
    void setWavelengths(const std::vector<double>& v) { _lambdav = v; }
    void setup() override
    {
        if (_lambdav.empty()) SKIRT_FATAL("There must be at least one wavelength in the grid");
        std::sort(_lambdav.begin(), _lambdav.end());
        _dlambdav.resize(_lambdav.size());
        for (size_t i = 0; i < _lambdav.size(); i++) _dlambdav[i] = 0.001 * _lambdav[i];     // OligoWavelengthGrid.cpp:25-26
    }
    bool issampledrange() const override { return false; }
};

class LogWavelengthGrid : public WavelengthGrid      // LogWavelengthGrid.cpp:18-28, PanWavelengthGrid.cpp:25-37
{
public:
    void setMinWavelength(double v) { _min = v; }
    void setMaxWavelength(double v) { _max = v; }
    void setPoints(int v) { _points = v; }
    void setup() override
    {
        if (_min <= 0) SKIRT_FATAL("the shortest wavelength should be positive");
        if (_max <= _min) SKIRT_FATAL("the longest wavelength should be larger than the shortest");
        if (_points < 3) SKIRT_FATAL("There must be at least three bins in a panchromatic wavelength grid");
        int n = _points - 1;
        double logxmin = std::log10(_min), dlogx = std::log10(_max / _min) / n;      // NR::loggrid, NR.hpp:269-275
        _lambdav.resize(n + 1); _dlambdav.resize(n + 1);
        for (int i = 0; i <= n; i++) _lambdav[i] = std::pow(10.0, logxmin + i * dlogx);
        for (int i = 0; i <= n; i++)
        {
            double lo = i == 0 ? _lambdav[0] : std::sqrt(_lambdav[i - 1] * _lambdav[i]);
            double hi = i == n ? _lambdav[n] : std::sqrt(_lambdav[i] * _lambdav[i + 1]);
            _dlambdav[i] = hi - lo;
        }
    }
    bool issampledrange() const override { return true; }
private:
    double _min = 0, _max = 0; int _points = 0;
};

// NestedLogWavelengthGrid.cpp:21-60: a low-resolution logarithmic grid whose points inside the zoom range are replaced by a
// high-resolution logarithmic subgrid; bin widths as for every PanWavelengthGrid (PanWavelengthGrid.cpp:25-37)
class NestedLogWavelengthGrid : public WavelengthGrid
{
public:
    void setMinWavelength(double v) { _min = v; }
    void setMaxWavelength(double v) { _max = v; }
    void setPoints(int v) { _points = v; }
    void setMinWavelengthSubGrid(double v) { _zoomMin = v; }
    void setMaxWavelengthSubGrid(double v) { _zoomMax = v; }
    void setPointsSubGrid(int v) { _zoomPoints = v; }
    void setup() override
    {
        if (_points < 2) SKIRT_FATAL("the number of points in the low-resolution grid should be at least 2");
        if (_zoomPoints < 2) SKIRT_FATAL("the number of points in the high-resolution subgrid should be at least 2");
        if (_min <= 0) SKIRT_FATAL("the shortest wavelength should be positive");
        if (_zoomMin <= _min || _zoomMax <= _zoomMin || _max <= _zoomMax)
            SKIRT_FATAL("the high-resolution subgrid should be properly nested in the low-resolution grid");
        auto loggrid = [](double xmin, double xmax, int n)                                  // NR::loggrid, NR.hpp:269-275
        {
            std::vector<double> xv(n + 1);
            const double logxmin = std::log10(xmin), dlogx = std::log10(xmax / xmin) / n;
            for (int i = 0; i <= n; i++) xv[i] = std::pow(10.0, logxmin + i * dlogx);
            return xv;
        };
        const std::vector<double> low = loggrid(_min, _max, _points - 1), zoom = loggrid(_zoomMin, _zoomMax, _zoomPoints - 1);
        _lambdav.clear();
        for (double v : low) if (v < _zoomMin) _lambdav.push_back(v);
        _lambdav.insert(_lambdav.end(), zoom.begin(), zoom.end());
        for (double v : low) if (v > _zoomMax) _lambdav.push_back(v);
        const int n = (int)_lambdav.size() - 1;
        if (n + 1 < 3) SKIRT_FATAL("There must be at least three bins in a panchromatic wavelength grid");
        _dlambdav.resize(n + 1);
        for (int i = 0; i <= n; i++)
        {
            const double lo = i == 0 ? _lambdav[0] : std::sqrt(_lambdav[i - 1] * _lambdav[i]);
            const double hi = i == n ? _lambdav[n] : std::sqrt(_lambdav[i] * _lambdav[i + 1]);
            _dlambdav[i] = hi - lo;
        }
    }
    bool issampledrange() const override { return true; }
private:
    double _min = 0, _max = 0, _zoomMin = 0, _zoomMax = 0; int _points = 0, _zoomPoints = 0;
};

// FileWavelengthGrid.cpp:22-47: the number of wavelengths, then the wavelengths in micron (divided by 1e6, sorted)
class FileWavelengthGrid : public WavelengthGrid
{
public:
    void setFilename(const std::string& v) { _filename = v; }
    void setup() override
    {
        std::ifstream file(_filename);
        if (!file.is_open()) SKIRT_FATAL("Could not open the data file " + _filename);
        int n = 0; file >> n;
        if (n < 3) SKIRT_FATAL("There must be at least three bins in a panchromatic wavelength grid");       // PanWavelengthGrid.cpp:30
        _lambdav.resize(n);
        for (int k = 0; k < n; k++) { file >> _lambdav[k]; _lambdav[k] /= 1e6; }
        if (!file) SKIRT_FATAL("the data file " + _filename + " holds fewer wavelengths than it announces");
        std::sort(_lambdav.begin(), _lambdav.end());
        _dlambdav.resize(n);
        for (int i = 0; i < n; i++)
        {
            const double lo = i == 0 ? _lambdav[0] : std::sqrt(_lambdav[i - 1] * _lambdav[i]);
            const double hi = i == n - 1 ? _lambdav[n - 1] : std::sqrt(_lambdav[i] * _lambdav[i + 1]);
            _dlambdav[i] = hi - lo;
        }
    }
    bool issampledrange() const override { return true; }
private:
    std::string _filename;
};

// ---- one-dimensional meshes (Mesh subclasses; NR.hpp:171-261) -------------------------------------------
class Mesh
{
public:
    virtual ~Mesh() {}
    void setNumBins(int n) { if (n <= 0) SKIRT_FATAL("the number of bins in the mesh should be positive"); _N = n; }
    int numBins() const { return _N; }
    virtual std::vector<double> mesh() const = 0;      // N+1 borders from 0 to 1
protected:
    int _N = 0;
};
class LinMesh : public Mesh
{
public:
    std::vector<double> mesh() const override
    { std::vector<double> v(_N + 1); for (int i = 0; i <= _N; i++) v[i] = 0.0 + i * ((1.0 - 0.0) / _N); return v; }
};
class PowMesh : public Mesh       // NR::powgrid, NR.hpp:205-221
{
public:
    void setRatio(double r) { if (r <= 0) SKIRT_FATAL("the bin width ratio should be positive"); _ratio = r; }
    std::vector<double> mesh() const override
    {
        if (std::fabs(_ratio - 1.) < 1e-3) { LinMesh l; l.setNumBins(_N); return l.mesh(); }
        std::vector<double> v(_N + 1);
        double q = std::pow(_ratio, 1. / (_N - 1));
        for (int i = 0; i <= _N; ++i) v[i] = 0.0 + (1. - std::pow(q, i)) / (1. - std::pow(q, _N)) * (1.0 - 0.0);
        return v;
    }
private:
    double _ratio = 1;
};
class LogMesh : public Mesh       // LogMesh.cpp:47-53; NR::zerologgrid, NR.hpp:283-289 (an anchored mesh: radial coordinates)
{
public:
    void setCentralBinFraction(double tc) { if (tc <= 0 || tc >= 1) SKIRT_FATAL("The central bin width fraction should be within range ]0,1["); _tc = tc; }
    std::vector<double> mesh() const override
    {
        if (_N <= 1) { LinMesh l; l.setNumBins(1); return l.mesh(); }
        if (_tc <= 0 || _tc >= 1) SKIRT_FATAL("The central bin width fraction should be within range ]0,1[");
        std::vector<double> v(_N + 1, 0.0);
        const double logxmin = std::log10(_tc), dlogx = std::log10(1.0 / _tc) / (_N - 1);
        for (int i = 0; i < _N; i++) v[i + 1] = std::pow(10, logxmin + i * dlogx);
        return v;
    }
private:
    double _tc = 0;
};
class SymPowMesh : public Mesh    // NR::sympowgrid, NR.hpp:225-261
{
public:
    void setRatio(double r) { if (r <= 0) SKIRT_FATAL("the bin width ratio should be positive"); _ratio = r; }
    std::vector<double> mesh() const override
    {
        int n = _N; double ratio = _ratio;
        if (std::fabs(ratio - 1.) < 1e-3) { LinMesh l; l.setNumBins(n); return l.mesh(); }
        std::vector<double> xv(n + 1); const double xmin = 0, xmax = 1, xc = 0.5 * (xmin + xmax);
        if (n % 2 == 0)
        {
            int M = n / 2; double q = std::pow(ratio, 1. / (M - 1)), qM = std::pow(q, M);
            xv[M] = xc;
            for (int i = 1; i <= M; ++i) { double dxi = (1. - std::pow(q, i)) / (1. - qM) * 0.5 * (xmax - xmin); xv[M + i] = xc + dxi; xv[M - i] = xc - dxi; }
        }
        else
        {
            int M = (n + 1) / 2; double q = std::pow(ratio, 1. / (M - 1)), qM = std::pow(q, M);
            for (int i = 1; i <= M; ++i) { double dxi = (0.5 + 0.5 * q - std::pow(q, i)) / (0.5 + 0.5 * q - qM) * 0.5 * (xmax - xmin); xv[M - 1 + i] = xc + dxi; xv[M - i] = xc - dxi; }
        }
        return xv;
    }
private:
    double _ratio = 1;
};

// ---- geometries ---------------------------------------------------------------------------------------------
class Geometry
{
public:
    virtual ~Geometry() {}
    virtual void setup() {}
    virtual double density(double x, double y, double z) const = 0;
    virtual double SigmaZ() const = 0;
    virtual skg_source sampler() const = 0;        // the launch sampler of StellarSystem::launch on the device
};

class ExpDiskGeometry : public Geometry          // ExpDiskGeometry.cpp:22-43,117-129,177-187
{
public:
    void setRadialScale(double v) { _hR = v; }
    void setAxialScale(double v) { _hz = v; }
    void setRadialTrunc(double v) { _Rmax = v; }
    void setAxialTrunc(double v) { _zmax = v; }
    void setInnerRadius(double v) { _Rmin = v; }
    void setup() override
    {
        if (_hR <= 0) SKIRT_FATAL("The radial scale length hR should be positive");
        if (_hz <= 0) SKIRT_FATAL("The axial scale height hz should be positive");
        if (_Rmax < 0) SKIRT_FATAL("The radial truncation length Rmax should be zero or positive");
        if (_zmax < 0) SKIRT_FATAL("The axial truncation length zmax should be zero or positive");
        double intphi = 2.0 * M_PI;
        double intz = _zmax > 0 ? -2.0 * _hz * std::expm1(-_zmax / _hz) : 2.0 * _hz;
        double tmin = _Rmin > 0 ? std::exp(-_Rmin / _hR) * (1.0 + _Rmin / _hR) : 1.0;
        double tmax = _Rmax > 0 ? std::exp(-_Rmax / _hR) * (1.0 + _Rmax / _hR) : 0.0;
        _rho0 = 1.0 / (_hR * _hR * (tmin - tmax) * intphi * intz);
    }
    double density(double x, double y, double z) const override
    {
        double R = std::sqrt(x * x + y * y), absz = std::fabs(z);
        if (_Rmax > 0.0 && R > _Rmax) return 0.0;
        if (_zmax > 0.0 && absz > _zmax) return 0.0;
        if (R < _Rmin) return 0.0;
        return _rho0 * std::exp(-R / _hR) * std::exp(-absz / _hz);
    }
    double SigmaZ() const override
    {
        if (_Rmin > 0) return 0.0;
        return _zmax > 0 ? -2.0 * _rho0 * _hz * std::expm1(-_zmax / _hz) : 2.0 * _rho0 * _hz;
    }
    skg_source sampler() const override
    { skg_source s{}; s.geometry = SKG_GEOM_EXPDISK; s.p[0] = _hR; s.p[1] = _hz; s.p[2] = _Rmax; s.p[3] = _zmax; s.p[4] = _Rmin; return s; }
private:
    double _hR = 0, _hz = 0, _Rmax = 0, _zmax = 0, _Rmin = 0, _rho0 = 0;
};

// SersicFunction.cpp:18-78: Sersic profile S(s) and cumulative mass M(s) tabulated on 101 logarithmic radii
class SersicFunction
{
public:
    explicit SersicFunction(double n)
    {
        if (n < 0.5 || n > 10.0) SKIRT_FATAL("The Sersic parameter should be between 0.5 and 10");
        double b = 2.0 * n - 1.0 / 3.0 + 4.0 / 405.0 / n + 46.0 / 25515.0 / (n * n) + 131.0 / 1148175.0 / (n * n * n);
        double I0 = std::pow(b, 2.0 * n) / (M_PI * std::tgamma(2.0 * n + 1));
        const int Ns = 101; sv.resize(Ns); Sv.resize(Ns); Mv.assign(Ns, 0.0);
        const double logsmin = -6.0, logsmax = 4.0, dlogs = (logsmax - logsmin) / (Ns - 1.0);
        const int Nu = 10000; const double tmax = 100.0, umax = std::sqrt((tmax + 1.0) * (tmax - 1.0)), du = umax / Nu;
        for (int i = 0; i < Ns; i++)
        {
            double s = std::pow(10.0, logsmin + i * dlogs); sv[i] = s;
            double alpha = b * std::pow(s, 1.0 / n), sum = 0.0;
            for (int j = 0; j <= Nu; j++)
            {
                double weight = (j == 0 || j == Nu) ? 0.5 : 1.0, u = j * du, u2 = u * u;
                double w = u > 1e-3 ? (std::pow(1.0 + u2, 2.0 * n) - 1.0) / u2
                                    : 2.0 * n + n * (2.0 * n - 1.0) * u2 + 2.0 / 3.0 * n * (2.0 * n - 1.0) * (n - 1.0) * u2 * u2;
                sum += weight * 2.0 * std::exp(-alpha * (1.0 + u2)) / std::sqrt(w);
            }
            Sv[i] = I0 * std::pow(b, n) * std::pow(alpha, 1.0 - n) / M_PI * du * sum;
        }
        for (int i = 1; i < Ns; i++)
        {
            double sum = 0.0, ds = (sv[i] - sv[i - 1]) / 32.0;
            for (int j = 0; j <= 32; j++) { double weight = (j == 0 || j == 32) ? 0.5 : 1.0, s = sv[i - 1] + j * ds; sum += weight * (*this)(s) * s * s * ds; }
            Mv[i] = Mv[i - 1] + 4.0 * M_PI * sum;
        }
        for (int i = 0; i < Ns; i++) Mv[i] /= Mv[Ns - 1];
    }
    double operator()(double s) const       // SersicFunction.cpp:82-93: log-log interpolation, clamped to the table
    {
        int Ns = (int)sv.size();
        if (s <= sv[0]) return Sv[0];
        if (s >= sv[Ns - 1]) return Sv[Ns - 1];
        int i = int(std::upper_bound(sv.begin(), sv.end(), s) - sv.begin()) - 1;
        i = std::max(0, std::min(Ns - 2, i));
        double x = std::log10(s), x1 = std::log10(sv[i]), x2 = std::log10(sv[i + 1]), f1 = std::log10(Sv[i]), f2 = std::log10(Sv[i + 1]);
        return std::pow(10.0, f1 + ((x - x1) / (x2 - x1)) * (f2 - f1));
    }
    std::vector<double> sv, Sv, Mv;
};

class SersicGeometry : public Geometry     // SersicGeometry.cpp:30-91 (+ SpheroidalGeometryDecorator.cpp:78-85 for q != 1)
{
public:
    void setIndex(double v) { _n = v; }
    void setRadius(double v) { _reff = v; }
    void setFlattening(double v) { _q = v; }
    void setup() override
    {
        if (_n <= 0.5 || _n > 10.0) SKIRT_FATAL("the Sersic index n should be between 0.5 and 10");
        if (_reff <= 0) SKIRT_FATAL("the effective radius should be positive");
        if (_q <= 0 || _q > 1) SKIRT_FATAL("the flattening parameter should be between 0 and 1");
        _fn.reset(new SersicFunction(_n));
    }
    double density(double x, double y, double z) const override
    { double r = std::sqrt(x * x + y * y + (z / _q) * (z / _q)); return (*_fn)(r / _reff) / (_reff * _reff * _reff) / _q; }
    double SigmaZ() const override { SKIRT_FATAL("SigmaZ of a Sersic geometry is not needed by the hot path"); }
    skg_source sampler() const override
    {
        skg_source s{}; s.geometry = SKG_GEOM_SERSIC; s.p[0] = _reff; s.p[1] = _q;
        s.ntab = (int)_fn->sv.size(); s.rv = _fn->sv.data(); s.Xv = _fn->Mv.data(); s.Sv = _fn->Sv.data();
        return s;
    }
private:
    double _n = 0, _reff = 0, _q = 1; std::unique_ptr<SersicFunction> _fn;
};

class SpiralStructureGeometryDecorator : public Geometry    // SpiralStructureGeometryDecorator.cpp:24-45,177-229
{
public:
    void setGeometry(Geometry* g) { _geometry.reset(g); }
    void setArms(int v) { _m = v; } void setPitch(double v) { _p = v; } void setRadius(double v) { _R0 = v; }
    void setPhase(double v) { _phi0 = v; } void setPerturbWeight(double v) { _w = v; } void setIndex(int v) { _N = v; }
    void setup() override
    {
        if (!_geometry) SKIRT_FATAL("the geometry to be decorated was not set");
        if (_m <= 0) SKIRT_FATAL("The number of spiral arms should be positive");
        if (_p <= 0 || _p >= M_PI / 2.) SKIRT_FATAL("The pitch angle should be between 0 and 90 degrees");
        if (_R0 <= 0) SKIRT_FATAL("The radius zero-point should be positive");
        if (_w <= 0 || _w > 1.) SKIRT_FATAL("The weight of the spiral perturbation should be between 0 and 1");
        if (_N < 0 || _N > 10) SKIRT_FATAL("The arm-interarm size ratio index should be between 0 and 10");
        _geometry->setup();
        _tanp = std::tan(_p); _CN = std::sqrt(M_PI) * std::tgamma(_N + 1.0) / std::tgamma(_N + 0.5);
    }
    double density(double x, double y, double z) const override
    {
        double R = std::sqrt(x * x + y * y), phi = std::atan2(y, x);
        double gamma = std::log(R / _R0) / _tanp + _phi0 + 0.5 * M_PI / _m;
        double pert = (1.0 - _w) + _w * _CN * std::pow(std::sin(0.5 * _m * (gamma - phi)), 2 * _N);
        return _geometry->density(x, y, z) * pert;
    }
    double SigmaZ() const override { return _geometry->SigmaZ(); }
    skg_source sampler() const override
    {
        skg_source s = _geometry->sampler();
        s.spiral_arms = _m; s.spiral_index = _N; s.spiral_pitch = _p; s.spiral_radius = _R0; s.spiral_phase = _phi0; s.spiral_weight = _w;
        return s;
    }
private:
    std::unique_ptr<Geometry> _geometry; int _m = 0, _N = 0; double _p = 0, _R0 = 0, _phi0 = 0, _w = 0, _tanp = 0, _CN = 0;
};

// ---- dust mixes -----------------------------------------------------------------------------------------------
class DustMix
{
public:
    virtual ~DustMix() {}
    virtual void setup(const WavelengthGrid& lg) = 0;
    std::vector<double> kappaabsv, kappascav, asymmparv;
    double kappaext(int ell) const { return kappaabsv[ell] + kappascav[ell]; }
};

// optical properties tabulated on any wavelength grid and resampled onto the simulation's grid like
// DustMix::addpopulation (DustMix.cpp:300-321): log-log for the opacities, log-lin for g
class TableDustMix : public DustMix
{
public:
    void setTable(const std::vector<double>& lambda, const std::vector<double>& kabs, const std::vector<double>& ksca, const std::vector<double>& g)
    { _lam = lambda; _kabs = kabs; _ksca = ksca; _g = g; }
    void setup(const WavelengthGrid& lg) override
    {
        size_t n = _lam.size();
        if (n < 1 || _kabs.size() != n || _ksca.size() != n || _g.size() != n) SKIRT_FATAL("dust mix table is incomplete");
        int N = lg.Nlambda();
        kappaabsv.resize(N); kappascav.resize(N); asymmparv.resize(N);
        for (int ell = 0; ell < N; ell++)
        {
            double lam = lg.lambda(ell);
            if (n == 1) { kappaabsv[ell] = _kabs[0]; kappascav[ell] = _ksca[0]; asymmparv[ell] = _g[0]; continue; }
            if (lam < _lam.front() * (1 - 0.5e-5) || lam > _lam.back() * (1 + 0.5e-5))
                SKIRT_FATAL("Properties for this dust population are only defined for wavelengths between the table limits");
            size_t i = std::upper_bound(_lam.begin(), _lam.end(), lam) - _lam.begin();
            i = std::max<size_t>(1, std::min(n - 1, i)) - 1;
            double t = (std::log10(lam) - std::log10(_lam[i])) / (std::log10(_lam[i + 1]) - std::log10(_lam[i]));
            auto loglog = [&](double f1, double f2) { return (f1 > 0 && f2 > 0) ? std::pow(10.0, std::log10(f1) + t * (std::log10(f2) - std::log10(f1))) : f1 + t * (f2 - f1); };
            kappaabsv[ell] = loglog(_kabs[i], _kabs[i + 1]); kappascav[ell] = loglog(_ksca[i], _ksca[i + 1]);
            asymmparv[ell] = _g[i] + t * (_g[i + 1] - _g[i]);
        }
    }
private:
    std::vector<double> _lam, _kabs, _ksca, _g;
};

// InterstellarDustMix (InterstellarDustMix.cpp:21-58) from the 256-point table shipped in skirt_b200/data
// (columns: lambda[m] kappa_abs kappa_sca g; produced by tools/make_dustmix_table.py)
class InterstellarDustMix : public TableDustMix
{
public:
    explicit InterstellarDustMix(const std::string& datafile)
    {
        std::ifstream in(datafile);
        if (!in) SKIRT_FATAL("Could not open the data file " + datafile);
        std::vector<double> lam, a, s, g; std::string line;
        while (std::getline(in, line))
        {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream is(line); double v[4];
            if (is >> v[0] >> v[1] >> v[2] >> v[3]) { lam.push_back(v[0]); a.push_back(v[1]); s.push_back(v[2]); g.push_back(v[3]); }
        }
        setTable(lam, a, s, g);
    }
};

// ---- dust grids -------------------------------------------------------------------------------------------------
class DustGrid
{
public:
    virtual ~DustGrid() {}
    virtual void setup() = 0;
    virtual int numCells() const = 0;
    virtual void upload(skg_engine* e) const = 0;                                          // replaces DustGrid::path et al.
    // set-up side: stratified sample points and volume of cell m (DustSystem::setSampleDensityBody, DustSystem.cpp:152-177,
    // uses random points; the engine only sees the resulting table)
    virtual void cellBox(int m, double b[6]) const = 0;
    // grids that need the engine to come into being or to sample the dust density (tree, Voronoi), or that bring their own
    // density field (adaptive mesh):
    virtual bool densityOnDevice() const { return false; }      // DustSystem::setSampleDensityBody through skg_sample_density
    virtual bool ownDensity(std::vector<double>&) const { return false; }
    virtual double weight(int) const { return 1.0; }           // DustGrid::weight(m): TwoPhaseDustGrid's density multiplier
    virtual void build(skg_engine*, const std::vector<skg_source>&, const std::vector<double>&, uint64_t) {}
    virtual std::vector<double> volumes() const
    { std::vector<double> v(numCells()); for (int m = 0; m < numCells(); m++) { double b[6]; cellBox(m, b); v[m] = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2]); } return v; }
};

class BoxDustGrid : public DustGrid
{
public:
    void setMinX(double v) { _ext[0] = v; } void setMaxX(double v) { _ext[1] = v; }
    void setMinY(double v) { _ext[2] = v; } void setMaxY(double v) { _ext[3] = v; }
    void setMinZ(double v) { _ext[4] = v; } void setMaxZ(double v) { _ext[5] = v; }
    void setup() override
    {
        if (_ext[1] <= _ext[0]) SKIRT_FATAL("The extent of the box should be positive in the X direction");
        if (_ext[3] <= _ext[2]) SKIRT_FATAL("The extent of the box should be positive in the Y direction");
        if (_ext[5] <= _ext[4]) SKIRT_FATAL("The extent of the box should be positive in the Z direction");
    }
protected:
    double _ext[6] = {0, 0, 0, 0, 0, 0};
};

// TreeDustGrid / OctTreeDustGrid / BinTreeDustGrid (TreeDustGrid.cpp:20-37 defaults, :50-233): grown level by level by
// skirt::TreeBuilder (GridBuilders.cpp), the dust mass of the candidate nodes estimated on the GPU (skg_sample_boxes)
class TreeDustGrid : public BoxDustGrid
{
public:
    enum SearchMethod { TopDown = 0, Neighbor = 1, Bookkeeping = 2 };
    explicit TreeDustGrid(int kind) : _kind(kind) {}
    void setMinLevel(int v) { _minlevel = v; } void setMaxLevel(int v) { _maxlevel = v; }
    void setSearchMethod(SearchMethod v) { _search = v; }
    void setSampleCount(int v) { _Nrandom = v; }
    void setMaxOpticalDepth(double v) { _maxOpticalDepth = v; }
    void setMaxMassFraction(double v) { _maxMassFraction = v; }
    void setMaxDensDispFraction(double v) { _maxDensDispFraction = v; }
    bool densityOnDevice() const override { return true; }
    void build(skg_engine* e, const std::vector<skg_source>& geoms, const std::vector<double>& norms, uint64_t seed) override;
    int numCells() const override { return _t.Ncells; }
    void upload(skg_engine* e) const override;
    void cellBox(int m, double b[6]) const override { const double* q = &_t.box[6 * (size_t)_cellNode[m]]; for (int c = 0; c < 6; c++) b[c] = q[c]; }
    const skirt::TreeTables& tables() const { return _t; }
private:
    int _kind, _minlevel = 2, _maxlevel = 6, _Nrandom = 100; SearchMethod _search = Neighbor;
    double _maxOpticalDepth = 0, _maxMassFraction = 1e-6, _maxDensDispFraction = 0;
    skirt::TreeTables _t; std::vector<int> _cellNode;
};
// Barycentric subdivision (OctTreeDustGrid.cpp:32-40, BinTreeDustGrid.cpp:41-52) needs the barycentre of a node's dust, which the
// device's box sampler does not return: asked for, it is refused rather than silently replaced by the regular subdivision
class OctTreeDustGrid : public TreeDustGrid
{
public:
    OctTreeDustGrid() : TreeDustGrid(0) {}
    void setBarycentric(bool v) { if (v) SKIRT_FATAL("barycentric subdivision is not supported by this host (use the regular octree)"); }
};
class BinTreeDustGrid : public TreeDustGrid
{
public:
    enum DirectionMethod { Alternating = 0, Barycenter = 1 };
    BinTreeDustGrid() : TreeDustGrid(1) {}
    void setDirectionMethod(DirectionMethod v) { if (v != Alternating) SKIRT_FATAL("the Barycenter direction method is not supported by this host (use Alternating)"); }
};

// Sphere1DDustGrid / Sphere2DDustGrid / Cylinder2DDustGrid (the grids with symmetries): the border arrays are the whole state
class SymmetricDustGrid : public DustGrid
{
public:
    bool densityOnDevice() const override { return true; }
    int numCells() const override { return _N2 > 0 ? _N1 * _N2 : _N1; }
    void cellBox(int, double b[6]) const override { for (int c = 0; c < 6; c++) b[c] = 0; }      // (no Cartesian box; densities are sampled on the device)
    std::vector<double> volumes() const override { return _volumes; }
protected:
    int _N1 = 0, _N2 = 0; std::vector<double> _v1, _v2, _cv, _volumes;
};
class Sphere1DDustGrid : public SymmetricDustGrid                        // Sphere1DDustGrid.cpp:24-33, :67-77
{
public:
    void setMaxR(double v) { _rmax = v; } void setMeshR(Mesh* m) { _meshr.reset(m); }
    void setup() override;
    void upload(skg_engine* e) const override;
private:
    double _rmax = 0; std::unique_ptr<Mesh> _meshr;
};
class Sphere2DDustGrid : public SymmetricDustGrid                        // Sphere2DDustGrid.cpp:27-75, :123-131
{
public:
    void setMaxR(double v) { _rmax = v; } void setMeshR(Mesh* m) { _meshr.reset(m); } void setMeshTheta(Mesh* m) { _mesht.reset(m); }
    void setup() override;
    void upload(skg_engine* e) const override;
private:
    double _rmax = 0; std::unique_ptr<Mesh> _meshr, _mesht;
};
class Cylinder2DDustGrid : public SymmetricDustGrid                      // Cylinder2DDustGrid.cpp:26-41, :85-93
{
public:
    void setMaxR(double v) { _Rmax = v; } void setMinZ(double v) { _zmin = v; } void setMaxZ(double v) { _zmax = v; }
    void setMeshR(Mesh* m) { _meshR.reset(m); } void setMeshZ(Mesh* m) { _meshz.reset(m); }
    void setup() override;
    void upload(skg_engine* e) const override;
private:
    double _Rmax = 0, _zmin = 0, _zmax = 0; std::unique_ptr<Mesh> _meshR, _meshz;
};

// ParticleTreeDustGrid (ParticleTreeDustGrid.cpp:76-152): an octree or binary tree grown around particle positions (one per
// line of a text file, or given directly); every leaf ends up with at most one particle; its own traversal (search = 3)
class ParticleTreeDustGrid : public BoxDustGrid
{
public:
    enum TreeType { OctTree = 0, BinTree = 1 };
    void setTreeType(TreeType v) { _kind = v; }
    void setExtraLevels(int v) { _extra = v; }
    void setParticleFile(const std::string& path) { _file = path; }
    void setParticles(const std::vector<double>& xyz) { _particles = xyz; }
    void setup() override;
    bool densityOnDevice() const override { return true; }
    int numCells() const override { return _t.Ncells; }
    void upload(skg_engine* e) const override;
    void cellBox(int m, double b[6]) const override { const double* q = &_t.box[6 * (size_t)_cellNode[m]]; for (int c = 0; c < 6; c++) b[c] = q[c]; }
    const skirt::TreeTables& tables() const { return _t; }
private:
    TreeType _kind = OctTree; int _extra = 0; std::string _file; std::vector<double> _particles;
    skirt::TreeTables _t; std::vector<int> _cellNode;
};

// AdaptiveMeshDustGrid + AdaptiveMeshDustDistribution over an adaptive mesh data file in the format of
// AdaptiveMeshAsciiFile (AdaptiveMeshAsciiFile.cpp:43-100): "! Nx Ny Nz" for a nonleaf, the field values for a leaf
class AdaptiveMeshDustGrid : public BoxDustGrid
{
public:
    void setAdaptiveMeshFile(const std::string& path) { _file = path; }
    void setMesh(const std::vector<int>& nxyz, const std::vector<double>& values) { _nxyz = nxyz; _values = values; }
    void setDensityIndex(int v) { _densityIndex = v; }
    void setDensityUnits(double v) { _units = v; }
    void setup() override;
    bool ownDensity(std::vector<double>& rho) const override
    { rho.resize(_t.Ncells); for (int m = 0; m < _t.Ncells; m++) rho[m] = std::max(_values[_t.fileIndex[m]], 0.0) * _units; return true; }
    int numCells() const override { return _t.Ncells; }
    void upload(skg_engine* e) const override;
    void cellBox(int m, double b[6]) const override { const double* q = &_t.box[6 * (size_t)_cellNode[m]]; for (int c = 0; c < 6; c++) b[c] = q[c]; }
    std::vector<double> volumes() const override { return _t.volume; }
private:
    std::string _file; int _densityIndex = 0; double _units = 1;
    std::vector<int> _nxyz; std::vector<double> _values;
    skirt::AMeshTables _t; std::vector<int> _cellNode;
};

// VoronoiDustGrid over particle positions read from a text file (x y z per line) or set directly
class VoronoiDustGrid : public BoxDustGrid
{
public:
    void setParticleFile(const std::string& path) { _file = path; }
    void setParticles(const std::vector<double>& xyz) { _particles = xyz; }
    void setup() override;
    bool densityOnDevice() const override { return true; }
    int numCells() const override { return _t.Ncells; }
    void upload(skg_engine* e) const override;
    void cellBox(int m, double b[6]) const override { for (int c = 0; c < 6; c++) b[c] = _t.cellBox[6 * (size_t)m + c]; }
    std::vector<double> volumes() const override { return _t.volume; }
private:
    std::string _file; std::vector<double> _particles;
    skirt::VoronoiTables _t;
};

class CartesianDustGrid : public DustGrid          // CartesianDustGrid.cpp:28-43
{
public:
    void setMinX(double v) { _xmin = v; } void setMaxX(double v) { _xmax = v; }
    void setMinY(double v) { _ymin = v; } void setMaxY(double v) { _ymax = v; }
    void setMinZ(double v) { _zmin = v; } void setMaxZ(double v) { _zmax = v; }
    void setMeshX(Mesh* m) { _meshx.reset(m); } void setMeshY(Mesh* m) { _meshy.reset(m); } void setMeshZ(Mesh* m) { _meshz.reset(m); }
    void setup() override
    {
        if (_xmax <= _xmin) SKIRT_FATAL("The extent of the box should be positive in the X direction");
        if (_ymax <= _ymin) SKIRT_FATAL("The extent of the box should be positive in the Y direction");
        if (_zmax <= _zmin) SKIRT_FATAL("The extent of the box should be positive in the Z direction");
        if (!_meshx || !_meshy || !_meshz) SKIRT_FATAL("the bin distribution was not set for all axes");
        auto scale = [](const Mesh& m, double lo, double hi) { std::vector<double> v = m.mesh(); for (double& t : v) t = t * (hi - lo) + lo; return v; };
        _xv = scale(*_meshx, _xmin, _xmax); _yv = scale(*_meshy, _ymin, _ymax); _zv = scale(*_meshz, _zmin, _zmax);
    }
    int numCells() const override { return (int)((_xv.size() - 1) * (_yv.size() - 1) * (_zv.size() - 1)); }
    void upload(skg_engine* e) const override;
    void cellBox(int m, double b[6]) const override
    {
        int Ny = (int)_yv.size() - 1, Nz = (int)_zv.size() - 1;
        int i = m / (Nz * Ny), j = (m / Nz) % Ny, k = m % Nz;          // CartesianDustGrid::box, :333-343
        b[0] = _xv[i]; b[1] = _yv[j]; b[2] = _zv[k]; b[3] = _xv[i + 1]; b[4] = _yv[j + 1]; b[5] = _zv[k + 1];
    }
protected:
    double _xmin = 0, _xmax = 0, _ymin = 0, _ymax = 0, _zmin = 0, _zmax = 0;
    std::unique_ptr<Mesh> _meshx, _meshy, _meshz;
    std::vector<double> _xv, _yv, _zv;
};

// TwoPhaseDustGrid (TwoPhaseDustGrid.cpp:18-39): a Cartesian grid whose cells belong at random to a high- or a low-density phase
class TwoPhaseDustGrid : public CartesianDustGrid
{
public:
    void setFillingFactor(double v) { _ff = v; } void setContrast(double v) { _contrast = v; } void setSeed(uint64_t s) { _seed = s; }
    void setup() override
    {
        CartesianDustGrid::setup();
        if (_ff <= 0 || _ff >= 1) SKIRT_FATAL("the volume filling factor of the high-density medium should be between 0 and 1");
        if (_contrast <= 0) SKIRT_FATAL("the density contrast between the high- and low-density medium should be positive");
        const double den = _contrast * _ff + 1.0 - _ff;
        _weightv.resize(numCells());
        uint64_t x = _seed * 6364136223846793005ull + 1442695040888963407ull;      // (a 64-bit LCG; the reference draws from its Random)
        for (double& w : _weightv) { x = x * 6364136223846793005ull + 1442695040888963407ull; w = ((x >> 11) * (1.0 / 9007199254740992.0)) < _ff ? _contrast / den : 1.0 / den; }
    }
    double weight(int m) const override { return m < 0 ? 0.0 : _weightv[m]; }
private:
    double _ff = 0, _contrast = 0; uint64_t _seed = 4357; std::vector<double> _weightv;
};

// ---- dust system --------------------------------------------------------------------------------------------------
class FaceOnDustCompNormalization      // FaceOnDustCompNormalization.cpp:67-74
{
public:
    void setWavelength(double v) { _lambda = v; }
    void setOpticalDepth(double v) { _tau = v; }
    double wavelength() const { return _lambda; } double opticalDepth() const { return _tau; }
private:
    double _lambda = 0, _tau = 0;
};

class DustComp
{
public:
    void setGeometry(Geometry* g) { geometry.reset(g); }
    void setMix(DustMix* m) { mix.reset(m); }
    void setNormalization(FaceOnDustCompNormalization* n) { norm.reset(n); }
    std::unique_ptr<Geometry> geometry; std::unique_ptr<DustMix> mix; std::unique_ptr<FaceOnDustCompNormalization> norm;
};

class DustSystem
{
public:
    void setDustGrid(DustGrid* g) { _grid.reset(g); }
    void addComponent(DustComp* c) { _comps.emplace_back(c); }
    void setSampleLattice(int n) { _nsub = n; }         // Cartesian grids: nsub^3 stratified points per cell stand in for setSampleCount
    void setSampleCount(int n) { _Nrandom = n; }        // tree / Voronoi grids: random positions per cell, drawn on the device
    void setStoreAbsorptionRates(bool v) { _storeabs = v; }
    bool storeabsorptionrates() const { return _storeabs; }
    int Ncells() const { return _grid->numCells(); }
    int Ncomp() const { return (int)_comps.size(); }
    DustGrid* dustGrid() const { return _grid.get(); }
    void presetup(const WavelengthGrid& lg);
    void setup(const WavelengthGrid& lg, skg_engine* e, uint64_t seed);
    void upload(skg_engine* e) const;
    const std::vector<double>& rho() const { return _rho; }
    const std::vector<double>& kappaabs() const { return _kabs; }       // [Ncomp*Nlambda]
    std::vector<double> volumes() const { return _grid->volumes(); }
private:
    std::unique_ptr<DustGrid> _grid; std::vector<std::unique_ptr<DustComp>> _comps;
    int _nsub = 2, _Nrandom = 100; bool _storeabs = false, _presetup = false; int _Nlambda = 0;
    std::vector<double> _rho, _kext, _ksca, _g, _kabs;
};

// ---- stellar system ------------------------------------------------------------------------------------------------
class StellarComp       // GeometricStellarComp with luminosities per wavelength bin
{
public:
    void setGeometry(Geometry* g) { geometry.reset(g); }
    void setLuminosities(const std::vector<double>& L) { Lv = L; }
    // PanStellarComp with a BlackBodySED and bolometric normalisation: L_ell = Lbol * B(lambda_ell,T) dlambda_ell / sum
    void setBlackBody(double T, double Lbol) { _T = T; _Lbol = Lbol; }
    void setup(const WavelengthGrid& lg)
    {
        if (!geometry) SKIRT_FATAL("the geometry of the stellar component was not set");
        geometry->setup();
        if (_T > 0)
        {
            int N = lg.Nlambda(); Lv.resize(N); double sum = 0;
            for (int ell = 0; ell < N; ell++)
            {
                double lam = lg.lambda(ell), x = Units::h * Units::c / (lam * Units::k * _T);
                Lv[ell] = 2.0 * Units::h * Units::c * Units::c / std::pow(lam, 5) / std::expm1(x) * lg.dlambda(ell); sum += Lv[ell];
            }
            for (double& v : Lv) v *= _Lbol / sum;
        }
        if ((int)Lv.size() != lg.Nlambda()) SKIRT_FATAL("the number of luminosities differs from the number of wavelengths");
    }
    std::unique_ptr<Geometry> geometry; std::vector<double> Lv;
private:
    double _T = 0, _Lbol = 0;
};

class StellarSystem
{
public:
    void addComponent(StellarComp* c) { _comps.emplace_back(c); }
    void setEmissionBias(double v) { if (v < 0 || v > 1) SKIRT_FATAL("the emission bias should be between 0 and 1"); _emissionBias = v; }
    void setup(const WavelengthGrid& lg) { if (_comps.empty()) SKIRT_FATAL("There are no stellar components"); for (auto& c : _comps) c->setup(lg); _Nlambda = lg.Nlambda(); }
    void upload(skg_engine* e) const;
    double luminosity(int ell) const { double s = 0; for (auto& c : _comps) s += c->Lv[ell]; return s; }
    int Ncomp() const { return (int)_comps.size(); }
private:
    std::vector<std::unique_ptr<StellarComp>> _comps; double _emissionBias = 0.5; int _Nlambda = 0;
};

// ---- instruments ------------------------------------------------------------------------------------------------------
class Instrument       // DistantInstrument + SingleFrameInstrument properties
{
public:
    virtual ~Instrument() {}
    virtual int kind() const = 0;
    void setInstrumentName(const std::string& v) { name = v; }
    void setDistance(double v) { d.distance = v; } void setInclination(double v) { d.inclination = v; }
    void setAzimuth(double v) { d.azimuth = v; } void setPositionAngle(double v) { d.positionAngle = v; }
    void setPixelsX(int v) { d.Nxp = v; } void setPixelsY(int v) { d.Nyp = v; }
    void setFieldOfViewX(double v) { d.fovxp = v; } void setFieldOfViewY(double v) { d.fovyp = v; }
    void setCenterX(double v) { d.xpc = v; } void setCenterY(double v) { d.ypc = v; }
    skg_instrument descriptor() const { skg_instrument s = d; s.kind = kind(); return s; }
    virtual int channels() const { return 0; }     // FullInstrument: 5 + scatteringLevels separate detector arrays
    std::string name;
    std::vector<double> ftotv, Ftotv;      // the detector arrays that Instrument::write() calibrates and saves (filled by fetch)
protected:
    skg_instrument d{};
};
class FrameInstrument : public Instrument { public: int kind() const override { return SKG_INSTR_FRAME; } };
class SEDInstrument : public Instrument { public: int kind() const override { return SKG_INSTR_SED; } };
class SimpleInstrument : public Instrument { public: int kind() const override { return SKG_INSTR_SIMPLE; } };
// FullInstrument (FullInstrument.cpp, unpolarised): transparent / direct / scattered stellar and direct / scattered dust
// emission flux in separate data cubes and SEDs, plus one per scattering level; channel order = SKG_CHAN_*
class FullInstrument : public Instrument
{
public:
    int kind() const override { return SKG_INSTR_FULL; }
    void setScatteringLevels(int v) { if (v < 0) SKIRT_FATAL("the number of scattering levels should be zero or positive"); d.scatteringLevels = v; }
    int scatteringLevels() const { return d.scatteringLevels; }
    int channels() const override { return SKG_CHAN_SCATTERING_LEVEL1 + d.scatteringLevels; }
    std::vector<std::vector<double>> fchanv, Fchanv;     // [channel][...] raw detector arrays (filled by fetch)
};

// MultiFrameInstrument (MultiFrameInstrument.cpp) with its InstrumentFrame items (InstrumentFrame.cpp:22-44): one pixel grid per
// wavelength; the total flux and / or the flux of every stellar component, one FITS file per array and wavelength
struct InstrumentFrame
{
    void setPixelsX(int v) { f.Nxp = v; } void setPixelsY(int v) { f.Nyp = v; }
    void setFieldOfViewX(double v) { f.fovxp = v; } void setFieldOfViewY(double v) { f.fovyp = v; }
    void setCenterX(double v) { f.xpc = v; } void setCenterY(double v) { f.ypc = v; }
    skg_instrument_frame f{};
};
class MultiFrameInstrument : public Instrument
{
public:
    MultiFrameInstrument() { d.writeTotal = 1; }
    int kind() const override { return SKG_INSTR_MULTIFRAME; }
    void setWriteTotal(bool v) { d.writeTotal = v ? 1 : 0; } void setWriteStellarComps(bool v) { d.writeStellarComps = v ? 1 : 0; }
    void addFrame(const InstrumentFrame& fr) { _frames.push_back(fr.f); }
    const std::vector<skg_instrument_frame>& frames() const { return _frames; }
    // filled by fetch: arrays[ell] = the arrays of frame ell in the order of InstrumentFrame::calibrateAndWriteData, with their names
    std::vector<std::vector<std::vector<double>>> arrays; std::vector<std::string> arrayNames;
    skg_instrument descriptorWithFrames() const { skg_instrument s = descriptor(); s.frames = _frames.data(); return s; }
private:
    std::vector<skg_instrument_frame> _frames;
};

// PerspectiveInstrument (PerspectiveInstrument.cpp): a pinhole camera; setWidth is the viewport width (the pixels are square)
class PerspectiveInstrument : public Instrument
{
public:
    int kind() const override { return SKG_INSTR_PERSPECTIVE; }
    void setWidth(double v) { d.fovxp = v; }
    void setViewX(double v) { d.viewX = v; } void setViewY(double v) { d.viewY = v; } void setViewZ(double v) { d.viewZ = v; }
    void setCrossX(double v) { d.crossX = v; } void setCrossY(double v) { d.crossY = v; } void setCrossZ(double v) { d.crossZ = v; }
    void setUpX(double v) { d.upX = v; } void setUpY(double v) { d.upY = v; } void setUpZ(double v) { d.upZ = v; }
    void setFocal(double v) { d.focal = v; }
};

class InstrumentSystem
{
public:
    void addInstrument(Instrument* i) { _instruments.emplace_back(i); }
    const std::vector<std::unique_ptr<Instrument>>& instruments() const { return _instruments; }
    void upload(skg_engine* e) const;
private:
    std::vector<std::unique_ptr<Instrument>> _instruments;
};

// ---- the simulation -------------------------------------------------------------------------------------------------------
// MonteCarloSimulation (MonteCarloSimulation.cpp:31-36 defaults, :251-261 runstellaremission).  One object per process /
// GPU; with several processes the packet budget is block-split like IdenticalAssigner does (IdenticalAssigner.cpp:37-58)
// and the detector arrays / absorption table are summed with NCCL (setCommunicator).
class MonteCarloSimulation
{
public:
    MonteCarloSimulation() {}
    ~MonteCarloSimulation() { if (_engine) skg_engine_destroy(_engine); }
    void setWavelengthGrid(WavelengthGrid* v) { _lambdagrid.reset(v); }
    void setStellarSystem(StellarSystem* v) { _ss.reset(v); }
    void setDustSystem(DustSystem* v) { _ds.reset(v); }
    void setInstrumentSystem(InstrumentSystem* v) { _is.reset(v); }
    void setPackages(double v) { if (v < 0) SKIRT_FATAL("Number of photon packages is negative"); if (v > 1e15) SKIRT_FATAL("Number of photon packages is larger than implementation limit of 1e15"); _packages = v; }
    void setMinWeightReduction(double v) { if (v < 1e3) SKIRT_FATAL("The minimum weight reduction factor should be larger than 1000"); _minWeightReduction = v; }
    void setMinScattEvents(double v) { if (v < 0 || v > 1000) SKIRT_FATAL("The minimum number of forced scattering events should be between 0 and 1000"); _minfs = v; }
    void setScattBias(double v) { if (v < 0 || v > 1) SKIRT_FATAL("The scattering bias should be between 0 and 1"); _xi = v; }
    void setContinuousScattering(bool v) { _continuousScattering = v; }
    void setSeed(int v) { _seed = v; }                                   // Random::setSeed
    void setDevice(int v) { _device = v; }
    // rank/size of the process group and the NCCL id shared by its members (replaces PeerToPeerCommunicator)
    void setCommunicator(int rank, int nranks, const void* ncclUniqueId128) { _rank = rank; _nranks = nranks; _uid = ncclUniqueId128; }
    double packages() const { return _packages; }
    skg_engine* engine() const { return _engine; }

    void setup();
    skg_mc_stats runstellaremission();
    // PanMonteCarloSimulation (PanMonteCarloSimulation.cpp:105-264); the emission spectra between the phases come from
    // the engine's dust library (AllCellsDustLib + GreyBodyDustEmissivity on the device)
    void setDustEmission(bool v) { _dustemission = v; }
    void setSelfAbsorption(bool v) { _selfabsorption = v; }
    void setCycles(int v) { _cycles = v; }
    void setEmissionBias(double v) { _dustBias = v; }
    void setEmissionBoost(double v) { _dustBoost = v; }
    bool dustemission() const { return _dustemission; }
    int rundustselfabsorption();         // returns the number of cycles performed
    int rundustselfabsorptionIfEnabled() { return _selfabsorption ? rundustselfabsorption() : 0; }
    skg_mc_stats rundustemission();
    void run() { runstellaremission(); if (_dustemission) { if (_selfabsorption) rundustselfabsorption(); rundustemission(); } }
    void fetchResults();                // fills Instrument::ftotv / Ftotv and Labs() on the host (what write() consumes)
    const std::vector<double>& Labs() const { return _Labs; }
    InstrumentSystem* instrumentSystem() const { return _is.get(); }
    WavelengthGrid* wavelengthGrid() const { return _lambdagrid.get(); }
    DustSystem* dustSystem() const { return _ds.get(); }
private:
    std::unique_ptr<WavelengthGrid> _lambdagrid; std::unique_ptr<StellarSystem> _ss; std::unique_ptr<DustSystem> _ds; std::unique_ptr<InstrumentSystem> _is;
    double _packages = 1e6, _minWeightReduction = 1e4, _minfs = 0, _xi = 0.5; int _seed = 4357, _device = 0;
    int _rank = 0, _nranks = 1; const void* _uid = nullptr;
    bool _continuousScattering = false;
    bool _dustemission = false, _selfabsorption = false; int _cycles = 0; double _dustBias = 0.5, _dustBoost = 1.0; int _phaseCounter = 0;
    skg_mc_stats shootDust(int phase, double packages);
    skg_engine* _engine = nullptr; std::vector<double> _Labs;
};

}   // namespace skirt
