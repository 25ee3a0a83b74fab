#include "Output.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>

namespace skirt
{

namespace
{
    const double pc = 3.08567758e16, AU = 1.49597871e11;                    // Units.cpp:21-22
    const double arcsec2 = std::pow(M_PI / (180. * 3600.), 2);
}

SIUnits::SIUnits()
{
    _ulength = "m"; _uwavelength = "m"; _unfd = "W/m2"; _unsb = "W/m2/sr"; _uwfd = "W/m3"; _uwsb = "W/m3/sr"; _uffd = "W/m2/Hz"; _ufsb = "W/m2/Hz/sr";
}
StellarUnits::StellarUnits()
{
    _ulength = "AU"; _uwavelength = "micron"; _unfd = "W/m2"; _unsb = "W/m2/arcsec2"; _uwfd = "W/m2/micron"; _uwsb = "W/m2/micron/arcsec2";
    _uffd = "Jy"; _ufsb = "MJy/sr";
    _clength = AU; _cwavelength = 1e-6; _cnsb = 1. / arcsec2; _cwfd = 1e6; _cwsb = 1e6 / arcsec2; _cffd = 1e-26; _cfsb = 1e-20;
}
ExtragalacticUnits::ExtragalacticUnits()
{
    _ulength = "pc"; _uwavelength = "micron"; _unfd = "W/m2"; _unsb = "W/m2/arcsec2"; _uwfd = "W/m2/micron"; _uwsb = "W/m2/micron/arcsec2";
    _uffd = "Jy"; _ufsb = "MJy/sr";
    _clength = pc; _cwavelength = 1e-6; _cnsb = 1. / arcsec2; _cwfd = 1e6; _cwsb = 1e6 / arcsec2; _cffd = 1e-26; _cfsb = 1e-20;
}

double UnitSystem::ofluxdensity(double lambda, double Flambda) const       // Units.cpp:995-1004
{
    switch (_style)
    {
    case Wavelength: return Flambda / _cwfd;
    case Frequency: return (lambda * lambda * Flambda / Units::c) / _cffd;
    default: return (lambda * Flambda) / _cnfd;
    }
}
double UnitSystem::osurfacebrightness(double lambda, double flambda) const  // Units.cpp:1033-1041
{
    switch (_style)
    {
    case Wavelength: return flambda / _cwsb;
    case Frequency: return (lambda * lambda * flambda / Units::c) / _cfsb;
    default: return (lambda * flambda) / _cnsb;
    }
}

static std::vector<double> calibrateCube(const Instrument& ins, const std::vector<double>& raw, const WavelengthGrid& lg, const UnitSystem& units)
{
    const skg_instrument d = ins.descriptor();
    const int Nl = lg.Nlambda(); const size_t Nframe = (size_t)d.Nxp * d.Nyp;
    if (raw.size() != Nframe * Nl) SKIRT_FATAL("the data cube of instrument " + ins.name + " has not been fetched");
    std::vector<double> f = raw;
    // step 1: W -> W/m
    for (int ell = 0; ell < Nl; ell++) { const double dl = lg.dlambda(ell); for (size_t l = 0; l < Nframe; l++) f[l + Nframe * ell] /= dl; }
    // step 2: per steradian
    const double xpsiz = d.fovxp / d.Nxp, ypsiz = d.fovyp / d.Nyp;
    const double area = (2.0 * std::atan(xpsiz / (2.0 * d.distance))) * (2.0 * std::atan(ypsiz / (2.0 * d.distance)));
    for (double& v : f) v /= area;
    // step 3: flux density at the observer
    const double fourpid2 = 4.0 * M_PI * d.distance * d.distance;
    for (double& v : f) v /= fourpid2;
    // output units
    for (int ell = 0; ell < Nl; ell++) { const double lam = lg.lambda(ell); for (size_t l = 0; l < Nframe; l++) f[l + Nframe * ell] = units.osurfacebrightness(lam, f[l + Nframe * ell]); }
    return f;
}

std::vector<double> calibrateDataCube(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units)
{ return calibrateCube(ins, ins.ftotv, lg, units); }

static std::vector<double> calibrateFluxes(const Instrument& ins, const std::vector<double>& raw, const WavelengthGrid& lg, const UnitSystem& units)
{
    const skg_instrument d = ins.descriptor();
    const int Nl = lg.Nlambda();
    if ((int)raw.size() != Nl) SKIRT_FATAL("the SED of instrument " + ins.name + " has not been fetched");
    std::vector<double> F = raw;
    for (int ell = 0; ell < Nl; ell++) F[ell] /= lg.dlambda(ell);
    const double fourpid2 = 4.0 * M_PI * d.distance * d.distance;
    for (double& v : F) v /= fourpid2;
    for (int ell = 0; ell < Nl; ell++) F[ell] = units.ofluxdensity(lg.lambda(ell), F[ell]);
    return F;
}
std::vector<double> calibrateSED(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units)
{ return calibrateFluxes(ins, ins.Ftotv, lg, units); }

namespace
{
    std::string pad80(std::string s) { s.resize(80, ' '); return s; }
    std::string cardLogical(const char* key, bool v, const char* comment)
    { char b[128]; std::snprintf(b, sizeof b, "%-8s= %20s / %s", key, v ? "T" : "F", comment); return pad80(b); }
    std::string cardInt(const char* key, long v, const char* comment)
    { char b[128]; std::snprintf(b, sizeof b, "%-8s= %20ld / %s", key, v, comment); return pad80(b); }
    std::string cardDouble(const char* key, double v, const char* comment)
    {
        char num[64]; std::snprintf(num, sizeof num, "%.15G", v);
        if (!std::strchr(num, 'E') && !std::strchr(num, '.')) std::strcat(num, ".");
        char b[160];
        if (comment[0]) std::snprintf(b, sizeof b, "%-8s= %20s / %s", key, num, comment); else std::snprintf(b, sizeof b, "%-8s= %20s", key, num);
        return pad80(b);
    }
    std::string cardString(const char* key, const std::string& v, const char* comment)
    {
        std::string q = "'"; for (char c : v) { q += c; if (c == '\'') q += c; }
        while (q.size() < 9) q += ' ';
        q += "'";
        char b[200]; std::snprintf(b, sizeof b, "%-8s= %-20s / %s", key, q.c_str(), comment); return pad80(b);
    }
}

void writeFITS(const std::string& path, const std::vector<double>& data, int nx, int ny, int nz, double incx, double incy,
               double xc, double yc, const std::string& dataUnits, const std::string& xyUnits, const std::string& stampIn)
{
    if (data.size() != (size_t)nx * ny * nz) SKIRT_FATAL("Inconsistent data size when creating FITS file " + path);
    std::string stamp = stampIn;
    if (stamp.empty()) { char b[32]; std::time_t t = std::time(nullptr); std::strftime(b, sizeof b, "%Y-%m-%dT%H:%M:%S", std::gmtime(&t)); stamp = b; }
    std::string h;
    h += cardLogical("SIMPLE", true, "file does conform to FITS standard");
    h += cardInt("BITPIX", -32, "number of bits per data pixel");
    h += cardInt("NAXIS", nz == 1 ? 2 : 3, "number of data axes");
    h += cardInt("NAXIS1", nx, "length of data axis 1");
    h += cardInt("NAXIS2", ny, "length of data axis 2");
    if (nz != 1) h += cardInt("NAXIS3", nz, "length of data axis 3");
    h += cardLogical("EXTEND", true, "FITS dataset may contain extensions");
    h += pad80("COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy");
    h += pad80("COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H");
    h += cardDouble("BSCALE", 1.0, ""); h += cardDouble("BZERO", 0.0, "");
    h += cardString("DATE", stamp, "Date and time of creation (UTC)");
    h += cardString("ORIGIN", "SKIRT simulation", "Astronomical Observatory, Ghent University");
    h += cardString("BUNIT", dataUnits, "Physical unit of the array values");
    h += cardDouble("CRPIX1", (nx + 1.0) / 2.0, "X-axis coordinate system reference pixel");
    h += cardDouble("CRVAL1", xc, "Coordinate system value at X-axis reference pixel");
    h += cardDouble("CDELT1", incx, "Coordinate increment along X-axis");
    h += cardString("CTYPE1", xyUnits, "Physical units of the X-axis increment");
    h += cardDouble("CRPIX2", (ny + 1.0) / 2.0, "Y-axis coordinate system reference pixel");
    h += cardDouble("CRVAL2", yc, "Coordinate system value at Y-axis reference pixel");
    h += cardDouble("CDELT2", incy, "Coordinate increment along Y-axis");
    h += cardString("CTYPE2", xyUnits, "Physical units of the Y-axis increment");
    h += pad80("END");
    h.resize((h.size() + 2879) / 2880 * 2880, ' ');
    std::string body(data.size() * 4, '\0');
    for (size_t i = 0; i < data.size(); i++)
    {
        const float v = (float)data[i]; unsigned u; std::memcpy(&u, &v, 4);
        body[4 * i] = (char)(u >> 24); body[4 * i + 1] = (char)(u >> 16); body[4 * i + 2] = (char)(u >> 8); body[4 * i + 3] = (char)u;     // big-endian
    }
    body.resize((body.size() + 2879) / 2880 * 2880, '\0');
    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    if (!out) SKIRT_FATAL("Error while creating FITS file " + path);
    out.write(h.data(), h.size()); out.write(body.data(), body.size());
    if (!out) SKIRT_FATAL("Error while writing FITS file " + path);
}

void writeSEDs(const std::string& path, const WavelengthGrid& lg, const std::vector<std::vector<double>>& Fs,
               const std::vector<std::string>& columnNames, const UnitSystem& units)
{
    std::ofstream out(path, std::ios::trunc);
    if (!out) SKIRT_FATAL("cannot create " + path);
    out << "# column 1: lambda (" << units.uwavelength() << ")\n";
    for (size_t q = 0; q < Fs.size(); q++)
        out << "# column " << q + 2 << ": " << columnNames[q] << "; " << units.sfluxdensity() << " (" << units.ufluxdensity() << ")\n";
    char b[32];
    for (int ell = 0; ell < lg.Nlambda(); ell++)
    {
        std::snprintf(b, sizeof b, "%.8e", units.owavelength(lg.lambda(ell))); out << b;       // QString::number(v, 'e', 8)
        for (const auto& F : Fs) { std::snprintf(b, sizeof b, " %.8e", F.empty() ? 0.0 : F[ell]); out << b; }
        out << "\n";
    }
}
void writeSED(const std::string& path, const WavelengthGrid& lg, const std::vector<double>& F, const std::string& columnName, const UnitSystem& units)
{ writeSEDs(path, lg, {F}, {columnName}, units); }

void writeInstrument(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units, const std::string& prefix, const std::string& stamp,
                     bool dustsystem, bool dustemission)
{
    const skg_instrument d = ins.descriptor();
    if (d.kind == SKG_INSTR_PERSPECTIVE)
    {
        // PerspectiveInstrument::write (PerspectiveInstrument.cpp:354-397): every sample times 1/(4 pi s^2) / dlambda, to output units
        const double sp = d.fovxp / d.Nxp, front = 1. / (4. * M_PI * sp * sp);
        const size_t Nf = (size_t)d.Nxp * d.Nyp;
        if (ins.ftotv.size() != Nf * lg.Nlambda()) SKIRT_FATAL("the detector array of instrument " + ins.name + " has not been fetched");
        std::vector<double> a = ins.ftotv;
        for (int ell = 0; ell < lg.Nlambda(); ell++)
            for (size_t l = 0; l < Nf; l++) a[l + Nf * ell] = units.osurfacebrightness(lg.lambda(ell), a[l + Nf * ell] * front / lg.dlambda(ell));
        writeFITS(prefix + "_" + ins.name + "_total.fits", a, d.Nxp, d.Nyp, lg.Nlambda(), units.olength(sp), units.olength(sp), 0., 0.,
                  units.usurfacebrightness(), units.ulength(), stamp);
        return;
    }
    if (const MultiFrameInstrument* mf = dynamic_cast<const MultiFrameInstrument*>(&ins))
    {
        // InstrumentFrame::calibrateAndWriteDataFrames (InstrumentFrame.cpp:216-262): per wavelength, every array divided by
        // dlambda * pixel solid angle * 4 pi d^2, converted to the output units, one FITS file per array
        if ((int)mf->arrays.size() != lg.Nlambda()) SKIRT_FATAL("the detector arrays of instrument " + ins.name + " have not been fetched");
        for (int ell = 0; ell < lg.Nlambda(); ell++)
        {
            const skg_instrument_frame& f = mf->frames()[ell];
            const double xpsiz = f.fovxp / f.Nxp, ypsiz = f.fovyp / f.Nyp;
            const double area = (2.0 * std::atan(xpsiz / (2.0 * d.distance))) * (2.0 * std::atan(ypsiz / (2.0 * d.distance)));
            const double fourpid2 = 4.0 * M_PI * d.distance * d.distance;
            const double factor = units.osurfacebrightness(lg.lambda(ell), 1.) / (lg.dlambda(ell) * area * fourpid2);
            for (size_t q = 0; q < mf->arrays[ell].size(); q++)
            {
                std::vector<double> a = mf->arrays[ell][q];
                for (double& v : a) v *= factor;
                writeFITS(prefix + "_" + ins.name + "_" + mf->arrayNames[q] + "_" + std::to_string(ell) + ".fits", a, f.Nxp, f.Nyp, 1,
                          units.olength(xpsiz), units.olength(ypsiz), f.xpc, f.ypc, units.usurfacebrightness(), units.ulength(), stamp);
            }
        }
        return;
    }
    if (const FullInstrument* fi = dynamic_cast<const FullInstrument*>(&ins))
    {
        // FullInstrument::write, FullInstrument.cpp:176-236: total = direct + scattered (+ dust); empty arrays are skipped as
        // data cubes (calibrateAndWriteDataCubes, SingleFrameInstrument.cpp:151-226) and written as zero SED columns
        if ((int)fi->fchanv.size() != fi->channels() || (int)fi->Fchanv.size() != fi->channels())
            SKIRT_FATAL("the detector arrays of instrument " + ins.name + " have not been fetched");
        typedef std::vector<double> V;
        auto sum = [](const V& a, const V& b) { V r(a); for (size_t i = 0; i < r.size(); i++) r[i] += b[i]; return r; };
        const V none;
        const V &ftra = fi->fchanv[SKG_CHAN_TRANSPARENT], &Ftra = fi->Fchanv[SKG_CHAN_TRANSPARENT];
        const V &fdir = fi->fchanv[SKG_CHAN_STELLAR_DIRECT], &Fdir = fi->Fchanv[SKG_CHAN_STELLAR_DIRECT];
        const V &fsca = fi->fchanv[SKG_CHAN_STELLAR_SCATTERED], &Fsca = fi->Fchanv[SKG_CHAN_STELLAR_SCATTERED];
        const V &fdd = fi->fchanv[SKG_CHAN_DUST_DIRECT], &Fdd = fi->Fchanv[SKG_CHAN_DUST_DIRECT];
        const V &fds = fi->fchanv[SKG_CHAN_DUST_SCATTERED], &Fds = fi->Fchanv[SKG_CHAN_DUST_SCATTERED];
        std::vector<V> f, F; std::vector<std::string> fn, Fn;
        auto add = [&](const V& fa, const V& Fa, const char* fname, const std::string& Fname) { f.push_back(fa); F.push_back(Fa); fn.push_back(fname); Fn.push_back(Fname); };
        if (dustemission) add(sum(sum(fdir, fsca), sum(fdd, fds)), sum(sum(Fdir, Fsca), sum(Fdd, Fds)), "total", "total flux");
        else if (dustsystem) add(sum(fdir, fsca), sum(Fdir, Fsca), "total", "total flux");
        else add(ftra, Ftra, "total", "total flux");
        add(dustsystem ? fdir : none, dustsystem ? Fdir : Ftra, "direct", "direct stellar flux");
        add(dustsystem ? fsca : none, dustsystem ? Fsca : none, "scattered", "scattered stellar flux");
        add(dustemission ? sum(fdd, fds) : none, dustemission ? sum(Fdd, Fds) : none, "dust", "total dust emission flux");
        add(dustemission ? fds : none, dustemission ? Fds : none, "dustscattered", "dust emission scattered flux");
        add(dustsystem ? ftra : none, Ftra, "transparent", "transparent flux");
        for (int n = 0; n < fi->scatteringLevels(); n++)
            add(dustsystem ? fi->fchanv[SKG_CHAN_SCATTERING_LEVEL1 + n] : none, dustsystem ? fi->Fchanv[SKG_CHAN_SCATTERING_LEVEL1 + n] : none,
                ("scatteringlevel" + std::to_string(n + 1)).c_str(), std::to_string(n + 1) + "-times scattered flux");
        for (size_t q = 0; q < f.size(); q++)
        {
            if (!f[q].empty())
                writeFITS(prefix + "_" + ins.name + "_" + fn[q] + ".fits", calibrateCube(ins, f[q], lg, units), d.Nxp, d.Nyp, lg.Nlambda(),
                          units.olength(d.fovxp / d.Nxp), units.olength(d.fovyp / d.Nyp), d.xpc, d.ypc, units.usurfacebrightness(), units.ulength(), stamp);
            if (!F[q].empty()) F[q] = calibrateFluxes(ins, F[q], lg, units);
        }
        writeSEDs(prefix + "_" + ins.name + "_sed.dat", lg, F, Fn, units);
        return;
    }
    if (d.kind != SKG_INSTR_SED)
        writeFITS(prefix + "_" + ins.name + "_total.fits", calibrateDataCube(ins, lg, units), d.Nxp, d.Nyp, lg.Nlambda(),
                  units.olength(d.fovxp / d.Nxp), units.olength(d.fovyp / d.Nyp), d.xpc, d.ypc, units.usurfacebrightness(), units.ulength(), stamp);
    if (d.kind != SKG_INSTR_FRAME)
        writeSED(prefix + "_" + ins.name + "_sed.dat", lg, calibrateSED(ins, lg, units), "total flux", units);
}

}   // namespace skirt
