// Output side of the instruments in the C++ host layer: unit systems, calibration of the detector arrays and the two
// wire formats of the reference (FITS data cubes with 32-bit float pixels, SED text files).
//
//   Units / SIUnits / StellarUnits / ExtragalacticUnits     Units.cpp:30-215,495-506,765-1040
//   SingleFrameInstrument::calibrateAndWriteDataCubes       SingleFrameInstrument.cpp:151-226
//   DistantInstrument::calibrateAndWriteSEDs                DistantInstrument.cpp:131-183
//   FITSInOut::write                                        FITSInOut.cpp:32-90
//   TextOutFile                                             TextOutFile.cpp:45-85
#pragma once
#include <string>
#include <vector>
#include "SimulationItems.hpp"

namespace skirt
{

class UnitSystem
{
public:
    enum FluxOutputStyle { Neutral, Wavelength, Frequency };
    virtual ~UnitSystem() {}
    void setFluxOutputStyle(FluxOutputStyle v) { _style = v; }
    FluxOutputStyle fluxOutputStyle() const { return _style; }
    std::string ulength() const { return _ulength; }
    std::string uwavelength() const { return _uwavelength; }
    std::string sfluxdensity() const { return _style == Wavelength ? "F_lambda" : _style == Frequency ? "F_nu" : "lambda*F_lambda"; }
    std::string ufluxdensity() const { return _style == Wavelength ? _uwfd : _style == Frequency ? _uffd : _unfd; }
    std::string usurfacebrightness() const { return _style == Wavelength ? _uwsb : _style == Frequency ? _ufsb : _unsb; }
    double olength(double x) const { return x / _clength; }
    double owavelength(double lambda) const { return lambda / _cwavelength; }
    double ofluxdensity(double lambda, double Flambda) const;
    double osurfacebrightness(double lambda, double flambda) const;
protected:
    FluxOutputStyle _style = Neutral;
    std::string _ulength, _uwavelength, _unfd, _unsb, _uwfd, _uwsb, _uffd, _ufsb;
    double _clength = 1, _cwavelength = 1, _cnfd = 1, _cnsb = 1, _cwfd = 1, _cwsb = 1, _cffd = 1, _cfsb = 1;
};
class SIUnits : public UnitSystem { public: SIUnits(); };
class StellarUnits : public UnitSystem { public: StellarUnits(); };
class ExtragalacticUnits : public UnitSystem { public: ExtragalacticUnits(); };

// calibrated copies of the detector arrays of one instrument (the reference calibrates in place)
std::vector<double> calibrateDataCube(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units);
std::vector<double> calibrateSED(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units);

// FITS primary HDU, BITPIX -32, nx x ny x nz pixels (two axes when nz == 1); stamp = "" takes the current UTC time
void writeFITS(const std::string& path, const std::vector<double>& data, int nx, int ny, int nz, double incx, double incy,
               double xc, double yc, const std::string& dataUnits, const std::string& xyUnits, const std::string& stamp = "");
void writeSED(const std::string& path, const WavelengthGrid& lg, const std::vector<double>& F, const std::string& columnName, const UnitSystem& units);
// several flux columns (DistantInstrument::calibrateAndWriteSEDs, DistantInstrument.cpp:131-183); an empty column is written as zeros
void writeSEDs(const std::string& path, const WavelengthGrid& lg, const std::vector<std::vector<double>>& Fs,
               const std::vector<std::string>& columnNames, const UnitSystem& units);

// Instrument::write() of FrameInstrument / SEDInstrument / SimpleInstrument: <prefix>_<name>_total.fits, <prefix>_<name>_sed.dat;
// of FullInstrument (FullInstrument.cpp:176-236): <prefix>_<name>_{total,direct,scattered,dust,dustscattered,transparent,
// scatteringlevelN}.fits for the non-empty channels and one SED file with a column per channel
void writeInstrument(const Instrument& ins, const WavelengthGrid& lg, const UnitSystem& units, const std::string& prefix, const std::string& stamp = "",
                     bool dustsystem = true, bool dustemission = false);

}   // namespace skirt
