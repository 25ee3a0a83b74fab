// Link-level stand-ins for the reference's file-output / console services (TextOutFile.cpp,
// DustGridPlotFile.cpp, Image.cpp, FilePaths.cpp, Console.cpp, MemoryStatistics.cpp).  None of them
// is on the photon-packet path; they are replaced so that oracle/_ref needs neither Qt's file
// classes nor cfitsio.  Declarations come from the reference's own headers.
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include "TextOutFile.hpp"
#include "DustGridPlotFile.hpp"
#include "Image.hpp"
#include "FilePaths.hpp"
#include "Console.hpp"
#include "MemoryStatistics.hpp"

// ---- TextOutFile: keep the numeric rows in memory, keyed by file name (SED files) ----
std::map<std::string, std::vector<std::vector<double>>>& skr_saved_rows()
{ static std::map<std::string, std::vector<std::vector<double>>> m; return m; }
static std::map<const TextOutFile*, std::string>& openNames() { static std::map<const TextOutFile*, std::string> m; return m; }
static std::mutex txtMutex;
TextOutFile::TextOutFile(const SimulationItem*, QString filename, QString, bool) : _log(0), _units(0), _ncolumns(0)
{ std::lock_guard<std::mutex> lock(txtMutex); openNames()[this] = filename.toStdString(); skr_saved_rows()[filename.toStdString()].clear(); }
TextOutFile::~TextOutFile() { std::lock_guard<std::mutex> lock(txtMutex); openNames().erase(this); }
void TextOutFile::writeLine(QString) {}
void TextOutFile::addColumn(QString, char, int) {}
void TextOutFile::writeRow(QList<double> values)
{
    std::lock_guard<std::mutex> lock(txtMutex);
    std::vector<double> row; for (int i = 0; i < values.size(); i++) row.push_back(values[i]);
    skr_saved_rows()[openNames()[this]].push_back(row);
}

// ---- DustGridPlotFile ----
DustGridPlotFile::DustGridPlotFile(const SimulationItem* item, QString filename) : TextOutFile(item, filename, "") {}
void DustGridPlotFile::writeLine(double, double, double, double) {}
void DustGridPlotFile::writeRectangle(double, double, double, double) {}
void DustGridPlotFile::writeCircle(double) {}
void DustGridPlotFile::writeLine(double, double, double, double, double, double) {}
void DustGridPlotFile::writeCube(double, double, double, double, double, double) {}
void DustGridPlotFile::writePolyhedron(const std::vector<double>&, const std::vector<int>&) {}

// ---- Image: keep the saved (calibrated) data cubes in memory, keyed by file name ----
std::map<std::string, std::vector<double>>& skr_saved_images()
{ static std::map<std::string, std::vector<double>> m; return m; }
static std::mutex imgMutex;
Image::Image() : _units(0), _xsize(0), _ysize(0), _nframes(0), _incx(0), _incy(0), _xc(0), _yc(0) {}
Image::Image(const SimulationItem*, int xsize, int ysize, int nframes, double xres, double yres, QString quantity, QString xyqty)
    : _units(0), _xsize(xsize), _ysize(ysize), _nframes(nframes), _incx(xres), _incy(yres), _xc(0), _yc(0), _dataunits(quantity), _xyunits(xyqty) {}
Image::Image(const SimulationItem*, int xsize, int ysize, int nframes, double xres, double yres, double xc, double yc, QString quantity, QString xyqty)
    : _units(0), _xsize(xsize), _ysize(ysize), _nframes(nframes), _incx(xres), _incy(yres), _xc(xc), _yc(yc), _dataunits(quantity), _xyunits(xyqty) {}
void Image::saveto(const SimulationItem*, const Array& data, QString filename, QString)
{
    std::lock_guard<std::mutex> lock(imgMutex);
    std::vector<double>& v = skr_saved_images()[filename.toStdString()];
    v.assign(&data[0], &data[0] + data.size());
}

// ---- FilePaths: resources resolve inside the reference's own dat/ tree ----
FilePaths::FilePaths() {}
void FilePaths::setInputPath(QString value) { _inputPath = value; }
QString FilePaths::inputPath() const { return _inputPath; }
void FilePaths::setOutputPath(QString value) { _outputPath = value; }
QString FilePaths::outputPath() const { return _outputPath; }
void FilePaths::setOutputPrefix(QString value) { _outputPrefix = value; }
QString FilePaths::outputPrefix() const { return _outputPrefix; }
QString FilePaths::input(QString name) const { return _inputPath + name; }
QString FilePaths::output(QString name) const { return _outputPath + _outputPrefix + "_" + name; }
QString FilePaths::application(QString name) { return name; }
QString FilePaths::resource(QString name)
{
    const char* root = getenv("SKIRT_REFERENCE_ROOT");
    return QString(root ? root : "/root/reference") + "/dat/" + name;
}
QString FilePaths::externalResource(QString name) { return resource(name); }

// ---- Console: silent log that counts warnings ("stuck packet" guard etc.) ----
long skr_warning_count = 0;
int skr_verbose = 0;
Console::Console() {}
void Console::output(QString message, Level level)
{
    if (level == Warning) __sync_fetch_and_add(&skr_warning_count, 1);
    if (skr_verbose || level == Error) fprintf(stderr, "[skirt-ref] %s\n", message.toStdString().c_str());
}
QString Console::promptForInput(QString) { return QString(); }

// ---- MemoryStatistics ----
size_t MemoryStatistics::availableMemory() { return 0; }
size_t MemoryStatistics::peakMemoryUsage() { return 0; }
size_t MemoryStatistics::currentMemoryUsage() { return 0; }
QString MemoryStatistics::reportAvailable(bool) { return QString(); }
QString MemoryStatistics::reportPeak(bool) { return QString(); }
QString MemoryStatistics::reportCurrent(bool) { return QString(); }
