// Photon-packet life cycle on the device: the stellar emission phase of
// MonteCarloSimulation::dostellaremissionchunk (MonteCarloSimulation.cpp:265-301) with
//   launch                StellarSystem::launch (StellarSystem.cpp:116-158) + geometry samplers
//   peel-off emission     MonteCarloSimulation.cpp:305-315
//   escape + absorption   :438-515 (DustSystem::absorb -> atomicAdd on Labs[m,ell], replaces LockFree::add)
//   forced propagation    :519-537 (+ DustGridPath::pathlength, DustGridPath.cpp:162-173)
//   peel-off scattering   :319-363 (HG phase function DustMix.cpp:665-668)
//   scattering            :541-549 (DustMix.cpp:607-614, Random::direction Random.cpp:188-222)
//   detection             FrameInstrument.cpp:32-47, SEDInstrument.cpp:32-42, SimpleInstrument.cpp:33-49
// The reference stores every path (DustGridPath) and then loops over it; here the walkers stream the
// segments straight into sinks, so no path is ever written to memory: pass 1 accumulates tau_path and
// the per-cell absorption, pass 2 re-walks up to the sampled interaction optical depth.
#include <cmath>
#include <dlfcn.h>
#include <vector>
#include "engine.h"
#include "geom.cuh"
#include "sinks.cuh"
#include "philox.cuh"

namespace skg
{

struct GridSetMC { CartGrid cart; TreeGrid tree; AMeshGrid amesh; VoroGrid voro; };

struct McDev
{
    Medium med;
    const SourceDev* sources; int Nsources;
    const double* L;        // [Nsources*Nlambda]
    const double* Ltot;     // [Nlambda]
    const double* Lcdf;     // [Nlambda*(Nsources+1)]
    double emissionBias;
    const InstrDev* instr; int Ninstr;
    double* labs;           // [Ncells*Nlambda] or null
    double Npp;             // packets per wavelength shot by this engine
    double Lscale;          // total packets per wavelength over all engines
    double minWeightReduction, minfs, xi;
    uint64_t seed, streamOffset;
    int ellBegin, ellEnd;
    unsigned long long NppInt;
};

// ---- samplers ------------------------------------------------------------------------------------------

// SpecialFunctions::LambertW1, SpecialFunctions.cpp:579-627 (branch W_-1 for -1/e <= z < 0)
static __device__ double lambertW1(double z)
{
    const double eps = 1.0e-12;
    const double em1 = 0.3678794411714423215955237701614608;
    if (z == 0.0) return -SKG_DBL_MAX;
    double q = z + em1;
    if (q < 0) q = 0;
    double r = -sqrt(q);
    double t8 = -8.401032217523977370984161688514 + r * (12.250753501314460424 + r * (-18.100697012472442755 + r * 27.029044799010561650));
    double t5 = 3.066858901050631912893148922704 + r * (-4.175335600258177138854984177460 + r * (5.858023729874774148815053846119 + r * t8));
    double t1 = 2.331643981597124203363536062168 + r * (-1.812187885639363490240191647568 + r * (1.936631114492359755363277457668
              + r * (-2.353551201881614516821543561516 + r * t5)));
    double w0 = -1.0 + r * t1;
    if (q < 3.0e-3) return w0;
    double w;
    if (z < -1e-6) w = w0;
    else { double l1 = log(-z); double l2 = log(-l1); w = l1 - l2 + l2 / l1; }
    for (int i = 0; i < 10; i++)
    {
        double e = exp(w);
        double t = w * e - z;
        double p = w + 1.0;
        t /= e * p - 0.5 * (p + 1.0) * t / p;
        w -= t;
        if (fabs(t) < eps * (1.0 + fabs(w))) return w;
    }
    return w;
}

// Direction(theta, phi), Direction.cpp:12-38
__device__ __forceinline__ void directionFromAngles(double theta, double phi, double& kx, double& ky, double& kz)
{
    const double eps = 1e-8;
    if (theta <= eps) { kx = 0; ky = 0; kz = 1; }
    else if (theta >= M_PI - eps) { kx = 0; ky = 0; kz = -1; }
    else { double st = sin(theta); kx = st * cos(phi); ky = st * sin(phi); kz = cos(theta); }
}

// Random::direction(), Random.cpp:179-184
__device__ __forceinline__ void randomDirection(Philox& rng, double& kx, double& ky, double& kz)
{
    double theta = acos(2.0 * rng.uniform() - 1.0);
    double phi = 2.0 * M_PI * rng.uniform();
    directionFromAngles(theta, phi, kx, ky, kz);
}

// Random::direction(bfk, costheta), Random.cpp:188-222
__device__ __forceinline__ void scatterDirection(Philox& rng, double costheta, double& kx, double& ky, double& kz)
{
    double phi = 2.0 * M_PI * rng.uniform();
    double cosphi = cos(phi), sinphi = sin(phi);
    double sintheta = sqrt(fabs((1.0 - costheta) * (1.0 + costheta)));
    double kxn, kyn, kzn;
    if (kz > 0.99999) { kxn = cosphi * sintheta; kyn = sinphi * sintheta; kzn = costheta; }
    else if (kz < -0.99999) { kxn = cosphi * sintheta; kyn = sinphi * sintheta; kzn = -costheta; }
    else
    {
        double root = sqrt((1.0 - kz) * (1.0 + kz));
        kxn = sintheta / root * (-kx * kz * cosphi + ky * sinphi) + kx * costheta;
        kyn = -sintheta / root * (ky * kz * cosphi + kx * sinphi) + ky * costheta;
        kzn = root * sintheta * cosphi + kz * costheta;
    }
    kx = kxn; ky = kyn; kz = kzn;
}

// Random::exponcutoff, Random.cpp:162-175
__device__ __forceinline__ double exponCutoff(Philox& rng, double xmax)
{
    if (xmax == 0.0) return 0.0;
    else if (xmax < 1e-10) return rng.uniform() * xmax;
    double x = -log(1.0 - rng.uniform() * (1.0 - exp(-xmax)));
    while (x > xmax) x = -log(1.0 - rng.uniform() * (1.0 - exp(-xmax)));
    return x;
}

// NR::interpolate_loglog, NR.hpp:321-345
__device__ __forceinline__ double interpLogLog(double x, double x1, double x2, double f1, double f2)
{
    x = log10(x); x1 = log10(x1); x2 = log10(x2);
    bool logf = f1 > 0 && f2 > 0;
    if (logf) { f1 = log10(f1); f2 = log10(f2); }
    double fx = f1 + ((x - x1) / (x2 - x1)) * (f2 - f1);
    if (logf) fx = pow(10.0, fx);
    return fx;
}

// Geometry::generatePosition for the supported geometries
static __device__ void generatePosition(const SourceDev& s, Philox& rng, double& x, double& y, double& z)
{
    if (s.geometry == SKG_GEOM_EXPDISK)
    {
        // SepAxGeometry::generatePosition (SepAxGeometry.cpp:21-30) + ExpDiskGeometry::randomR/randomz (:134-161)
        const double hR = s.p[0], hz = s.p[1], Rmax = s.p[2], zmax = s.p[3], Rmin = s.p[4];
        double R, zz;
        do
        {
            double X = rng.uniform();
            R = hR * (-1.0 - lambertW1((X - 1.0) / M_E));
        }
        while ((Rmax > 0.0 && R >= Rmax) || R <= Rmin);
        double phi = 2.0 * M_PI * rng.uniform();
        do
        {
            double X = rng.uniform();
            zz = (X <= 0.5) ? hz * log(2.0 * X) : -hz * log(2.0 * (1.0 - X));
        }
        while (zmax > 0.0 && fabs(zz) >= zmax);
        x = R * cos(phi); y = R * sin(phi); z = zz;     // Position(R,phi,z,CYLINDRICAL), Position.cpp:23-31
    }
    else
    {
        // SpheGeometry::generatePosition (SpheGeometry.cpp:36-44) with SersicGeometry::randomradius (:85-91),
        // SersicFunction::inversemass (SersicFunction.cpp:112-124); SpheroidalGeometryDecorator (:78-85)
        const double reff = s.p[0], q = s.p[1];
        double X = rng.uniform();
        int Ns = s.ntab; double sval;
        if (X <= s.Xv[0]) sval = s.rv[0];
        else if (X >= s.Xv[Ns - 1]) sval = s.rv[Ns - 1];
        else
        {
            int i = locateClip(s.Xv, X, Ns);
            sval = interpLogLog(X, s.Xv[i], s.Xv[i + 1], s.rv[i], s.rv[i + 1]);
        }
        double r = reff * sval;
        double kx, ky, kz; randomDirection(rng, kx, ky, kz);
        x = r * kx; y = r * ky; z = r * kz;             // Position(r,bfk), Position.cpp:51-54
        z = q * z;
    }
    if (s.spiral_arms > 0)
    {
        // SpiralStructureGeometryDecorator::generatePosition, SpiralStructureGeometryDecorator.cpp:177-192
        double R = sqrt(x * x + y * y);
        double c = s.spiral_c;
        double phi, t;
        do
        {
            phi = 2.0 * M_PI * rng.uniform();
            // perturbation(R,phi), :224-229
            double gamma = log(R / s.spiral_radius) / s.spiral_tanp + s.spiral_phase + 0.5 * M_PI / s.spiral_arms;
            double pert = (1.0 - s.spiral_weight) + s.spiral_weight * s.spiral_cn * pow(sin(0.5 * s.spiral_arms * (gamma - phi)), 2 * s.spiral_index);
            t = rng.uniform() * c / pert;
        }
        while (t > 1);
        x = R * cos(phi); y = R * sin(phi);
    }
}

// ---- sinks for the life cycle --------------------------------------------------------------------------

// pass 1: DustSystem::fillOpticalDepth + simulateescapeandabsorption streamed per segment
struct AbsorbSink
{
    KappaRho kr; const Medium* med; int ell;
    double L;               // packet luminosity at the start of the path
    double albedo;          // Ncomp==1: DustMix::albedo(ell)
    double* labs;           // Labs + ell (stride Nlambda) or null
    double tau = 0, Lsca = 0;
    int n = 0, nAbs = 0;
    __device__ __forceinline__ bool add(int m, double ds)
    {
        n++;
        if (m < 0) return true;             // rho(-1,h) = 0: dtau = 0, nothing absorbed
        nAbs++;
        int Ncomp = med->Ncomp;
        if (Ncomp == 1)
        {
            double dtau = kr(m) * ds;
            if (labs)
            {
                double Lintm = L * exp(-tau) * (-expm1(-dtau));
                double Labsm = (1.0 - albedo) * Lintm;
                atomicAdd(labs + (size_t)m * med->Nlambda, Labsm);
            }
            tau += dtau;
        }
        else
        {
            double ksca = 0.0, kext = 0.0, krr = 0.0;
            for (int h = 0; h < Ncomp; h++)
            {
                double rho = __ldg(med->rho + (size_t)m * Ncomp + h);
                ksca += rho * __ldg(med->ksca + (size_t)h * med->Nlambda + ell);
                double ke = __ldg(med->kext + (size_t)h * med->Nlambda + ell);
                kext += rho * ke;
                krr += ke * rho;
            }
            double alb = (kext > 0.0) ? ksca / kext : 0.0;
            double dtau = krr * ds;
            double Lintm = L * exp(-tau) * (-expm1(-dtau));
            Lsca += alb * Lintm;
            if (labs) atomicAdd(labs + (size_t)m * med->Nlambda, (1.0 - alb) * Lintm);
            tau += dtau;
        }
        return true;
    }
};

// pass 2: DustGridPath::pathlength(tau) evaluated on the fly (DustGridPath.cpp:162-173)
struct PropagateSink
{
    KappaRho kr; double target;
    double sPrev = 0, tauPrev = 0, result = 0;
    bool found = false;
    int n = 0;
    __device__ __forceinline__ bool add(int m, double ds)
    {
        n++;
        double sNew = sPrev + ds;
        double tauNew = tauPrev + kr(m) * ds;
        if (target < tauNew)
        {
            result = sPrev + ((target - tauPrev) / (tauNew - tauPrev)) * (sNew - sPrev);     // NR::interpolate_linlin
            found = true;
            return false;
        }
        sPrev = sNew; tauPrev = tauNew;
        return true;
    }
    __device__ __forceinline__ double s() const { return found ? result : sPrev; }
};

template<int KIND, class Sink>
__device__ __forceinline__ void walkMC(const GridSetMC& G, const CartGrid& cart, Counters* ctr,
                                       double x, double y, double z, double kx, double ky, double kz, Sink& sink)
{
    if (KIND == GRID_CART) walkCart(cart, x, y, z, kx, ky, kz, sink);
    else if (KIND == GRID_TREE) walkTree(G.tree, ctr, x, y, z, kx, ky, kz, sink);
    else if (KIND == GRID_AMESH) walkAMesh(G.amesh, ctr, x, y, z, kx, ky, kz, sink);
    else walkVoro(G.voro, ctr, x, y, z, kx, ky, kz, sink);
}

template<int KIND>
__device__ __forceinline__ int whichCellMC(const GridSetMC& G, const CartGrid& cart, double x, double y, double z)
{
    if (KIND == GRID_CART) return cartWhichCell(cart, x, y, z);
    else if (KIND == GRID_TREE) { int node = treeWhichNode(G.tree, x, y, z); return node >= 0 ? G.tree.cell[node] : -1; }
    else if (KIND == GRID_AMESH) { int node = ameshWhichNode(G.amesh, x, y, z); return node >= 0 ? G.amesh.cell[node] : -1; }
    else return voroCellIndex(G.voro, x, y, z);
}

// Instrument::detect for the peel-off packet (r, kobs, L): returns the number of segments walked
template<int KIND>
__device__ __forceinline__ int detect(const GridSetMC& G, const CartGrid& cart, Counters* ctr, const McDev& P, const InstrDev& I,
                                      int ell, double x, double y, double z, double L, unsigned long long& nDet)
{
    int l = -1;
    if (I.kind != SKG_INSTR_SED)
    {
        // SingleFrameInstrument::pixelondetector, SingleFrameInstrument.cpp:130-147
        double xpp = -I.sinphi * x + I.cosphi * y;
        double ypp = -I.cosphi * I.costheta * x - I.sinphi * I.costheta * y + I.sintheta * z;
        double xp = I.cospa * xpp - I.sinpa * ypp;
        double yp = I.sinpa * xpp + I.cospa * ypp;
        int i = (int)floor((xp - I.xpmin) / I.xpsiz);
        int j = (int)floor((yp - I.ypmin) / I.ypsiz);
        if (!(i < 0 || i >= I.Nxp || j < 0 || j >= I.Nyp)) l = i + I.Nxp * j;
        if (I.kind == SKG_INSTR_FRAME && l < 0) return 0;       // FrameInstrument.cpp:36: no path for off-frame packets
    }
    TauSink sink;
    sink.kr = KappaRho{P.med.rho, P.med.kext + ell, P.med.Ncomp, P.med.Nlambda};
    sink.distance = SKG_DBL_MAX;
    if (P.med.rho) walkMC<KIND>(G, cart, ctr, x, y, z, I.kobsx, I.kobsy, I.kobsz, sink);     // Instrument::opticalDepth: 0 without dust
    double Lextf = L * exp(-sink.tau);
    if (I.kind != SKG_INSTR_FRAME) { atomicAdd(I.sed + ell, Lextf); nDet++; }
    if (I.kind != SKG_INSTR_SED && l >= 0) { atomicAdd(I.frame + (size_t)l + (size_t)ell * I.Nxp * I.Nyp, Lextf); nDet++; }
    return sink.n;
}

template<int KIND>
__global__ void __launch_bounds__(128) stellarKernel(const __grid_constant__ GridSetMC G, const __grid_constant__ McDev P,
                                                     Counters* ctr, bool cartSmem)
{
    extern __shared__ double smem[];
    CartGrid cart = G.cart;
    if (KIND == GRID_CART && cartSmem)
    {
        int nx = cart.Nx + 1, ny = cart.Ny + 1, nz = cart.Nz + 1;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) smem[i] = cart.xv[i];
        for (int i = threadIdx.x; i < ny; i += blockDim.x) smem[nx + i] = cart.yv[i];
        for (int i = threadIdx.x; i < nz; i += blockDim.x) smem[nx + ny + i] = cart.zv[i];
        __syncthreads();
        cart.xv = smem; cart.yv = smem + nx; cart.zv = smem + nx + ny;
    }

    unsigned long long nSeg = 0, nPaths = 0, nScatt = 0, nPackets = 0, nAbs = 0, nDet = 0;
    const unsigned long long total = P.NppInt * (unsigned long long)(P.ellEnd - P.ellBegin);
    const int Ncomp = P.med.Ncomp, Nlambda = P.med.Nlambda;

    for (unsigned long long gidx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; gidx < total;
         gidx += (unsigned long long)gridDim.x * blockDim.x)
    {
        const int ell = P.ellBegin + (int)(gidx / P.NppInt);
        const unsigned long long ipkt = gidx % P.NppInt;
        // MonteCarloSimulation.cpp:267-271
        double L = __ldg(P.Ltot + ell) / P.Lscale;
        if (!(L > 0)) continue;
        const double Lthreshold = L / P.minWeightReduction;
        Philox rng; rng.init(P.seed, (P.streamOffset + ipkt) * (unsigned long long)Nlambda + ell);
        nPackets++;

        // ---- StellarSystem::launch, StellarSystem.cpp:116-158 ----
        int h = 0;
        if (P.Nsources > 1)
        {
            int N = P.Nsources;
            double X = rng.uniform();
            if (X < P.emissionBias) h = max(0, min(N - 1, (int)(N * X / P.emissionBias)));
            else h = locateClip(P.Lcdf + (size_t)ell * (N + 1), (X - P.emissionBias) / (1.0 - P.emissionBias), N + 1);
            double Lh = __ldg(P.L + (size_t)h * Nlambda + ell);
            if (Lh > 0)
            {
                double Lmean = __ldg(P.Ltot + ell) / N;
                double weight = 1.0 / (1.0 - P.emissionBias + P.emissionBias * Lmean / Lh);
                L = L * weight;
            }
            else L = 0;
        }
        if (!(L > 0)) continue;
        double x, y, z, kx, ky, kz;
        generatePosition(P.sources[h], rng, x, y, z);        // GeometricStellarComp::launch, GeometricStellarComp.cpp:75-81
        randomDirection(rng, kx, ky, kz);                    // Geometry::generateDirection, Geometry.cpp:33
        int nscatt = 0;

        // ---- peeloffemission, MonteCarloSimulation.cpp:305-315 (isotropic emitter: probabilityForDirection = 1) ----
        for (int q = 0; q < P.Ninstr; q++)
        {
            int ns = detect<KIND>(G, cart, ctr, P, P.instr[q], ell, x, y, z, L, nDet);
            nSeg += ns; nPaths++;
        }

        if (P.med.rho) while (true)
        {
            // ---- fillOpticalDepth + simulateescapeandabsorption, :286-288, :438-515 ----
            AbsorbSink ab;
            ab.kr = KappaRho{P.med.rho, P.med.kext + ell, Ncomp, Nlambda};
            ab.med = &P.med; ab.ell = ell; ab.L = L;
            ab.labs = P.labs ? P.labs + ell : nullptr;
            double kext0 = __ldg(P.med.kext + ell), ksca0 = __ldg(P.med.ksca + ell);
            ab.albedo = kext0 > 0 ? ksca0 / kext0 : 0.0;        // DustMix::albedo(ell) (DustMix.cpp:55-90)
            walkMC<KIND>(G, cart, ctr, x, y, z, kx, ky, kz, ab);
            nSeg += ab.n; nPaths++; if (ab.labs) nAbs += ab.nAbs;
            const double taupath = ab.tau;
            if (Ncomp == 1) L = L * ab.albedo * (-expm1(-taupath));
            else L = ab.Lsca;

            // ---- termination test, :289 ----
            if (L <= 0 || (L <= Lthreshold && nscatt >= P.minfs)) break;

            // ---- simulatepropagation, :519-537 ----
            if (taupath != 0.0)
            {
                double tau;
                if (P.xi == 0.0) tau = exponCutoff(rng, taupath);
                else
                {
                    double X = rng.uniform();
                    tau = (X < P.xi) ? rng.uniform() * taupath : exponCutoff(rng, taupath);
                    double p = -exp(-tau) / expm1(-taupath);
                    double q = (1.0 - P.xi) * p + P.xi / taupath;
                    L = L * (p / q);
                }
                double s = 0;
                if (tau > 0)
                {
                    PropagateSink pr;
                    pr.kr = KappaRho{P.med.rho, P.med.kext + ell, Ncomp, Nlambda};
                    pr.target = tau;
                    walkMC<KIND>(G, cart, ctr, x, y, z, kx, ky, kz, pr);
                    nSeg += pr.n; nPaths++;
                    s = pr.s();
                }
                x += s * kx; y += s * ky; z += s * kz;      // PhotonPackage::propagate, PhotonPackage.cpp:93-96
            }

            // ---- peeloffscattering, :319-363 ----
            double wv[8]; bool peel = true; int mcell = -2;
            if (Ncomp == 1) wv[0] = 1.0;
            else
            {
                mcell = whichCellMC<KIND>(G, cart, x, y, z);
                if (mcell == -1) peel = false;
                else
                {
                    double sum = 0;
                    for (int c = 0; c < Ncomp && c < 8; c++)
                    { wv[c] = __ldg(P.med.ksca + (size_t)c * Nlambda + ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c); sum += wv[c]; }
                    if (sum <= 0) peel = false;
                    else for (int c = 0; c < Ncomp && c < 8; c++) wv[c] /= sum;
                }
            }
            if (peel) for (int q = 0; q < P.Ninstr; q++)
            {
                const InstrDev& I = P.instr[q];
                double cosalpha = kx * I.kobsx + ky * I.kobsy + kz * I.kobsz;       // Direction::dot
                double w = 0;
                for (int c = 0; c < Ncomp && c < 8; c++)
                {
                    // DustMix::phaseFunctionValue (HG), DustMix.cpp:665-668
                    double g = __ldg(P.med.g + (size_t)c * Nlambda + ell);
                    double t = 1.0 + g * g - 2 * g * cosalpha;
                    w += wv[c] * ((1.0 - g) * (1.0 + g) / sqrt(t * t * t));
                }
                int ns = detect<KIND>(G, cart, ctr, P, I, ell, x, y, z, L * w, nDet);      // launchScatteringPeelOff, PhotonPackage.cpp:51-62
                nSeg += ns; nPaths++;
            }

            // ---- simulatescattering, :541-549 ----
            int hmix = 0;
            if (Ncomp > 1)
            {
                // DustSystem::randomMixForPosition, DustSystem.cpp:879-893
                if (mcell == -2) mcell = whichCellMC<KIND>(G, cart, x, y, z);
                if (mcell >= 0)
                {
                    double Xv[9]; Xv[0] = 0;
                    for (int c = 0; c < Ncomp && c < 8; c++)
                        Xv[c + 1] = Xv[c] + __ldg(P.med.ksca + (size_t)c * Nlambda + ell) * __ldg(P.med.rho + (size_t)mcell * Ncomp + c);
                    double tot = Xv[Ncomp];
                    for (int c = 0; c <= Ncomp; c++) Xv[c] /= tot;
                    hmix = locateClip(Xv, rng.uniform(), Ncomp + 1);
                }
            }
            {
                // DustMix::scatteringDirectionAndPolarization (HG branch), DustMix.cpp:607-614
                double g = __ldg(P.med.g + (size_t)hmix * Nlambda + ell);
                if (fabs(g) < 1e-6) randomDirection(rng, kx, ky, kz);
                else
                {
                    double f = ((1.0 - g) * (1.0 + g)) / (1.0 - g + 2.0 * g * rng.uniform());
                    double costheta = (1.0 + g * g - f * f) / (2.0 * g);
                    scatterDirection(rng, costheta, kx, ky, kz);
                }
            }
            nscatt++; nScatt++;
        }
    }

    // statistics: warp-aggregated, one atomic per warp and counter
    for (int o = 16; o > 0; o >>= 1)
    {
        nSeg += __shfl_down_sync(0xffffffffu, nSeg, o); nPaths += __shfl_down_sync(0xffffffffu, nPaths, o);
        nScatt += __shfl_down_sync(0xffffffffu, nScatt, o); nPackets += __shfl_down_sync(0xffffffffu, nPackets, o);
        nAbs += __shfl_down_sync(0xffffffffu, nAbs, o); nDet += __shfl_down_sync(0xffffffffu, nDet, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicAdd(&ctr->segments, nSeg); atomicAdd(&ctr->paths, nPaths);
        atomicAdd(&ctr->scatterings, nScatt); atomicAdd(&ctr->packets, nPackets);
        atomicAdd(&ctr->absorbSegments, nAbs); atomicAdd(&ctr->detections, nDet);
    }
}

// ---- host side -------------------------------------------------------------------------------------------

void mcSetSources(Engine& e, int Ncomp, const skg_source* comps, int Nlambda, const double* L, double emissionBias)
{
    if (Ncomp < 1 || !comps || Nlambda < 1 || !L) throw Error("skg_sources: bad arguments");
    if (e.med.Nlambda && Nlambda != e.med.Nlambda) throw Error("sources and medium disagree on the number of wavelengths");
    for (DevBuf* b : e.sourceBufs) delete b;
    e.sourceBufs.clear(); e.sources.clear();
    for (int h = 0; h < Ncomp; h++)
    {
        const skg_source& c = comps[h];
        SourceDev s{};
        s.geometry = c.geometry;
        for (int j = 0; j < 8; j++) s.p[j] = c.p[j];
        if (c.geometry == SKG_GEOM_EXPDISK)
        {
            if (!(c.p[0] > 0) || !(c.p[1] > 0)) throw Error("The radial scale length hR and axial scale height hz should be positive");  // ExpDiskGeometry.cpp:27-28
        }
        else if (c.geometry == SKG_GEOM_SERSIC)
        {
            if (!(c.p[0] > 0)) throw Error("the effective radius should be positive");
            if (c.ntab < 2 || !c.rv || !c.Xv) throw Error("Sersic geometry needs the tabulated inverse mass function");
            if (c.p[1] == 0) s.p[1] = 1.0;
            DevBuf* a = new DevBuf(); DevBuf* b = new DevBuf(); e.sourceBufs.push_back(a); e.sourceBufs.push_back(b);
            a->upload(c.rv, sizeof(double) * c.ntab, e.stream); b->upload(c.Xv, sizeof(double) * c.ntab, e.stream);
            s.ntab = c.ntab; s.rv = a->as<double>(); s.Xv = b->as<double>();
        }
        else throw Error("unsupported source geometry (no CPU fallback): " + std::to_string(c.geometry));
        s.spiral_arms = c.spiral_arms; s.spiral_index = c.spiral_index; s.spiral_pitch = c.spiral_pitch;
        s.spiral_radius = c.spiral_radius; s.spiral_phase = c.spiral_phase; s.spiral_weight = c.spiral_weight;
        if (c.spiral_arms > 0)
        {
            // SpiralStructureGeometryDecorator::setupSelfBefore, SpiralStructureGeometryDecorator.cpp:24-40
            if (c.spiral_pitch <= 0 || c.spiral_pitch >= M_PI / 2.) throw Error("The pitch angle should be between 0 and 90 degrees");
            if (c.spiral_radius <= 0) throw Error("The radius zero-point should be positive");
            if (c.spiral_weight <= 0 || c.spiral_weight > 1.) throw Error("The weight of the spiral perturbation should be between 0 and 1");
            if (c.spiral_index < 0 || c.spiral_index > 10) throw Error("The arm-interarm size ratio index should be between 0 and 10");
            s.spiral_tanp = std::tan(c.spiral_pitch);
            s.spiral_cn = std::sqrt(M_PI) * std::tgamma(c.spiral_index + 1.0) / std::tgamma(c.spiral_index + 0.5);
            s.spiral_c = 1.0 + (s.spiral_cn - 1.0) * c.spiral_weight;
        }
        e.sources.push_back(s);
    }
    e.sourcesDev.upload(e.sources.data(), sizeof(SourceDev) * Ncomp, e.stream);
    // StellarSystem::setupSelfAfter, StellarSystem.cpp:35-52: total luminosities and per-wavelength CDFs
    e.lumHost.assign(L, L + (size_t)Ncomp * Nlambda);
    e.lumTotHost.assign(Nlambda, 0.0);
    std::vector<double> cdf((size_t)Nlambda * (Ncomp + 1), 0.0);
    for (int ell = 0; ell < Nlambda; ell++)
    {
        for (int h = 0; h < Ncomp; h++) e.lumTotHost[ell] += L[(size_t)h * Nlambda + ell];
        double* X = cdf.data() + (size_t)ell * (Ncomp + 1);
        for (int h = 0; h < Ncomp; h++) X[h + 1] = X[h] + L[(size_t)h * Nlambda + ell];     // NR::cdf, NR.hpp:404-409
        double norm = X[Ncomp];
        for (int h = 0; h <= Ncomp; h++) X[h] /= norm;
    }
    e.lumDev.upload(e.lumHost.data(), sizeof(double) * e.lumHost.size(), e.stream);
    e.lumTotDev.upload(e.lumTotHost.data(), sizeof(double) * Nlambda, e.stream);
    e.lumCdfDev.upload(cdf.data(), sizeof(double) * cdf.size(), e.stream);
    e.Nsources = Ncomp; e.NlambdaSrc = Nlambda; e.emissionBias = emissionBias;
    e.sync();
}

void mcSetInstruments(Engine& e, int n, const skg_instrument* instr)
{
    if (n < 0 || (n > 0 && !instr)) throw Error("skg_instruments: bad arguments");
    if (!e.med.Nlambda) throw Error("skg_instruments needs skg_medium first (number of wavelengths)");
    for (DevBuf* b : e.instrBufs) delete b;
    e.instrBufs.clear(); e.instr.clear();
    for (int i = 0; i < n; i++)
    {
        const skg_instrument& s = instr[i];
        InstrDev d{};
        d.kind = s.kind;
        if (s.kind < SKG_INSTR_FRAME || s.kind > SKG_INSTR_SIMPLE) throw Error("unsupported instrument kind");
        if (s.distance <= 0) throw Error("Distance was not set");                    // DistantInstrument.cpp:32
        // DistantInstrument::setupSelfBefore, DistantInstrument.cpp:27-50
        d.costheta = std::cos(s.inclination); d.sintheta = std::sin(s.inclination);
        d.cosphi = std::cos(s.azimuth); d.sinphi = std::sin(s.azimuth);
        d.cospa = std::cos(s.positionAngle); d.sinpa = std::sin(s.positionAngle);
        {
            const double eps = 1e-8; double theta = s.inclination, phi = s.azimuth;      // Direction(theta,phi), Direction.cpp:12-38
            if (theta < -eps || theta > M_PI + eps) throw Error("Theta should be between 0 and pi.");
            if (theta <= eps) { d.kobsx = 0; d.kobsy = 0; d.kobsz = 1; }
            else if (theta >= M_PI - eps) { d.kobsx = 0; d.kobsy = 0; d.kobsz = -1; }
            else { double st = std::sin(theta); d.kobsx = st * std::cos(phi); d.kobsy = st * std::sin(phi); d.kobsz = std::cos(theta); }
        }
        if (s.kind != SKG_INSTR_SED)
        {
            // SingleFrameInstrument::setupSelfBefore, SingleFrameInstrument.cpp:26-42
            if (s.Nxp <= 0 || s.Nyp <= 0) throw Error("Number of pixels was not set");
            if (s.fovxp <= 0 || s.fovyp <= 0) throw Error("Field of view was not set");
            d.Nxp = s.Nxp; d.Nyp = s.Nyp;
            d.xpmin = s.xpc - 0.5 * s.fovxp; d.xpsiz = s.fovxp / s.Nxp;
            d.ypmin = s.ypc - 0.5 * s.fovyp; d.ypsiz = s.fovyp / s.Nyp;
            DevBuf* f = new DevBuf(); e.instrBufs.push_back(f);
            size_t bytes = sizeof(double) * (size_t)s.Nxp * s.Nyp * e.med.Nlambda;
            f->ensure(bytes); SKG_CUDA(cudaMemsetAsync(f->p, 0, bytes, e.stream));
            d.frame = f->as<double>();
        }
        if (s.kind != SKG_INSTR_FRAME)
        {
            DevBuf* f = new DevBuf(); e.instrBufs.push_back(f);
            f->ensure(sizeof(double) * e.med.Nlambda); SKG_CUDA(cudaMemsetAsync(f->p, 0, sizeof(double) * e.med.Nlambda, e.stream));
            d.sed = f->as<double>();
        }
        e.instr.push_back(d);
    }
    e.instrDev.upload(e.instr.data(), sizeof(InstrDev) * std::max(n, 1), e.stream);
    e.sync();
}

void mcResetResults(Engine& e)
{
    for (const InstrDev& d : e.instr)
    {
        if (d.frame) SKG_CUDA(cudaMemsetAsync(d.frame, 0, sizeof(double) * (size_t)d.Nxp * d.Nyp * e.med.Nlambda, e.stream));
        if (d.sed) SKG_CUDA(cudaMemsetAsync(d.sed, 0, sizeof(double) * e.med.Nlambda, e.stream));
    }
    if (e.labs.p && e.labsCount) SKG_CUDA(cudaMemsetAsync(e.labs.p, 0, sizeof(double) * e.labsCount, e.stream));
    e.sync();
}

void mcRunStellar(Engine& e, const skg_mc_params& p, skg_mc_stats* stats)
{
    if (e.gridKind == GRID_NONE && e.med.rho) throw Error("no dust grid has been set");
    if (!e.Nsources) throw Error("no sources have been set");
    if (e.med.Nlambda && e.NlambdaSrc != e.med.Nlambda) throw Error("sources and medium disagree on the number of wavelengths");
    int Nlambda = e.NlambdaSrc;
    if (p.ellBegin < 0 || p.ellEnd > Nlambda || p.ellBegin > p.ellEnd) throw Error("wavelength range out of bounds");
    if (p.scattBias < 0 || p.scattBias > 1) throw Error("scattBias should be between 0 and 1");
    if (e.med.Ncomp > 8) throw Error("at most 8 dust components are supported");
    if (p.storeAbsorption)
    {
        int64_t count = (int64_t)e.Ncells * Nlambda;
        if (e.labsCount != count)
        {
            e.labs.ensure(sizeof(double) * count); e.labsCount = count;
            SKG_CUDA(cudaMemsetAsync(e.labs.p, 0, sizeof(double) * count, e.stream));
        }
    }
    McDev P{};
    P.med = e.med; if (!e.med.rho) { P.med.Nlambda = Nlambda; P.med.Ncomp = 0; }
    P.sources = e.sourcesDev.as<SourceDev>(); P.Nsources = e.Nsources;
    P.L = e.lumDev.as<double>(); P.Ltot = e.lumTotDev.as<double>(); P.Lcdf = e.lumCdfDev.as<double>();
    P.emissionBias = e.emissionBias;
    P.instr = e.instrDev.as<InstrDev>(); P.Ninstr = (int)e.instr.size();
    P.labs = p.storeAbsorption ? e.labs.as<double>() : nullptr;
    P.NppInt = (unsigned long long)std::ceil(p.packages);
    P.Npp = (double)P.NppInt; P.Lscale = p.luminosityScale > 0 ? p.luminosityScale : P.Npp;
    P.minWeightReduction = p.minWeightReduction; P.minfs = p.minScattEvents; P.xi = p.scattBias;
    P.seed = p.seed; P.streamOffset = p.streamOffset; P.ellBegin = p.ellBegin; P.ellEnd = p.ellEnd;

    GridSetMC G; G.cart = e.cart; G.tree = e.tree; G.amesh = e.amesh; G.voro = e.voro;
    unsigned long long total = P.NppInt * (unsigned long long)(p.ellEnd - p.ellBegin);
    Counters before = e.readCounters();
    cudaEvent_t ev0, ev1; SKG_CUDA(cudaEventCreate(&ev0)); SKG_CUDA(cudaEventCreate(&ev1));
    SKG_CUDA(cudaEventRecord(ev0, e.stream));
    if (total > 0)
    {
        long long want = (long long)((total + 127) / 128);
        int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)e.smCount * 16));
        size_t smem = 0; bool cartSmem = false;
        if (e.gridKind == GRID_CART)
        {
            size_t need = sizeof(double) * (size_t)(e.cart.Nx + e.cart.Ny + e.cart.Nz + 3);
            if (need <= 40 * 1024) { smem = need; cartSmem = true; }
        }
        switch (e.gridKind)
        {
        case GRID_CART: stellarKernel<GRID_CART><<<blocks, 128, smem, e.stream>>>(G, P, e.ctr(), cartSmem); break;
        case GRID_TREE: stellarKernel<GRID_TREE><<<blocks, 128, 0, e.stream>>>(G, P, e.ctr(), false); break;
        case GRID_AMESH: stellarKernel<GRID_AMESH><<<blocks, 128, 0, e.stream>>>(G, P, e.ctr(), false); break;
        case GRID_VORO: stellarKernel<GRID_VORO><<<blocks, 128, 0, e.stream>>>(G, P, e.ctr(), false); break;
        default: throw Error("no dust grid has been set");
        }
        e.launches++; SKG_CUDA(cudaGetLastError());
    }
    SKG_CUDA(cudaEventRecord(ev1, e.stream));
    e.sync();
    float ms = 0; SKG_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    Counters after = e.readCounters();
    if (stats)
    {
        stats->packets = after.packets - before.packets; stats->pathSegments = after.segments - before.segments;
        stats->paths = after.paths - before.paths; stats->scatterings = after.scatterings - before.scatterings;
        stats->kernel_ms = ms;
        stats->absorbSegments = after.absorbSegments - before.absorbSegments; stats->detections = after.detections - before.detections;
    }
}

// ---- NCCL (loaded at run time: libnccl.so.2 is already in the process when torch.distributed is) ---------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int (*fnGetUniqueId)(ncclUniqueId*);
typedef int (*fnCommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*fnAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fnGroup)(void);
typedef const char* (*fnErr)(int);
static struct { void* lib = nullptr; fnGetUniqueId getUniqueId; fnCommInitRank commInitRank; fnAllReduce allReduce; fnGroup groupStart, groupEnd; fnErr errString; } nccl;

static void loadNccl()
{
    if (nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (nccl.lib) break; }
    if (!nccl.lib) throw Error(std::string("cannot load NCCL: ") + dlerror());
    nccl.getUniqueId = (fnGetUniqueId)dlsym(nccl.lib, "ncclGetUniqueId");
    nccl.commInitRank = (fnCommInitRank)dlsym(nccl.lib, "ncclCommInitRank");
    nccl.allReduce = (fnAllReduce)dlsym(nccl.lib, "ncclAllReduce");
    nccl.groupStart = (fnGroup)dlsym(nccl.lib, "ncclGroupStart");
    nccl.groupEnd = (fnGroup)dlsym(nccl.lib, "ncclGroupEnd");
    nccl.errString = (fnErr)dlsym(nccl.lib, "ncclGetErrorString");
    if (!nccl.getUniqueId || !nccl.commInitRank || !nccl.allReduce || !nccl.groupStart || !nccl.groupEnd) throw Error("NCCL symbols missing");
}
#define SKG_NCCL(call) do { int rc__ = (call); if (rc__ != 0) throw skg::Error(std::string(#call) + ": " + (nccl.errString ? nccl.errString(rc__) : "NCCL error")); } while (0)

}   // namespace skg

using namespace skg;
extern "C"
{
int skg_comm_unique_id(void* out)
{
    try { loadNccl(); ncclUniqueId id; SKG_NCCL(nccl.getUniqueId(&id)); memcpy(out, &id, 128); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
int skg_comm_init(skg_engine* eh, int rank, int nranks, const void* uid)
{
    try
    {
        Engine& e = *reinterpret_cast<Engine*>(eh);
        SKG_CUDA(cudaSetDevice(e.device));
        loadNccl();
        ncclUniqueId id; memcpy(&id, uid, 128);
        ncclComm_t comm; SKG_NCCL(nccl.commInitRank(&comm, nranks, id, rank));
        e.nccl = comm; e.rank = rank; e.nranks = nranks;
        return 0;
    }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
// replaces Instrument::sumResults (Instrument.cpp:57-65) and PanDustSystem::sumResults (PanDustSystem.cpp:394-404):
// one grouped in-place ncclAllReduce(double, sum) over Labs and every detector array
int skg_allreduce_results(skg_engine* eh)
{
    try
    {
        Engine& e = *reinterpret_cast<Engine*>(eh);
        SKG_CUDA(cudaSetDevice(e.device));
        if (!e.nccl || e.nranks <= 1) return 0;
        const int ncclDouble = 8, ncclSum = 0;       // nccl.h: ncclFloat64 = 8, ncclSum = 0
        ncclComm_t comm = (ncclComm_t)e.nccl;
        SKG_NCCL(nccl.groupStart());
        if (e.labs.p && e.labsCount) SKG_NCCL(nccl.allReduce(e.labs.p, e.labs.p, (size_t)e.labsCount, ncclDouble, ncclSum, comm, e.stream));
        for (const InstrDev& d : e.instr)
        {
            if (d.frame) SKG_NCCL(nccl.allReduce(d.frame, d.frame, (size_t)d.Nxp * d.Nyp * e.med.Nlambda, ncclDouble, ncclSum, comm, e.stream));
            if (d.sed) SKG_NCCL(nccl.allReduce(d.sed, d.sed, (size_t)e.med.Nlambda, ncclDouble, ncclSum, comm, e.stream));
        }
        SKG_NCCL(nccl.groupEnd());
        e.sync();
        return 0;
    }
    catch (std::exception& ex) { setLastError(ex.what()); return 1; }
}
}
