// Segment sinks for the grid walkers (see geom.cuh).  Every sink reproduces
// DustGridPath::addSegment's running length `_s += ds` (DustGridPath.cpp:46-53); walkers call add()
// only for ds > 0.
#pragma once
#include "tables.h"

namespace skg
{

// KappaRho functor, DustSystem.cpp:465-491: sum over components h (in order) of kext[h][ell]*rho(m,h),
// with rho(-1,h) = 0 (DustSystem.cpp:918-921).
struct KappaRho
{
    const double* rho; const double* kextEll;   // kextEll = kext + ell, stride Nlambda
    int Ncomp, Nlambda;
    __device__ __forceinline__ double operator()(int m) const
    {
        double result = 0;
        for (int h = 0; h < Ncomp; h++)
            result += __ldg(kextEll + (size_t)h * Nlambda) * (m >= 0 ? __ldg(rho + (size_t)m * Ncomp + h) : 0.0);
        return result;
    }
};

// counts segments (first pass of the batched path())
struct CountSink
{
    int n = 0;
    __device__ __forceinline__ bool add(int, double) { n++; return true; }
};

// records Segment{m, ds, s, dtau, tau}: DustGridPath::addSegment + fillOpticalDepth (DustGridPath.hpp:117-129)
struct RecordSink
{
    int* m; double* ds; double* s; double* dtau; double* tau;   // already offset to this ray's first segment
    KappaRho kr; bool optical;
    double sacc = 0, tacc = 0;
    int n = 0;
    __device__ __forceinline__ bool add(int mm, double d)
    {
        sacc += d;
        double dt = 0;
        if (optical) { dt = kr(mm) * d; tacc += dt; }
        m[n] = mm; ds[n] = d; s[n] = sacc; dtau[n] = dt; tau[n] = tacc;
        n++;
        return true;
    }
};

// DustGridPath::opticalDepth(kapparho, distance), DustGridPath.hpp:97-108: the overshooting segment is
// counted in full, then the walk stops.
struct TauSink
{
    KappaRho kr; double distance;
    double sacc = 0, tau = 0;
    int n = 0;
    __device__ __forceinline__ bool add(int mm, double d)
    {
        sacc += d; n++;
        tau += kr(mm) * d;
        return !(sacc > distance);
    }
};

}   // namespace skg
