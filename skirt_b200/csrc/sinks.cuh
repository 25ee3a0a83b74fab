// The opacity functor shared by every job that turns path segments into optical depth.
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

}   // namespace skg
