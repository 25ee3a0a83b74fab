// prints NR::zerologgrid (the reference's Fundamentals/NR.hpp, LogMesh::mesh) as hex floats: n tc then n+1 borders
#include <cstdio>
#include "NR.hpp"
int main()
{
    const int ns[] = {2, 3, 10, 40, 157};
    const double tcs[] = {1e-3, 0.01, 0.25, 0.9};
    for (int n : ns) for (double tc : tcs)
    {
        Array tv; NR::zerologgrid(tv, tc, 1.0, n);
        std::printf("%d %a", n, tc);
        for (size_t i = 0; i < tv.size(); i++) std::printf(" %a", tv[i]);
        std::printf("\n");
    }
    return 0;
}
