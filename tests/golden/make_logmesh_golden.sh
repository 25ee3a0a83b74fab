#!/bin/bash
# regenerates tests/golden/logmesh.txt from the reference's own NR::zerologgrid (Fundamentals/NR.hpp:283-289, what
# LogMesh::mesh calls, LogMesh.cpp:47-53); run in the container that has /root/reference
set -e
cd "$(dirname "$0")"
/usr/bin/g++ -std=c++11 -O3 -ffp-contract=off -w -I${REF:-/root/reference}/Fundamentals make_logmesh_golden.cpp -o /tmp/make_logmesh_golden
/tmp/make_logmesh_golden > logmesh.txt
wc -l logmesh.txt
