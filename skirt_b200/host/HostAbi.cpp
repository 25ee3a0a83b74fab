// C ABI of the host-side set-up library (include/skirthost.h) over GridBuilders.
#include <algorithm>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include "../../include/skirthost.h"
#include "GridBuilders.hpp"

static thread_local std::string g_err;
template<class F> static int guarded(F f)
{
    try { f(); return 0; }
    catch (std::exception& ex) { g_err = ex.what(); }
    catch (...) { g_err = "unknown error"; }
    return 1;
}
template<class T> static void put(T* dst, const std::vector<T>& src) { if (dst && !src.empty()) std::copy(src.begin(), src.end(), dst); }

struct skh_tree { skirt::TreeBuilder b; skh_tree(int k, const double* e, int lo, int hi) : b(k, e, lo, hi) {} };
struct skh_amesh { skirt::AMeshTables t; };
struct skh_voronoi { skirt::VoronoiTables t; };

extern "C"
{
const char* skh_last_error(void) { return g_err.c_str(); }

int skh_tree_create(int kind, const double* extent6, int minLevel, int maxLevel, skh_tree** out)
{ return guarded([&]{ if (!extent6 || !out) throw std::runtime_error("null argument"); *out = new skh_tree(kind, extent6, minLevel, maxLevel); }); }
void skh_tree_destroy(skh_tree* t) { delete t; }
int skh_tree_frontier(skh_tree* t, int* level, int64_t* size, int* needsDecision)
{ return guarded([&]{ if (!t) throw std::runtime_error("null tree"); if (level) *level = t->b.frontierLevel(); if (size) *size = (int64_t)t->b.frontierSize();
                      if (needsDecision) *needsDecision = t->b.frontierNeedsDecision() ? 1 : 0; }); }
int skh_tree_frontier_boxes(skh_tree* t, double* box6)
{ return guarded([&]{ if (!t || !box6) throw std::runtime_error("null argument"); t->b.frontierBoxes(box6); }); }
int skh_tree_subdivide(skh_tree* t, const unsigned char* flags)
{ return guarded([&]{ if (!t) throw std::runtime_error("null tree"); t->b.subdivide(flags); }); }
int skh_tree_subdivide_at(skh_tree* t, const unsigned char* flags, const double* barycenters)
{ return guarded([&]{ if (!t) throw std::runtime_error("null tree"); t->b.subdivide(flags, barycenters); }); }
int skh_tree_finish(skh_tree* t, int search, int* Nnodes, int* Ncells, int64_t* Nneighbours)
{ return guarded([&]{ if (!t) throw std::runtime_error("null tree"); t->b.finish(search); const skirt::TreeTables& T = t->b.tables();
                      if (Nnodes) *Nnodes = T.Nnodes;
                      if (Ncells) *Ncells = T.Ncells;
                      if (Nneighbours) *Nneighbours = (int64_t)T.nbrIds.size(); }); }
int skh_ptree_build(int kind, const double* extent6, const double* particles, int64_t n, int extraLevels, skh_tree** out, int* Nnodes, int* Ncells)
{ return guarded([&]{ if (!extent6 || !out || n < 0 || extraLevels < 0) throw std::runtime_error("bad arguments"); skh_tree* t = new skh_tree(kind, extent6, 0, 2);      // (the level limits of TreeDustGrid play no role here)
                      try { t->b.addParticles(particles, (size_t)n, extraLevels); } catch (...) { delete t; throw; }
                      *out = t; if (Nnodes) *Nnodes = t->b.tables().Nnodes; if (Ncells) *Ncells = t->b.tables().Ncells; }); }
int skh_tree_tables(skh_tree* t, double* box, int* child0, int* parent, int* cell, int* dir, int* level, int* nbrStart, int* nbrIds)
{ return guarded([&]{ if (!t) throw std::runtime_error("null tree"); const skirt::TreeTables& T = t->b.tables();
                      if (T.Nnodes == 0) throw std::runtime_error("skh_tree_finish has not been called");
                      put(box, T.box); put(child0, T.child0); put(parent, T.parent); put(cell, T.cell); put(dir, T.dir); put(level, T.level);
                      put(nbrStart, T.nbrStart); put(nbrIds, T.nbrIds); }); }

int skh_amesh_build(const double* extent6, const int* nxyz, int64_t n, skh_amesh** out, int* Nnodes, int* Ncells)
{ return guarded([&]{ if (!extent6 || !nxyz || !out || n < 1) throw std::runtime_error("bad arguments"); skh_amesh* a = new skh_amesh();
                      try { a->t = skirt::buildAdaptiveMesh(extent6, nxyz, (size_t)n); } catch (...) { delete a; throw; }
                      *out = a; if (Nnodes) *Nnodes = a->t.Nnodes; if (Ncells) *Ncells = a->t.Ncells; }); }
int skh_amesh_tables(skh_amesh* a, double* box, int* nxyz, int* child0, int* cell, int* wallNbr, double* volume, int* fileIndex)
{ return guarded([&]{ if (!a) throw std::runtime_error("null mesh"); const skirt::AMeshTables& T = a->t;
                      put(box, T.box); put(nxyz, T.nxyz); put(child0, T.child0); put(cell, T.cell); put(wallNbr, T.wallNbr); put(volume, T.volume); put(fileIndex, T.fileIndex); }); }
void skh_amesh_destroy(skh_amesh* a) { delete a; }

int skh_voronoi_available(void) { return skirt::voronoiAvailable() ? 1 : 0; }
int skh_voronoi_build(const double* extent6, const double* particles, int64_t n, skh_voronoi** out, int64_t* sizes5)
{ return guarded([&]{ if (!extent6 || !particles || !out || n < 1) throw std::runtime_error("bad arguments"); skh_voronoi* v = new skh_voronoi();
                      try { v->t = skirt::buildVoronoiMesh(extent6, particles, (size_t)n); } catch (...) { delete v; throw; }
                      *out = v;
                      if (sizes5) { sizes5[0] = v->t.Ncells; sizes5[1] = (int64_t)v->t.nbrIds.size(); sizes5[2] = v->t.nb; sizes5[3] = (int64_t)v->t.blkIds.size(); sizes5[4] = (int64_t)v->t.kdM.size(); } }); }
int skh_voronoi_tables(skh_voronoi* v, double* cellBox, double* volume, double* centroid, int* nbrStart, int* nbrIds, int* blkStart, int* blkIds,
                       int* blkTree, int* kdM, int* kdAxis, int* kdUp, int* kdLeft, int* kdRight)
{ return guarded([&]{ if (!v) throw std::runtime_error("null mesh"); const skirt::VoronoiTables& T = v->t;
                      put(cellBox, T.cellBox); put(volume, T.volume); put(centroid, T.centroid); put(nbrStart, T.nbrStart); put(nbrIds, T.nbrIds);
                      put(blkStart, T.blkStart); put(blkIds, T.blkIds); put(blkTree, T.blkTree); put(kdM, T.kdM); put(kdAxis, T.kdAxis);
                      put(kdUp, T.kdUp); put(kdLeft, T.kdLeft); put(kdRight, T.kdRight); }); }
void skh_voronoi_destroy(skh_voronoi* v) { delete v; }
}
