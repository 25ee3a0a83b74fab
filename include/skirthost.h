/* skirthost.h -- C ABI of the host-side set-up library (libskirthost.so): construction of the hierarchical and
 * unstructured dust grids as the flat tables that skg_grid_tree / skg_grid_amesh / skg_grid_voronoi (skirtgpu.h) take.
 *
 * This is NOT the drop-in boundary of the hot path (that is skirtgpu.h); it is the set-up side of the reference's grid
 * classes, which a maintainer binding the engine into SKIRT does not need (the adapter flattens SKIRT's own objects,
 * INTEGRATION.md 2.1) but which a stand-alone host (skirt_b200/host, skirt_b200/simulation.py, bench.py) does:
 *   skh_tree_*      TreeDustGrid::setupSelfBefore / subdivide (TreeDustGrid.cpp:50-233), OctTreeNode / BinTreeNode
 *                   createchildren + addneighbors, TreeNode::sortneighbors (TreeNode.cpp:104-158)
 *   skh_amesh_*     AdaptiveMesh::AdaptiveMesh + addNeighbors (AdaptiveMesh.cpp:21-99, AdaptiveMeshNode.cpp:14-80)
 *   skh_voronoi_*   VoronoiMesh::buildMesh / buildTree (VoronoiMesh.cpp:310-393) over the Voro++ library
 * All functions return 0 on success; skh_last_error() holds the message otherwise.  Host pointers only; extents are
 * xmin,xmax,ymin,ymax,zmin,zmax (the order of the Box setters).
 */
#ifndef SKIRTHOST_H
#define SKIRTHOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* skh_last_error(void);

/* ---- octree / binary tree, grown level by level --------------------------------------------------------------------
 * kind 0 octree, 1 binary tree.  Loop: skh_tree_frontier -> (size > 0) -> skh_tree_frontier_boxes -> estimate the mass in
 * every box (skg_sample_boxes, or skg_sample_boxes_dispersion when maxDensDispFraction > 0) -> skh_tree_subdivide(flags) ... until size == 0; then skh_tree_finish + skh_tree_tables.
 * needsDecision is 0 for the levels <= minLevel (every node is subdivided: flags may be NULL). */
typedef struct skh_tree skh_tree;
int skh_tree_create(int kind, const double* extent6, int minLevel, int maxLevel, skh_tree** out);
void skh_tree_destroy(skh_tree* t);
int skh_tree_frontier(skh_tree* t, int* level, int64_t* size, int* needsDecision);
int skh_tree_frontier_boxes(skh_tree* t, double* box6 /* [6*size] xmin,ymin,zmin,xmax,ymax,zmax */);
int skh_tree_subdivide(skh_tree* t, const unsigned char* flags /* [size] */);
/* the same with the barycentre of the dust in every frontier node (TreeNodeSampleDensityCalculator::barycenter,
 * TreeNodeSampleDensityCalculator.cpp:48-58), barycenters[3*size] or null: an octree node is split at its barycentre
 * (OctTreeDustGrid::barycentric, BaryOctTreeNode.cpp:27-30), a binary-tree node across the axis along which the barycentre is
 * relatively nearest to a wall (BinTreeDustGrid::directionMethod Barycenter, BaryBinTreeNode.cpp:34-58).  Levels <= minLevel
 * split regularly (TreeDustGrid.cpp:172-176). */
int skh_tree_subdivide_at(skh_tree* t, const unsigned char* flags /* [size] */, const double* barycenters /* [3*size] */);
/* search: 0 TopDown, 1 Neighbor (builds the sorted neighbour lists), 2 Bookkeeping (octree only) */
int skh_tree_finish(skh_tree* t, int search, int* Nnodes, int* Ncells, int64_t* Nneighbours);
/* ParticleTreeDustGrid::setupSelfBefore (ParticleTreeDustGrid.cpp:76-152): an octree (kind 0) or binary tree (kind 1) grown around
 * particles[3n] -- added in order, a leaf holding a particle is subdivided until the two are apart -- plus extraLevels subdivisions
 * of every leaf.  The tables (skh_tree_tables; no neighbour lists) go to skg_grid_tree with search = 3. */
int skh_ptree_build(int kind, const double* extent6, const double* particles, int64_t n, int extraLevels, skh_tree** out, int* Nnodes, int* Ncells);
int skh_tree_tables(skh_tree* t, double* box, int* child0, int* parent, int* cell, int* dir, int* level, int* nbrStart, int* nbrIds);

/* ---- adaptive mesh ---------------------------------------------------------------------------------------------------
 * nxyz[3n]: the nodes in the order of the mesh file (AdaptiveMeshAsciiFile.cpp:43-100: depth first, children k -> j -> i),
 * 0,0,0 for a leaf.  fileIndex[Ncells]: position of every cell's line in that sequence (to pick up its field values). */
typedef struct skh_amesh skh_amesh;
int skh_amesh_build(const double* extent6, const int* nxyz, int64_t n, skh_amesh** out, int* Nnodes, int* Ncells);
int skh_amesh_tables(skh_amesh* a, double* box, int* nxyz, int* child0, int* cell, int* wallNbr, double* volume, int* fileIndex);
void skh_amesh_destroy(skh_amesh* a);

/* ---- Voronoi mesh ----------------------------------------------------------------------------------------------------
 * sizes: [0] Ncells [1] neighbours in total [2] nb (blocks per axis) [3] block references in total [4] kd-tree nodes */
typedef struct skh_voronoi skh_voronoi;
int skh_voronoi_available(void);
int skh_voronoi_build(const double* extent6, const double* particles, int64_t n, skh_voronoi** out, int64_t* sizes5);
int skh_voronoi_tables(skh_voronoi* v, double* cellBox, double* volume, double* centroid, int* nbrStart, int* nbrIds, int* blkStart, int* blkIds,
                       int* blkTree, int* kdM, int* kdAxis, int* kdUp, int* kdLeft, int* kdRight);
void skh_voronoi_destroy(skh_voronoi* v);

#ifdef __cplusplus
}
#endif
#endif
