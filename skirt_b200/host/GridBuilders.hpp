// Host-side construction of the hierarchical / unstructured dust grids, flattened straight into the tables of
// include/skirtgpu.h (skg_grid_tree / skg_grid_amesh / skg_grid_voronoi).  These are the set-up parts of the reference's
// grid classes, restated on flat arrays:
//
//   TreeBuilder          TreeDustGrid::setupSelfBefore (TreeDustGrid.cpp:50-164), subdivide (:168-233),
//                        OctTreeNode / BinTreeNode::createchildren + addneighbors (OctTreeNode.cpp:38-180,
//                        BinTreeNode.cpp:40-330), TreeNode::sortneighbors (TreeNode.cpp:104-158)
//                        ParticleTreeDustGrid::setupSelfBefore (ParticleTreeDustGrid.cpp:36-152) for trees grown around particles
//   AdaptiveMeshBuilder  AdaptiveMesh::AdaptiveMesh (AdaptiveMesh.cpp:21-57), AdaptiveMeshNode (AdaptiveMeshNode.cpp:14-80)
//   VoronoiBuilder       VoronoiMesh::buildMesh / buildTree (VoronoiMesh.cpp:310-393) on top of the Voro++ library
//
// The tree is grown level by level: the caller asks for the boxes of the childless nodes of the current level, estimates
// the dust mass in each (on the GPU: skg_sample_boxes, the device version of TreeNodeSampleDensityCalculator), and hands
// back the subdivision decisions -- node ids, box arithmetic, neighbour lists and their order come out exactly as the
// reference produces them for the same decisions (tests/test_builders.py compares against the reference's own objects).
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace skirt
{

struct TreeTables
{
    int kind = 0, search = 1, Nnodes = 0, Ncells = 0;
    std::vector<double> box;                    // [6N] xmin,ymin,zmin,xmax,ymax,zmax
    std::vector<int> child0, parent, cell, dir, level;
    std::vector<int> nbrStart, nbrIds;          // CSR over (node, wall), lists in the reference's stored order
};

class TreeBuilder
{
public:
    // kind: 0 octree, 1 binary tree (k-d); extent: xmin,xmax,ymin,ymax,zmin,zmax
    TreeBuilder(int kind, const double extent[6], int minLevel, int maxLevel);
    bool done() const { return _frontier.empty(); }
    int frontierLevel() const { return _level; }
    size_t frontierSize() const { return _frontier.size(); }
    // level <= minLevel: every node of the level is subdivided; minLevel < level < maxLevel: the caller decides; level == maxLevel: none
    bool frontierNeedsDecision() const { return _level > _minLevel && _level < _maxLevel; }
    void frontierBoxes(double* box6) const;     // [6 * frontierSize()]
    // flags[frontierSize()] (ignored unless frontierNeedsDecision()); bary[3 * frontierSize()] or null: the barycentres of the dust
    // in the nodes (TreeNodeSampleDensityCalculator::barycenter) for OctTreeDustGrid::barycentric / BinTreeDustGrid's Barycenter method
    void subdivide(const unsigned char* flags, const double* bary = nullptr);
    void finish(int search);                    // cell numbers; neighbour lists when search == 1 (Neighbor)
    // ParticleTreeDustGrid::setupSelfBefore (ParticleTreeDustGrid.cpp:76-152) on a tree that is still its root: the particles
    // are added one by one, a leaf that already holds one is subdivided until the two sit in different leaves; then
    // extraLevels more subdivisions of every leaf; finishes the tables with search = 3 (that grid's own traversal)
    void addParticles(const double* xyz, size_t n, int extraLevels);
    const TreeTables& tables() const { return _t; }
private:
    void createChildren(int l, const double* bary = nullptr);
    int whichNodeFrom(int start, double x, double y, double z) const;
    int addParticle(int p, int start, const double* xyz, std::vector<int>& particlev);
    void addNeighbors(int l);
    void makeNeighbors(int wall1, int node1, int node2);
    void deleteNeighbor(int node, int wall, int other);
    int _kind, _minLevel, _maxLevel, _level = 0;
    std::vector<int> _frontier;
    std::vector<std::vector<int>> _nbr;         // [6N]
    TreeTables _t;
};

struct AMeshTables
{
    int Nnodes = 0, Ncells = 0;
    std::vector<double> box; std::vector<int> nxyz, child0, cell, wallNbr;
    std::vector<double> volume;                 // [Ncells]
    std::vector<int> fileIndex;                 // [Ncells] position of every leaf in the input sequence (for the field values)
};
// nxyz[3n] in the order of the mesh file (AdaptiveMeshAsciiFile.cpp:43-100: depth first, children k -> j -> i); a leaf has 0,0,0
AMeshTables buildAdaptiveMesh(const double extent[6], const int* nxyz, size_t n);

struct VoronoiTables
{
    int Ncells = 0, nb = 0;
    std::vector<double> particles, cellBox, volume, centroid;
    std::vector<int> nbrStart, nbrIds, blkStart, blkIds, blkTree, kdM, kdAxis, kdUp, kdLeft, kdRight;
};
// particles[3n] inside the extent (VoronoiMesh.cpp:262-263 drops the others before the mesh is built)
VoronoiTables buildVoronoiMesh(const double extent[6], const double* particles, size_t n);
bool voronoiAvailable();                        // false when the library was built without Voro++

}   // namespace skirt
