// See GridBuilders.hpp.  Plain C++17 on flat arrays; the only external code is Voro++ (optional, SKIRT_WITH_VORO).
#include "GridBuilders.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <stdexcept>
#ifdef SKIRT_WITH_VORO
#include "container.hh"
#endif

namespace skirt
{

// ---------------------------------------------------------------------------------------------------------------------
// trees
// ---------------------------------------------------------------------------------------------------------------------
// walls as in TreeNode.hpp:100: BACK(-x) FRONT(+x) LEFT(-y) RIGHT(+y) BOTTOM(-z) TOP(+z); axis = wall / 2, high side = wall & 1
static inline int complementWall(int w) { return w ^ 1; }

TreeBuilder::TreeBuilder(int kind, const double extent[6], int minLevel, int maxLevel)
    : _kind(kind), _minLevel(minLevel), _maxLevel(maxLevel)
{
    // TreeDustGrid::setupSelfBefore, TreeDustGrid.cpp:57-59
    if (kind != 0 && kind != 1) throw std::runtime_error("tree kind must be 0 (octree) or 1 (binary tree)");
    if (minLevel < 0) throw std::runtime_error("The minimum tree level should be at least 0");
    if (maxLevel < 2) throw std::runtime_error("The maximum tree level should be at least 2");
    if (maxLevel <= minLevel) throw std::runtime_error("Maximum tree level should be larger than minimum tree level");
    if (!(extent[1] > extent[0]) || !(extent[3] > extent[2]) || !(extent[5] > extent[4])) throw std::runtime_error("The extent of the box should be positive");
    _t.kind = kind;
    const double root[6] = {extent[0], extent[2], extent[4], extent[1], extent[3], extent[5]};
    _t.box.assign(root, root + 6);
    _t.child0.push_back(-1); _t.parent.push_back(-1); _t.dir.push_back(0); _t.level.push_back(0);
    _frontier.push_back(0);
}

void TreeBuilder::frontierBoxes(double* box6) const
{
    for (size_t q = 0; q < _frontier.size(); q++) std::copy_n(_t.box.begin() + 6 * (size_t)_frontier[q], 6, box6 + 6 * q);
}

// OctTreeNode::createchildren (split at the centre, OctTreeNode.cpp:38-57) / BinTreeNode::createchildren (split across
// level % 3, BinTreeNode.cpp:40-77); children get consecutive ids at the end of the node vector
void TreeBuilder::createChildren(int l, const double* bary)
{
    const double b[6] = {_t.box[6 * (size_t)l], _t.box[6 * (size_t)l + 1], _t.box[6 * (size_t)l + 2], _t.box[6 * (size_t)l + 3], _t.box[6 * (size_t)l + 4], _t.box[6 * (size_t)l + 5]};
    const int id = (int)_t.child0.size();
    _t.child0[l] = id;
    auto add = [&](double x0, double y0, double z0, double x1, double y1, double z1)
    {
        const double c[6] = {x0, y0, z0, x1, y1, z1};
        _t.box.insert(_t.box.end(), c, c + 6);
        _t.child0.push_back(-1); _t.parent.push_back(l); _t.dir.push_back(0); _t.level.push_back(_t.level[l] + 1);
    };
    if (_kind == 0)
    {
        // Box::center, or the barycentre of the node's dust (BaryOctTreeNode::createchildren, BaryOctTreeNode.cpp:27-30)
        const double xc = bary ? bary[0] : 0.5 * (b[0] + b[3]), yc = bary ? bary[1] : 0.5 * (b[1] + b[4]), zc = bary ? bary[2] : 0.5 * (b[2] + b[5]);
        add(b[0], b[1], b[2], xc, yc, zc);   add(xc, b[1], b[2], b[3], yc, zc);
        add(b[0], yc, b[2], xc, b[4], zc);   add(xc, yc, b[2], b[3], b[4], zc);
        add(b[0], b[1], zc, xc, yc, b[5]);   add(xc, b[1], zc, b[3], yc, b[5]);
        add(b[0], yc, zc, xc, b[4], b[5]);   add(xc, yc, zc, b[3], b[4], b[5]);
    }
    else
    {
        int dir = _t.level[l] % 3;
        if (bary)
        {
            // BaryBinTreeNode::createchildren, BaryBinTreeNode.cpp:34-58: across the axis along which the barycentre is
            // relatively nearest to a wall (the split itself stays at the centre)
            const double dx = std::min(bary[0] - b[0], b[3] - bary[0]) / (b[3] - b[0]);
            const double dy = std::min(bary[1] - b[1], b[4] - bary[1]) / (b[4] - b[1]);
            const double dz = std::min(bary[2] - b[2], b[5] - bary[2]) / (b[5] - b[2]);
            if (dx < dy) dir = dx < dz ? 0 : 2; else dir = dy < dz ? 1 : 2;
        }
        _t.dir[l] = dir;
        if (dir == 0) { const double xc = 0.5 * (b[0] + b[3]); add(b[0], b[1], b[2], xc, b[4], b[5]); add(xc, b[1], b[2], b[3], b[4], b[5]); }
        else if (dir == 1) { const double yc = 0.5 * (b[1] + b[4]); add(b[0], b[1], b[2], b[3], yc, b[5]); add(b[0], yc, b[2], b[3], b[4], b[5]); }
        else { const double zc = 0.5 * (b[2] + b[5]); add(b[0], b[1], b[2], b[3], b[4], zc); add(b[0], b[1], zc, b[3], b[4], b[5]); }
    }
}

// TreeDustGrid::subdivide (TreeDustGrid.cpp:168-233) for every childless node of the current level, in id order
void TreeBuilder::subdivide(const unsigned char* flags, const double* bary)
{
    std::vector<int> next;
    const bool always = _level <= _minLevel, decide = frontierNeedsDecision();
    if (decide && !flags) throw std::runtime_error("subdivision decisions are needed for this level");
    for (size_t q = 0; q < _frontier.size(); q++)
    {
        if (!(always || (decide && flags[q]))) continue;
        const int l = _frontier[q];
        createChildren(l, (bary && !always) ? bary + 3 * q : nullptr);       // forced levels split regularly, TreeDustGrid.cpp:172-176
        const int nchild = _kind == 0 ? 8 : 2;
        for (int c = 0; c < nchild; c++) next.push_back(_t.child0[l] + c);
    }
    _frontier.swap(next);
    _level++;
}

void TreeBuilder::makeNeighbors(int wall1, int node1, int node2)        // TreeNode::makeneighbors, TreeNode.cpp:188-193
{
    _nbr[6 * (size_t)node1 + wall1].push_back(node2);
    _nbr[6 * (size_t)node2 + complementWall(wall1)].push_back(node1);
}

void TreeBuilder::deleteNeighbor(int node, int wall, int other)         // TreeNode::deleteneighbor, TreeNode.cpp:171-184
{
    std::vector<int>& v = _nbr[6 * (size_t)node + wall];
    for (size_t i = 0; i < v.size(); i++) if (v[i] == other) { v.erase(v.begin() + i); break; }
}

// OctTreeNode::addneighbors (OctTreeNode.cpp:66-180) and BinTreeNode::addneighbors (BinTreeNode.cpp:88-330) in one form.
// A node is split across the axes S (octree: x, y, z; binary tree: its direction); child c sits on the high side of
// axis a when bit a of its octant is set (binary tree: c == 1).
//  - internal neighbours: for every child in index order and every split axis on whose low side it sits, the child and
//    its sibling across that axis -- the order of the reference's explicit list;
//  - for every wall of the node, in wall order, and every neighbour of that wall in list order: the neighbour forgets
//    the node and becomes the neighbour of those children on the wall (in index order) that it touches: across every
//    other split axis b, a low-side child needs neighbour.min_b <= split_b, a high-side child neighbour.max_b >= split_b.
void TreeBuilder::addNeighbors(int l)
{
    const int c0 = _t.child0[l];
    if (c0 < 0) return;
    const int nchild = _kind == 0 ? 8 : 2;
    const int dir = _t.dir[l];
    auto split = [&](int a) { return _kind == 0 || a == dir; };
    auto high = [&](int c, int a) { return _kind == 0 ? ((c >> a) & 1) != 0 : c == 1; };       // only asked for split axes
    auto sibling = [&](int c, int a) { return _kind == 0 ? c + (1 << a) : 1; };
    const double sp[3] = {_t.box[6 * (size_t)c0 + 3], _t.box[6 * (size_t)c0 + 4], _t.box[6 * (size_t)c0 + 5]};    // max corner of child 0
    for (int c = 0; c < nchild; c++)
        for (int a = 0; a < 3; a++)
            if (split(a) && !high(c, a)) makeNeighbors(2 * a + 1, c0 + c, c0 + sibling(c, a));
    for (int w = 0; w < 6; w++)
    {
        const int a = w / 2; const bool hi = (w & 1) != 0;
        const std::vector<int> list = _nbr[6 * (size_t)l + w];      // (the node's own list does not change below; copied for clarity)
        for (int nb : list)
        {
            deleteNeighbor(nb, complementWall(w), l);
            const double* q = &_t.box[6 * (size_t)nb];
            for (int c = 0; c < nchild; c++)
            {
                if (split(a) && high(c, a) != hi) continue;         // the child is not on this wall
                bool touches = true;
                for (int b2 = 0; b2 < 3; b2++)
                {
                    if (b2 == a || !split(b2)) continue;
                    if (high(c, b2) ? !(q[b2 + 3] >= sp[b2]) : !(q[b2] <= sp[b2])) { touches = false; break; }
                }
                if (touches) makeNeighbors(complementWall(w), nb, c0 + c);
            }
        }
    }
}

void TreeBuilder::finish(int search)
{
    if (!done()) throw std::runtime_error("the tree has not been subdivided down to its last level");
    if (search < 0 || search > 3) throw std::runtime_error("invalid search method");
    if (search == 2 && _kind != 0) throw std::runtime_error("Bookkeeping method is not compatible with binary tree");     // BinTreeDustGrid.cpp:19-25
    const int N = (int)_t.child0.size();
    _t.Nnodes = N; _t.search = search;
    _t.cell.assign(N, -1);
    int m = 0;
    for (int l = 0; l < N; l++) if (_t.child0[l] < 0) _t.cell[l] = m++;     // TreeDustGrid.cpp:112-122
    _t.Ncells = m;
    _t.nbrStart.clear(); _t.nbrIds.clear();
    if (search != 1) return;
    _nbr.assign(6 * (size_t)N, std::vector<int>());
    for (int l = 0; l < N; l++) addNeighbors(l);
    // TreeNode::sortneighbors (TreeNode.cpp:144-151): every list by decreasing overlap with the node's wall
    for (int l = 0; l < N; l++)
        for (int w = 0; w < 6; w++)
        {
            std::vector<int>& v = _nbr[6 * (size_t)l + w];
            if (v.size() < 2) continue;
            const int a1 = w < 2 ? 1 : 0, a2 = w < 4 ? 2 : 1;     // the wall's in-plane axes: (y,z), (x,z), (x,y)
            const double* b = &_t.box[6 * (size_t)l];
            auto overlap = [&](int n)
            {
                const double* q = &_t.box[6 * (size_t)n];
                return std::max(std::min(b[a1 + 3], q[a1 + 3]) - std::max(b[a1], q[a1]), 0.) * std::max(std::min(b[a2 + 3], q[a2 + 3]) - std::max(b[a2], q[a2]), 0.);
            };
            std::sort(v.begin(), v.end(), [&](int n1, int n2) { return overlap(n1) > overlap(n2); });
        }
    _t.nbrStart.resize(6 * (size_t)N + 1);
    for (size_t q = 0; q < 6 * (size_t)N; q++) { _t.nbrStart[q] = (int)_t.nbrIds.size(); _t.nbrIds.insert(_t.nbrIds.end(), _nbr[q].begin(), _nbr[q].end()); }
    _t.nbrStart[6 * (size_t)N] = (int)_t.nbrIds.size();
    std::vector<std::vector<int>>().swap(_nbr);
}

// TreeNode::whichnode(Vec) (TreeNode.cpp:70-80) from a given node: closed-box test, then OctTreeNode::child (OctTreeNode.cpp:184-189)
// or BinTreeNode::child (BinTreeNode.cpp:326-335) down to a leaf
int TreeBuilder::whichNodeFrom(int start, double x, double y, double z) const
{
    const double* b = &_t.box[6 * (size_t)start];
    if (!(x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5])) return -1;
    int l = start;
    while (_t.child0[l] >= 0)
    {
        const int c0 = _t.child0[l]; const double* c = &_t.box[6 * (size_t)c0];      // child 0: its max corner is the split point
        if (_kind == 0) l = c0 + (x < c[3] ? 0 : 1) + (y < c[4] ? 0 : 2) + (z < c[5] ? 0 : 4);
        else { const int d = _t.dir[l]; const double v = d == 0 ? x : d == 1 ? y : z; l = c0 + (v < c[3 + d] ? 0 : 1); }
    }
    return l;
}

// addParticleToNode (ParticleTreeDustGrid.cpp:36-72); returns the level of the leaf that received the particle, -1 if outside
int TreeBuilder::addParticle(int p, int start, const double* xyz, std::vector<int>& particlev)
{
    const int node = whichNodeFrom(start, xyz[3 * (size_t)p], xyz[3 * (size_t)p + 1], xyz[3 * (size_t)p + 2]);
    if (node < 0) return -1;
    if (particlev[node] < 0) { particlev[node] = p; return _t.level[node]; }
    if (_t.level[node] > 1000) throw std::runtime_error("two particles share a position: the particle tree cannot separate them");
    createChildren(node);
    particlev.resize(_t.child0.size(), -1);
    addParticle(particlev[node], node, xyz, particlev);
    return addParticle(p, node, xyz, particlev);
}

void TreeBuilder::addParticles(const double* xyz, size_t n, int extraLevels)
{
    if (_t.child0.size() != 1) throw std::runtime_error("particles are added to a tree that is still its root node");
    if (!xyz && n) throw std::runtime_error("null particle list");
    std::vector<int> particlev(1, -1);
    for (size_t i = 0; i < n; i++) addParticle((int)i, 0, xyz, particlev);
    for (int e = 0; e < extraLevels; e++)
    {
        const int N = (int)_t.child0.size();
        for (int l = 0; l < N; l++) if (_t.child0[l] < 0) createChildren(l);
    }
    _frontier.clear();
    finish(3);
}

// ---------------------------------------------------------------------------------------------------------------------
// adaptive mesh
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
struct AmNode { double b[6]; int nx, ny, nz, firstChild, m; };       // children are consecutive in the depth-first vector? no: see below

inline bool contains(const double* b, double x, double y, double z) { return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5]; }
inline int cellIndex1(double v, double lo, double hi, int n)          // Box::cellindices, Box.hpp:134-139 (C truncation)
{ return std::max(0, std::min(n - 1, static_cast<int>(n * (v - lo) / (hi - lo)))); }
}

AMeshTables buildAdaptiveMesh(const double extent[6], const int* nxyz, size_t n)
{
    if (n < 1) throw std::runtime_error("Reached end of file in mesh data before all nodes were read");
    // pass 1: the recursion of AdaptiveMeshNode::AdaptiveMeshNode over the depth-first sequence, with an explicit stack;
    // nodes are collected breadth first so that the children of a node are consecutive in local Morton order
    struct Pending { int node; };
    AMeshTables T;
    // depth-first construction needs the subtree sizes; do it recursively on the input sequence with a cursor
    struct Tmp { double b[6]; int nx, ny, nz, m; std::vector<int> kids; };
    std::vector<Tmp> tmp; tmp.reserve(n);
    size_t cursor = 0; int leaves = 0;
    // iterative DFS: stack of (tmp index, next child ordinal)
    std::vector<std::pair<int, int>> stack;
    auto newNode = [&](const double* b) -> int
    {
        if (cursor >= n) throw std::runtime_error("Reached end of file in mesh data before all nodes were read");
        Tmp t; std::copy_n(b, 6, t.b); t.nx = nxyz[3 * cursor]; t.ny = nxyz[3 * cursor + 1]; t.nz = nxyz[3 * cursor + 2]; cursor++;
        const bool leaf = t.nx == 0 && t.ny == 0 && t.nz == 0;
        if (!leaf && (t.nx < 1 || t.ny < 1 || t.nz < 1)) throw std::runtime_error("invalid number of child nodes in mesh data");
        t.m = leaf ? leaves++ : -1;
        if (leaf) T.fileIndex.push_back((int)cursor - 1);
        tmp.push_back(std::move(t));
        return (int)tmp.size() - 1;
    };
    const double root[6] = {extent[0], extent[2], extent[4], extent[1], extent[3], extent[5]};
    stack.emplace_back(newNode(root), 0);
    while (!stack.empty())
    {
        const int p = stack.back().first; const int ord = stack.back().second;
        const int Nx = tmp[p].nx, Ny = tmp[p].ny, Nz = tmp[p].nz;
        if (tmp[p].m >= 0 || ord >= Nx * Ny * Nz) { stack.pop_back(); continue; }
        stack.back().second++;
        const int i = ord % Nx, j = (ord / Nx) % Ny, k = ord / (Nx * Ny);          // loops k -> j -> i, i fastest
        const double* e = tmp[p].b;
        // Box::fracpos(int...), Box.hpp:129-130: min + d*(max-min)/n
        const double cb[6] = {e[0] + i * (e[3] - e[0]) / Nx, e[1] + j * (e[4] - e[1]) / Ny, e[2] + k * (e[5] - e[2]) / Nz,
                              e[0] + (i + 1) * (e[3] - e[0]) / Nx, e[1] + (j + 1) * (e[4] - e[1]) / Ny, e[2] + (k + 1) * (e[5] - e[2]) / Nz};
        const int c = newNode(cb);
        tmp[p].kids.push_back(c);
        stack.emplace_back(c, 0);
    }
    if (cursor != n) throw std::runtime_error("Superfluous data in mesh data after all nodes were read");
    // pass 2: breadth-first numbering
    const int N = (int)tmp.size();
    std::vector<int> order; order.reserve(N); std::vector<int> index(N, -1);
    order.push_back(0); index[0] = 0;
    for (size_t q = 0; q < order.size(); q++) for (int c : tmp[order[q]].kids) { index[c] = (int)order.size(); order.push_back(c); }
    T.Nnodes = N; T.Ncells = leaves;
    T.box.resize(6 * (size_t)N); T.nxyz.resize(3 * (size_t)N); T.child0.resize(N); T.cell.resize(N); T.wallNbr.assign(6 * (size_t)N, -1);
    T.volume.resize(leaves);
    for (int l = 0; l < N; l++)
    {
        const Tmp& t = tmp[order[l]];
        std::copy_n(t.b, 6, &T.box[6 * (size_t)l]);
        T.nxyz[3 * (size_t)l] = t.nx; T.nxyz[3 * (size_t)l + 1] = t.ny; T.nxyz[3 * (size_t)l + 2] = t.nz;
        T.cell[l] = t.m; T.child0[l] = t.kids.empty() ? -1 : index[t.kids[0]];
        if (t.m >= 0) T.volume[t.m] = (t.b[3] - t.b[0]) * (t.b[4] - t.b[1]) * (t.b[5] - t.b[2]);
    }
    // AdaptiveMeshNode::whichnode(Vec) from the root (AdaptiveMeshNode.cpp:132-142 + child :109-128)
    auto whichnode = [&](double x, double y, double z) -> int
    {
        if (!contains(&T.box[0], x, y, z)) return -1;
        int node = 0;
        while (T.child0[node] >= 0)
        {
            const double* b = &T.box[6 * (size_t)node];
            const int Nx = T.nxyz[3 * (size_t)node], Ny = T.nxyz[3 * (size_t)node + 1], Nz = T.nxyz[3 * (size_t)node + 2];
            int i = cellIndex1(x, b[0], b[3], Nx), j = cellIndex1(y, b[1], b[4], Ny), k = cellIndex1(z, b[2], b[5], Nz);
            int child = T.child0[node] + (k * Ny + j) * Nx + i;
            const double* c = &T.box[6 * (size_t)child];
            if (!contains(c, x, y, z))
            {
                if (x < c[0]) i--; else if (x > c[3]) i++;
                if (y < c[1]) j--; else if (y > c[4]) j++;
                if (z < c[2]) k--; else if (z > c[5]) k++;
                if (i < 0 || i >= Nx || j < 0 || j >= Ny || k < 0 || k >= Nz) throw std::runtime_error("Can't locate the appropriate child node");
                child = T.child0[node] + (k * Ny + j) * Nx + i;
                if (!contains(&T.box[6 * (size_t)child], x, y, z)) throw std::runtime_error("Can't locate the appropriate child node");
            }
            node = child;
        }
        return node;
    };
    // AdaptiveMesh::addNeighbors (AdaptiveMesh.cpp:93-99) -> AdaptiveMeshNode::addNeighbors (AdaptiveMeshNode.cpp:61-80)
    const double wx = root[3] - root[0], wy = root[4] - root[1], wz = root[5] - root[2];
    const double eps = 1e-12 * std::sqrt(wx * wx + wy * wy + wz * wz);       // AdaptiveMesh.cpp:52
    for (int l = 0; l < N; l++)
    {
        if (T.cell[l] < 0) continue;
        const double* b = &T.box[6 * (size_t)l];
        const double xc = (b[0] + b[3]) / 2., yc = (b[1] + b[4]) / 2., zc = (b[2] + b[5]) / 2.;
        int* w = &T.wallNbr[6 * (size_t)l];
        w[0] = whichnode(b[0] - eps, yc, zc); w[1] = whichnode(b[3] + eps, yc, zc);
        w[2] = whichnode(xc, b[1] - eps, zc); w[3] = whichnode(xc, b[4] + eps, zc);
        w[4] = whichnode(xc, yc, b[2] - eps); w[5] = whichnode(xc, yc, b[5] + eps);
    }
    return T;
}

// ---------------------------------------------------------------------------------------------------------------------
// Voronoi
// ---------------------------------------------------------------------------------------------------------------------
bool voronoiAvailable()
{
#ifdef SKIRT_WITH_VORO
    return true;
#else
    return false;
#endif
}

#ifdef SKIRT_WITH_VORO
namespace
{
// lessthan(p1, p2, axis), VoronoiMesh.cpp:77-105: lexicographic from the split axis onwards
inline bool lessThan(const double* p1, const double* p2, int axis)
{
    for (int q = 0; q < 3; q++)
    {
        const int a = (axis + q) % 3;
        if (p1[a] < p2[a]) return true;
        if (p1[a] > p2[a]) return false;
    }
    return false;
}
}
#endif

VoronoiTables buildVoronoiMesh(const double extent[6], const double* particles, size_t n)
{
#ifndef SKIRT_WITH_VORO
    (void)extent; (void)particles; (void)n;
    throw std::runtime_error("this build of the host library has no Voro++ (Voronoi tessellation); rebuild where the Voro++ sources are available");
#else
    VoronoiTables T;
    if (n < 1) throw std::runtime_error("a Voronoi grid needs at least one particle");
    if (n > 2000000000ull) throw std::runtime_error("too many particles");
    const int N = (int)n;
    const double xmin = extent[0], xmax = extent[1], ymin = extent[2], ymax = extent[3], zmin = extent[4], zmax = extent[5];
    for (int m = 0; m < N; m++)
    {
        const double* p = particles + 3 * (size_t)m;
        if (!(p[0] >= xmin && p[0] <= xmax && p[1] >= ymin && p[1] <= ymax && p[2] >= zmin && p[2] <= zmax))
            throw std::runtime_error("particle " + std::to_string(m) + " lies outside the extent of the grid");
    }
    // VoronoiMesh::buildMesh, VoronoiMesh.cpp:310-376
    T.Ncells = N;
    T.nb = std::max(3, std::min(1000, static_cast<int>(3. * std::pow((double)N, 1. / 3.))));
    const int nb = T.nb; const size_t nb3 = (size_t)nb * nb * nb;
    const double wx = xmax - xmin, wy = ymax - ymin, wz = zmax - zmin;
    const double eps = 1e-12 * std::sqrt(wx * wx + wy * wy + wz * wz);
    T.particles.assign(particles, particles + 3 * n);
    T.cellBox.resize(6 * n); T.volume.resize(n); T.centroid.resize(3 * n);
    std::vector<std::vector<int>> nbrs(n), blocks(nb3);
    voro::container con(xmin, xmax, ymin, ymax, zmin, zmax, nb, nb, nb, false, false, false, 8);
    for (int m = 0; m < N; m++) con.put(m, particles[3 * (size_t)m], particles[3 * (size_t)m + 1], particles[3 * (size_t)m + 2]);
    voro::c_loop_all loop(con);
    if (loop.start()) do
    {
        voro::voronoicell_neighbor cell;
        if (!con.compute_cell(cell, loop)) throw std::runtime_error("Can't compute Voronoi cell " + std::to_string(loop.pid()));
        const int m = loop.pid();
        const double* p = particles + 3 * (size_t)m;
        // VoronoiCell::init, VoronoiMesh.cpp:36-57
        double cx, cy, cz; cell.centroid(cx, cy, cz);
        T.centroid[3 * (size_t)m] = cx + p[0]; T.centroid[3 * (size_t)m + 1] = cy + p[1]; T.centroid[3 * (size_t)m + 2] = cz + p[2];
        T.volume[m] = cell.volume();
        std::vector<double> coords; cell.vertices(p[0], p[1], p[2], coords);
        double b[6] = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (size_t i = 0; i + 2 < coords.size(); i += 3)
            for (int a = 0; a < 3; a++) { b[a] = std::min(b[a], coords[i + a]); b[a + 3] = std::max(b[a + 3], coords[i + a]); }
        std::copy_n(b, 6, &T.cellBox[6 * (size_t)m]);
        cell.neighbors(nbrs[m]);
        // the blocks the cell's enclosing box overlaps, VoronoiMesh.cpp:352-361 (Box::cellindices)
        const int i1 = cellIndex1(b[0] - eps, xmin, xmax, nb), j1 = cellIndex1(b[1] - eps, ymin, ymax, nb), k1 = cellIndex1(b[2] - eps, zmin, zmax, nb);
        const int i2 = cellIndex1(b[3] + eps, xmin, xmax, nb), j2 = cellIndex1(b[4] + eps, ymin, ymax, nb), k2 = cellIndex1(b[5] + eps, zmin, zmax, nb);
        for (int i = i1; i <= i2; i++) for (int j = j1; j <= j2; j++) for (int k = k1; k <= k2; k++)
            blocks[((size_t)i * nb + j) * nb + k].push_back(m);
    }
    while (loop.inc());
    T.nbrStart.resize(n + 1);
    for (int m = 0; m < N; m++) { T.nbrStart[m] = (int)T.nbrIds.size(); T.nbrIds.insert(T.nbrIds.end(), nbrs[m].begin(), nbrs[m].end()); }
    T.nbrStart[N] = (int)T.nbrIds.size();
    // per-block search trees for blocks with more than five cells, VoronoiMesh.cpp:366-393 (nth_element medians;
    // the ids of a block are permuted by the construction, like the reference's _blocklists)
    T.blkTree.assign(nb3, -1);
    struct Frame { size_t first, last; int depth, up; bool left; };
    for (size_t bq = 0; bq < nb3; bq++)
    {
        std::vector<int>& ids = blocks[bq];
        if (ids.size() <= 5) continue;
        // recursive buildTree(first, last, depth) in pre-order: node, left subtree, right subtree
        std::vector<Frame> st; st.push_back({0, ids.size(), 0, -1, false});
        while (!st.empty())
        {
            const Frame f = st.back(); st.pop_back();
            const size_t length = f.last - f.first;
            if (length == 0) continue;
            const int axis = f.depth % 3;
            const size_t median = length >> 1;
            std::nth_element(ids.begin() + f.first, ids.begin() + f.first + median, ids.begin() + f.last,
                             [&](int m1, int m2) { return m1 != m2 && lessThan(&T.particles[3 * (size_t)m1], &T.particles[3 * (size_t)m2], axis); });
            const int node = (int)T.kdM.size();
            T.kdM.push_back(ids[f.first + median]); T.kdAxis.push_back(axis); T.kdUp.push_back(f.up); T.kdLeft.push_back(-1); T.kdRight.push_back(-1);
            if (f.up >= 0) { if (f.left) T.kdLeft[f.up] = node; else T.kdRight[f.up] = node; }
            else T.blkTree[bq] = node;
            // pre-order: the left subtree is numbered before the right one (push right first)
            st.push_back({f.first + median + 1, f.last, f.depth + 1, node, false});
            st.push_back({f.first, f.first + median, f.depth + 1, node, true});
        }
    }
    T.blkStart.resize(nb3 + 1);
    for (size_t bq = 0; bq < nb3; bq++) { T.blkStart[bq] = (int)T.blkIds.size(); T.blkIds.insert(T.blkIds.end(), blocks[bq].begin(), blocks[bq].end()); }
    T.blkStart[nb3] = (int)T.blkIds.size();
    return T;
#endif
}

}   // namespace skirt
