// Engine object + C ABI (include/skirtgpu.h).  Product path: fails loudly without a CUDA device; there is
// no CPU fallback and nothing here touches the test oracles.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include "engine.h"

namespace skg
{

static thread_local std::string g_lastError;
void setLastError(const std::string& msg) { g_lastError = msg; }
const char* lastErrorCStr() { return g_lastError.c_str(); }

Engine::Engine(int dev) : device(dev)
{
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        throw Error(std::string("no CUDA device available (") + cudaGetErrorString(err) + "); the engine has no CPU fallback");
    if (dev < 0 || dev >= count) throw Error("invalid CUDA device index " + std::to_string(dev));
    SKG_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop; SKG_CUDA(cudaGetDeviceProperties(&prop, dev));
    smCount = prop.multiProcessorCount;
    SKG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    counters.ensure(sizeof(Counters));
    SKG_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(Counters), stream));
    sync();
}

Engine::~Engine()
{
    cudaSetDevice(device);
    destroyComm(*this);
    freeGrid();
    for (DevBuf* b : sourceBufs) delete b;
    for (DevBuf* b : instrBufs) delete b;
    for (DevBuf* b : spareBufs) delete b;
    for (Shadow& sh : shadows) delete sh.buf;
    if (copyStream) cudaStreamDestroy(copyStream);
    if (snapEvent) cudaEventDestroy(snapEvent);
    if (mcHostCounts) cudaFreeHost(mcHostCounts);
    if (smallHost) cudaFreeHost(smallHost);
    for (cudaEvent_t ev : mcEvents) if (ev) cudaEventDestroy(ev);
    if (stream) cudaStreamDestroy(stream);
}

DevBuf* Engine::takeBuf(size_t bytes)
{
    int best = -1;
    for (int i = 0; i < (int)spareBufs.size(); i++)
        if (spareBufs[i]->bytes >= bytes && (best < 0 || spareBufs[i]->bytes < spareBufs[best]->bytes)) best = i;
    if (best < 0) return new DevBuf();
    DevBuf* b = spareBufs[best]; spareBufs.erase(spareBufs.begin() + best);
    return b;
}
void Engine::recycle(std::vector<DevBuf*>& list)
{
    for (DevBuf* b : list) { if (spareBufs.size() < 256) spareBufs.push_back(b); else delete b; }
    list.clear();
}

void Engine::freeGrid()
{
    recycle(gridBufs);
    gridKind = GRID_NONE; Ncells = 0;
}

__global__ void publishWords(const unsigned long long* __restrict__ src, unsigned long long* host, int n)
{
    if ((int)threadIdx.x < n) reinterpret_cast<volatile unsigned long long*>(host)[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
void Engine::readSmall(void* dst, const void* devSrc, size_t bytes)
{
    if (bytes > 512 || bytes % 8) throw Error("readSmall: at most 512 bytes in 8-byte words");
    if (!smallHost) SKG_CUDA(cudaMallocHost(&smallHost, 512));
    publishWords<<<1, 64, 0, stream>>>(static_cast<const unsigned long long*>(devSrc), static_cast<unsigned long long*>(smallHost), (int)(bytes / 8));
    SKG_CUDA(cudaGetLastError());
    sync();
    std::memcpy(dst, smallHost, bytes);
}

Counters Engine::readCounters()
{
    static_assert(sizeof(Counters) % 8 == 0 && sizeof(Counters) <= 512, "Counters travel through readSmall");
    Counters c;
    readSmall(&c, counters.p, sizeof(Counters));
    return c;
}

template<class T> static const T* up(Engine& e, const T* host, size_t n)
{
    DevBuf* b = e.takeBuf(n * sizeof(T)); e.gridBufs.push_back(b);
    b->upload(host, n * sizeof(T), e.stream);
    return b->as<T>();
}

// eps = 1e-12 * extent.widths().norm(), TreeDustGrid.cpp:76 / VoronoiMesh.cpp:234 / AdaptiveMesh.cpp:52
static double epsFor(double wx, double wy, double wz)
{
    volatile double a = wx * wx, b = wy * wy, c = wz * wz;   // no contraction
    volatile double sum = a + b; sum = sum + c;
    return 1e-12 * std::sqrt(sum);
}

}   // namespace skg

using namespace skg;

template<class F> static int guarded(F f)
{
    try { f(); return 0; }
    catch (std::exception& ex) { setLastError(ex.what()); }
    catch (...) { setLastError("unknown error"); }
    return 1;
}

static Engine& E(skg_engine* e)
{
    if (!e) throw Error("null engine");
    Engine& en = *reinterpret_cast<Engine*>(e);
    SKG_CUDA(cudaSetDevice(en.device));
    return en;
}

extern "C"
{

const char* skg_last_error(void) { return skg::lastErrorCStr(); }
int skg_version(void) { return 1; }

int skg_engine_create(int device, skg_engine** out)
{
    return guarded([&]{ if (!out) throw Error("null out pointer"); *out = reinterpret_cast<skg_engine*>(new Engine(device)); });
}

void skg_engine_destroy(skg_engine* e) { delete reinterpret_cast<Engine*>(e); }

int skg_stream(skg_engine* eh, void** stream)
{ return guarded([&]{ if (!stream) throw Error("null output"); *stream = (void*)E(eh).stream; }); }
int skg_launch_count(skg_engine* eh, uint64_t* launches)
{ return guarded([&]{ if (!launches) throw Error("null output"); *launches = E(eh).launches; }); }

int skg_host_alloc(size_t bytes, void** ptr)
{ return guarded([&]{ if (!ptr) throw Error("null output"); SKG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault)); }); }
int skg_host_free(void* ptr) { return guarded([&]{ if (ptr) SKG_CUDA(cudaFreeHost(ptr)); }); }

int skg_copy_to_host(skg_engine* eh, const void* d_src, void* host, size_t bytes)
{ return guarded([&]{ Engine& e = E(eh); if (!d_src || !host) throw Error("null pointer"); SKG_CUDA(cudaMemcpyAsync(host, d_src, bytes, cudaMemcpyDeviceToHost, e.stream)); e.sync(); }); }

int skg_num_cells(skg_engine* e) { return e ? reinterpret_cast<Engine*>(e)->Ncells : 0; }

int skg_grid_cartesian(skg_engine* eh, const double* xv, int Nx, const double* yv, int Ny, const double* zv, int Nz)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (Nx < 1 || Ny < 1 || Nz < 1 || !xv || !yv || !zv) throw Error("cartesian grid needs at least one bin per axis");
        if ((int64_t)Nx * Ny * Nz > 2147483647LL) throw Error("too many cells for int32 cell numbers");
        for (int i = 0; i < Nx; i++) if (!(xv[i] < xv[i+1])) throw Error("x borders must be strictly ascending");
        for (int i = 0; i < Ny; i++) if (!(yv[i] < yv[i+1])) throw Error("y borders must be strictly ascending");
        for (int i = 0; i < Nz; i++) if (!(zv[i] < zv[i+1])) throw Error("z borders must be strictly ascending");
        e.freeGrid();
        e.cart.xv = up(e, xv, Nx + 1); e.cart.yv = up(e, yv, Ny + 1); e.cart.zv = up(e, zv, Nz + 1);
        e.cart.Nx = Nx; e.cart.Ny = Ny; e.cart.Nz = Nz; e.cart.sx = e.cart.sy = e.cart.sz = 0; e.cart.staged = 0; e.cart.rhoAhead = nullptr; e.cart.rhoAheadStride = 0;
        // BoxDustGrid extent: for every Mesh of the reference mesh[0]=0 and mesh[N]=1, so that the borders'
        // end points equal the extent (CartesianDustGrid.cpp:34-36)
        e.cart.ext[0] = xv[0]; e.cart.ext[1] = xv[Nx]; e.cart.ext[2] = yv[0]; e.cart.ext[3] = yv[Ny]; e.cart.ext[4] = zv[0]; e.cart.ext[5] = zv[Nz];
        // uniform meshes (LinMesh on every axis, NR::lingrid NR.hpp:171-176): the shooting stages then step the exit
        // parameters by a constant per axis instead of reading the borders (CartFastWalker<true>)
        auto uniformAxis = [](const double* v, int n) {
            const double w = (v[n] - v[0]) / n;
            for (int i = 0; i <= n; i++) if (std::fabs(v[i] - (v[0] + i * w)) > 1e-12 * std::fabs(v[n] - v[0])) return false;
            return true; };
        e.cart.uniform = uniformAxis(xv, Nx) && uniformAxis(yv, Ny) && uniformAxis(zv, Nz);
        if (const char* v = getenv("SKG_CART_UNIFORM")) e.cart.uniform = e.cart.uniform && atoi(v) != 0;
        e.cart.wx = (xv[Nx] - xv[0]) / Nx; e.cart.wy = (yv[Ny] - yv[0]) / Ny; e.cart.wz = (zv[Nz] - zv[0]) / Nz;
        e.gridKind = GRID_CART; e.Ncells = Nx * Ny * Nz;
        e.sync();
    });
}

// ---- grids with symmetries ------------------------------------------------------------------------------------------------
static void checkBorders(const double* v, int n, const char* what)
{
    if (!v || n < 1) throw Error(std::string("grid borders missing: ") + what);
    for (int i = 0; i < n; i++) if (!(v[i + 1] > v[i])) throw Error(std::string("grid borders must increase: ") + what);
}
int skg_grid_sphere1d(skg_engine* eh, int Nr, const double* rv)
{
    return guarded([&]{
        Engine& e = E(eh);
        checkBorders(rv, Nr, "radial");
        if (!(rv[Nr] > 0)) throw Error("The outer radius of the grid should be positive");       // SphereDustGrid.cpp:21-26
        e.freeGrid();
        e.sym = SymGrid{}; e.sym.sub = 0; e.sym.N1 = Nr; e.sym.N2 = 0; e.sym.v1 = up(e, rv, Nr + 1); e.sym.rmax = rv[Nr];
        e.gridKind = GRID_SYM; e.Ncells = Nr;
        e.sync();
    });
}
int skg_grid_sphere2d(skg_engine* eh, int Nr, const double* rv, int Ntheta, const double* thetav, const double* cv)
{
    return guarded([&]{
        Engine& e = E(eh);
        checkBorders(rv, Nr, "radial"); checkBorders(thetav, Ntheta, "polar");
        if (!cv) throw Error("grid borders missing: cosines of the polar borders");
        if (!(rv[Nr] > 0)) throw Error("The outer radius of the grid should be positive");
        // Sphere2DDustGrid::path needs a border in the xy-plane (Sphere2DDustGrid.cpp:39-72 inserts one when the mesh has none)
        int zeros = 0; for (int k = 1; k < Ntheta; k++) if (cv[k] == 0.0) zeros++;
        if (zeros != 1) throw Error("the polar borders need exactly one grid point at pi/2 (cosine exactly 0)");
        e.freeGrid();
        e.sym = SymGrid{}; e.sym.sub = 1; e.sym.N1 = Nr; e.sym.N2 = Ntheta; e.sym.v1 = up(e, rv, Nr + 1); e.sym.v2 = up(e, thetav, Ntheta + 1);
        e.sym.cv = up(e, cv, Ntheta + 1); e.sym.rmax = rv[Nr];
        e.gridKind = GRID_SYM; e.Ncells = Nr * Ntheta;
        e.sync();
    });
}
int skg_grid_cylinder2d(skg_engine* eh, int NR, const double* Rv, int Nz, const double* zv)
{
    return guarded([&]{
        Engine& e = E(eh);
        checkBorders(Rv, NR, "radial"); checkBorders(zv, Nz, "vertical");
        if (!(Rv[NR] > 0)) throw Error("The outer radius of the grid should be positive");       // CylinderDustGrid.cpp
        e.freeGrid();
        e.sym = SymGrid{}; e.sym.sub = 2; e.sym.N1 = NR; e.sym.N2 = Nz; e.sym.v1 = up(e, Rv, NR + 1); e.sym.v2 = up(e, zv, Nz + 1);
        e.sym.rmax = Rv[NR]; e.sym.zmin = zv[0]; e.sym.zmax = zv[Nz];
        e.gridKind = GRID_SYM; e.Ncells = NR * Nz;
        e.sync();
    });
}

int skg_grid_tree(skg_engine* eh, int kind, int search, int N, const double* box, const int* child0,
                  const int* parent, const int* cell, const int* dir, const int* nbrStart, const int* nbrIds)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (N < 1 || !box || !child0 || !parent || !cell) throw Error("tree grid tables missing");
        if (kind != 0 && kind != 1) throw Error("tree kind must be 0 (octree) or 1 (binary tree)");
        if (search < 0 || search > 3) throw Error("invalid search method");
        if (search == 2 && kind != 0) throw Error("Bookkeeping method is not compatible with binary tree");   // BinTreeDustGrid.cpp:19-25
        if (search == 1 && (!nbrStart || !nbrIds)) throw Error("Neighbor search needs the neighbour lists");
        if (kind == 1 && !dir) throw Error("binary tree needs the split directions");
        int nchild = kind == 0 ? 8 : 2; int ncells = 0;
        for (int l = 0; l < N; l++)
        {
            if (child0[l] >= 0 && (child0[l] <= l || child0[l] + nchild > N)) throw Error("invalid child index in tree tables");
            if (cell[l] >= 0) ncells++;
            if ((child0[l] < 0) != (cell[l] >= 0)) throw Error("leaf/cell tables are inconsistent");
        }
        if (search == 1)
        {
            if (nbrStart[0] != 0) throw Error("neighbour list offsets must start at 0");
            for (size_t q = 0; q < 6 * (size_t)N; q++) if (nbrStart[q + 1] < nbrStart[q]) throw Error("neighbour list offsets must not decrease");
            const size_t total = (size_t)nbrStart[6 * (size_t)N];
            for (size_t q = 0; q < total; q++) if (nbrIds[q] < 0 || nbrIds[q] >= N) throw Error("invalid neighbour id in tree tables");
        }
        e.freeGrid();
        e.tree.box = up(e, box, 6 * (size_t)N); e.tree.child0 = up(e, child0, N); e.tree.parent = up(e, parent, N);
        e.tree.cell = up(e, cell, N);
        std::vector<int> zeros; if (!dir) { zeros.assign(N, 0); dir = zeros.data(); }
        e.tree.dir = up(e, dir, N);
        // point-location accelerator: a G^3 lattice over the root box; every lattice cell remembers the deepest node
        // whose (closed) box contains the whole cell.  TreeNode::whichnode descends from there instead of from the root
        // when the point lies strictly inside that node's box -- the same leaf, several levels fewer dependent reads.
        {
            int G = N > 4096 ? 64 : (N > 64 ? 8 : 1);
            if (const char* v = getenv("SKG_TREE_LATTICE")) G = std::max(1, std::min(256, atoi(v)));
            std::vector<int> lut((size_t)G * G * G, 0);
            const double* rb = box;
            for (int i = 0; i < G; i++) for (int j = 0; j < G; j++) for (int k = 0; k < G; k++)
            {
                double lo[3] = {rb[0] + (rb[3] - rb[0]) * i / G, rb[1] + (rb[4] - rb[1]) * j / G, rb[2] + (rb[5] - rb[2]) * k / G};
                double hi[3] = {rb[0] + (rb[3] - rb[0]) * (i + 1) / G, rb[1] + (rb[4] - rb[1]) * (j + 1) / G, rb[2] + (rb[5] - rb[2]) * (k + 1) / G};
                int node = 0;
                while (child0[node] >= 0)
                {
                    int next = -1;
                    for (int c = 0; c < nchild; c++)
                    {
                        const double* cb = box + 6 * (size_t)(child0[node] + c);
                        if (lo[0] >= cb[0] && hi[0] <= cb[3] && lo[1] >= cb[1] && hi[1] <= cb[4] && lo[2] >= cb[2] && hi[2] <= cb[5]) { next = child0[node] + c; break; }
                    }
                    if (next < 0) break;
                    node = next;
                }
                lut[((size_t)i * G + j) * G + k] = node;
            }
            e.tree.lookup = up(e, lut.data(), lut.size()); e.tree.lookupG = G;
            for (int c = 0; c < 3; c++) e.tree.lookupInv[c] = G / (rb[3 + c] - rb[c]);
        }
        e.tree.nodeRec = nullptr; e.tree.nbrHint = nullptr;
        if (search == 1)
        {
            const size_t total = (size_t)std::max(0, nbrStart[6 * (size_t)N]);
            e.tree.nbrStart = up(e, nbrStart, 6 * (size_t)N + 1); e.tree.nbrIds = up(e, nbrIds, std::max<size_t>(1, total));
            std::vector<TreeNodeRec> rec(N); std::vector<int> hints;
            for (int l = 0; l < N; l++)
            {
                TreeNodeRec& r = rec[l];
                for (int c = 0; c < 6; c++) r.box[c] = box[6 * (size_t)l + c];
                r.cell = cell[l]; r.hbase = (int)(hints.size() / 4); r.hmeta = 0; r.pad0 = 0; r.pad1[0] = r.pad1[1] = 0;
                for (int w = 0; w < 6; w++)
                {
                    const int beg = nbrStart[6 * (size_t)l + w], cnt = nbrStart[6 * (size_t)l + w + 1] - beg;
                    r.first[w] = cnt > 0 ? nbrIds[beg] : -1;
                    if (cnt > 1)
                    {
                        // which neighbour covers the centre of each of the G x G bins of this wall (in-plane axes a, b); G = 2, 4, 8
                        // or 16, fine enough to resolve the smallest neighbour where that is possible
                        const int a = w < 2 ? 1 : 0, b = w < 4 ? 2 : 1;
                        const double* nb = box + 6 * (size_t)l;
                        double ratio = 1;
                        for (int q = beg; q < beg + cnt; q++)
                        {
                            const double* c = box + 6 * (size_t)nbrIds[q];
                            const double ea = c[a + 3] - c[a], eb = c[b + 3] - c[b];
                            if (ea > 0) ratio = std::max(ratio, (nb[a + 3] - nb[a]) / ea);
                            if (eb > 0) ratio = std::max(ratio, (nb[b + 3] - nb[b]) / eb);
                        }
                        int lg = 0; while (lg < 3 && (2 << lg) < ratio * 0.99) lg++;
                        const int G = 2 << lg;
                        if (hints.size() / 4 >= 2000000000u) throw Error("too many multi-neighbour walls");
                        // does the first neighbour (the lists are sorted by decreasing overlap, TreeNode::sortneighbors) cover at
                        // least half of the wall?  then it is tested first, like a single neighbour
                        const double* c0 = box + 6 * (size_t)nbrIds[beg];
                        const double oa = std::min(nb[a + 3], c0[a + 3]) - std::max(nb[a], c0[a]), ob = std::min(nb[b + 3], c0[b + 3]) - std::max(nb[b], c0[b]);
                        const bool dominant = oa > 0 && ob > 0 && oa * ob >= 0.5 * (nb[a + 3] - nb[a]) * (nb[b + 3] - nb[b]);
                        r.hmeta |= (1u | ((unsigned)lg << 1) | (dominant ? 8u : 0u)) << (5 * w);
                        for (int ia = 0; ia < G; ia++) for (int ib = 0; ib < G; ib++)
                        {
                            const double ca = nb[a] + (nb[a + 3] - nb[a]) * (ia + 0.5) / G, cb = nb[b] + (nb[b + 3] - nb[b]) * (ib + 0.5) / G;
                            int pick = nbrIds[beg];
                            for (int q = beg; q < beg + cnt; q++)
                            {
                                const double* c = box + 6 * (size_t)nbrIds[q];
                                if (ca >= c[a] && ca <= c[a + 3] && cb >= c[b] && cb <= c[b + 3]) { pick = nbrIds[q]; break; }
                            }
                            hints.push_back(pick);
                        }
                        // a wall shared by four finer siblings (the usual level transition of an octree): their ids follow from the
                        // id of the first one and the bin, ids = base + sa*ia + sb*ib -- the walkers then need no table read at all
                        // (bit 4 of the wall's meta; the record's `first` entry of such a wall holds the base)
                        if (G == 2 && !dominant)
                        {
                            const int* h = hints.data() + hints.size() - 4;
                            const int sa = w < 2 ? 2 : 1, sb = w < 4 ? 4 : 2;
                            if (h[2] - h[0] == sa && h[1] - h[0] == sb && h[3] - h[0] == sa + sb) { r.hmeta |= 16u << (5 * w); r.first[w] = h[0]; }
                        }
                    }
                }
            }
            e.tree.nodeRec = up(e, rec.data(), (size_t)N);
            if (hints.empty()) hints.assign(16, 0);
            e.tree.nbrHint = up(e, hints.data(), hints.size());
        }
        else { e.tree.nbrStart = nullptr; e.tree.nbrIds = nullptr; }
        { std::vector<int> cn(std::max(ncells, 1), 0); for (int l = 0; l < N; l++) if (cell[l] >= 0 && cell[l] < ncells) cn[cell[l]] = l; e.tree.cellNode = up(e, cn.data(), cn.size()); }
        e.tree.N = N; e.tree.kind = kind; e.tree.search = search;
        e.tree.eps = epsFor(box[3] - box[0], box[4] - box[1], box[5] - box[2]);
        e.gridKind = GRID_TREE; e.Ncells = ncells;
        e.sync();
    });
}

int skg_grid_amesh(skg_engine* eh, int N, const double* box, const int* nxyz, const int* child0, const int* cell, const int* wallNbr)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (N < 1 || !box || !nxyz || !child0 || !cell || !wallNbr) throw Error("adaptive mesh tables missing");
        int ncells = 0;
        for (int l = 0; l < N; l++)
        {
            if (cell[l] >= 0) ncells++;
            for (int w = 0; w < 6; w++) if (wallNbr[6 * (size_t)l + w] >= N) throw Error("invalid wall neighbour in adaptive mesh tables");
            if (child0[l] >= 0)
            {
                int64_t nc = (int64_t)nxyz[3*l] * nxyz[3*l+1] * nxyz[3*l+2];
                if (nc < 1 || child0[l] <= l || child0[l] + nc > N) throw Error("invalid child index in adaptive mesh tables");
            }
        }
        e.freeGrid();
        e.amesh.box = up(e, box, 6 * (size_t)N); e.amesh.nxyz = up(e, nxyz, 3 * (size_t)N); e.amesh.child0 = up(e, child0, N);
        e.amesh.cell = up(e, cell, N); e.amesh.wallNbr = up(e, wallNbr, 6 * (size_t)N);
        { std::vector<int> cn(std::max(ncells, 1), 0); for (int l = 0; l < N; l++) if (cell[l] >= 0 && cell[l] < ncells) cn[cell[l]] = l; e.amesh.cellNode = up(e, cn.data(), cn.size()); }
        {
            std::vector<AMeshNodeRec> rec(N);
            for (int l = 0; l < N; l++)
            {
                AMeshNodeRec& r = rec[l];
                for (int c = 0; c < 6; c++) { r.box[c] = box[6 * (size_t)l + c]; r.wallNbr[c] = wallNbr[6 * (size_t)l + c]; }
                r.cell = cell[l]; r.child0 = child0[l]; r.nx = nxyz[3 * (size_t)l]; r.ny = nxyz[3 * (size_t)l + 1]; r.nz = nxyz[3 * (size_t)l + 2]; r.parent = 0;
            }
            // parents: the walkers climb from a wall's neighbour to the ancestor a root search would pass through
            for (int l = 0; l < N; l++)
                if (child0[l] >= 0) { const int64_t nc = (int64_t)nxyz[3 * (size_t)l] * nxyz[3 * (size_t)l + 1] * nxyz[3 * (size_t)l + 2]; for (int64_t c = 0; c < nc; c++) rec[child0[l] + c].parent = l; }
            e.amesh.nodeRec = up(e, rec.data(), (size_t)N);
        }
        e.amesh.N = N;
        e.amesh.eps = epsFor(box[3] - box[0], box[4] - box[1], box[5] - box[2]);
        e.gridKind = GRID_AMESH; e.Ncells = ncells;
        e.sync();
    });
}

int skg_grid_voronoi(skg_engine* eh, int N, const double* particles, const int* nbrStart, const int* nbrIds,
                     const double* extent, int nb, const int* blkStart, const int* blkIds, const int* blkTree,
                     int Nkd, const int* kdM, const int* kdAxis, const int* kdUp, const int* kdLeft,
                     const int* kdRight, const double* cellBox)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (N < 1 || !particles || !nbrStart || !nbrIds || !extent || nb < 1 || !blkStart || !blkIds || !blkTree)
            throw Error("voronoi tables missing");
        size_t nb3 = (size_t)nb * nb * nb;
        // every table is validated before the previous grid is released
        if (nbrStart[0] != 0) throw Error("neighbour list offsets must start at 0");
        for (int m = 0; m < N; m++) if (nbrStart[m + 1] < nbrStart[m]) throw Error("neighbour list offsets must not decrease");
        for (int q = 0; q < nbrStart[N]; q++) if (nbrIds[q] >= N || nbrIds[q] < -6) throw Error("invalid neighbour id in Voronoi tables");
        if ((size_t)std::max(0, nbrStart[N]) + N + 8 > 2147483647ull) throw Error("too many Voronoi neighbours for int32 record indices");
        for (size_t b = 0; b < nb3; b++) if (blkStart[b + 1] < blkStart[b] || blkTree[b] >= Nkd) throw Error("invalid block tables in Voronoi grid");
        for (int q = 0; q < blkStart[nb3]; q++) if (blkIds[q] < 0 || blkIds[q] >= N) throw Error("invalid cell id in Voronoi block lists");
        e.freeGrid();
        e.voro.particles = up(e, particles, 3 * (size_t)N);
        e.voro.nbrStart = up(e, nbrStart, (size_t)N + 1); e.voro.nbrIds = up(e, nbrIds, (size_t)std::max(1, nbrStart[N]));
        e.voro.blkStart = up(e, blkStart, nb3 + 1); e.voro.blkIds = up(e, blkIds, (size_t)std::max(1, blkStart[nb3]));
        e.voro.blkTree = up(e, blkTree, nb3);
        int one = -1; size_t nk = (size_t)std::max(1, Nkd);
        e.voro.kdM = up(e, Nkd ? kdM : &one, nk); e.voro.kdAxis = up(e, Nkd ? kdAxis : &one, nk); e.voro.kdUp = up(e, Nkd ? kdUp : &one, nk);
        e.voro.kdLeft = up(e, Nkd ? kdLeft : &one, nk); e.voro.kdRight = up(e, Nkd ? kdRight : &one, nk);
        e.voro.cellBox = cellBox ? up(e, cellBox, 6 * (size_t)N) : nullptr;
        {
            const size_t total = (size_t)std::max(0, nbrStart[N]);
            std::vector<double> rec(4 * (total + N + 8), 0.0);
            auto pack = [](int lo, int hi) { const long long v = (long long)(unsigned)lo | ((long long)hi << 32); double d; std::memcpy(&d, &v, 8); return d; };
            for (int m = 0; m < N; m++)
            {
                const int beg = nbrStart[m], cnt = nbrStart[m + 1] - beg;
                double* h = rec.data() + 4 * ((size_t)beg + m);
                h[0] = particles[3 * (size_t)m]; h[1] = particles[3 * (size_t)m + 1]; h[2] = particles[3 * (size_t)m + 2]; h[3] = pack(cnt, 0);
                for (int q = 0; q < cnt; q++)
                {
                    const int id = nbrIds[beg + q];
                    double* s4 = h + 4 * (size_t)(q + 1);
                    if (id >= 0) { s4[0] = particles[3 * (size_t)id]; s4[1] = particles[3 * (size_t)id + 1]; s4[2] = particles[3 * (size_t)id + 2]; s4[3] = pack(id, nbrStart[id] + id); }
                    else s4[3] = pack(id, 0);
                }
            }
            e.voro.rec = up(e, rec.data(), rec.size());
        }
        e.voro.planes = nullptr;
        // plane records of the shooting stages (tables.h): the layout of the crossing records with normals in place of positions.
        // Built where the table stays resident in L2 (126 MB): there the neighbour loop is bound by instructions and the
        // planes' shorter loop pays (+31 % at 200 000 cells); a mesh whose records come from DRAM at every crossing measured
        // 8 % slower with them (1e6 cells) and keeps walking the crossing records.  SKG_VORO_PLANES=0/1 overrides the rule.
        bool wantPlanes = 32.0 * ((double)std::max(0, nbrStart[N]) + N) <= 128.0 * 1024 * 1024;
        if (const char* v = getenv("SKG_VORO_PLANES")) wantPlanes = atoi(v) != 0;
        if (wantPlanes)
        {
            const size_t total = (size_t)std::max(0, nbrStart[N]);
            std::vector<double> pl(4 * (total + N + 8), 0.0);
            auto pack = [](int lo, int hi) { const long long v = (long long)(unsigned)lo | ((long long)hi << 32); double d; std::memcpy(&d, &v, 8); return d; };
            const double lo[3] = {extent[0], extent[2], extent[4]}, hi[3] = {extent[1], extent[3], extent[5]};
            for (int m = 0; m < N; m++)
            {
                const int beg = nbrStart[m], cnt = nbrStart[m + 1] - beg;
                double* h = pl.data() + 4 * ((size_t)beg + m);
                const double* pr = particles + 3 * (size_t)m;
                h[0] = pr[0]; h[1] = pr[1]; h[2] = pr[2]; h[3] = pack(cnt, 0);
                for (int q = 0; q < cnt; q++)
                {
                    const int id = nbrIds[beg + q];
                    double* s4 = h + 4 * (size_t)(q + 1);
                    if (id >= 0)
                    {
                        // the bisector plane of p and p_i: n = p_i - p, n.(r - p) = |n|^2 / 2
                        const double* pi = particles + 3 * (size_t)id;
                        for (int c = 0; c < 3; c++) s4[c] = pi[c] - pr[c];
                        s4[3] = pack(id, nbrStart[id] + id);
                    }
                    else
                    {
                        // walls -1 .. -6: xmin, xmax, ymin, ymax, zmin, zmax (VoronoiMesh.cpp:800-812) at distance D along the axis: the
                        // bisector plane towards the particle's mirror image, n = 2 D along the outward normal
                        const int a = (-1 - id) >> 1; const bool upper = ((-1 - id) & 1) != 0;
                        s4[a] = upper ? 2.0 * (hi[a] - pr[a]) : -2.0 * (pr[a] - lo[a]);
                        s4[3] = pack(id, 0);
                    }
                }
            }
            e.voro.planes = up(e, pl.data(), pl.size());
        }
        // extent arrives as xmin,xmax,ymin,ymax,zmin,zmax (Box setters order); stored as min corner, max corner
        e.voro.ext[0] = extent[0]; e.voro.ext[1] = extent[2]; e.voro.ext[2] = extent[4];
        e.voro.ext[3] = extent[1]; e.voro.ext[4] = extent[3]; e.voro.ext[5] = extent[5];
        e.voro.eps = epsFor(extent[1] - extent[0], extent[3] - extent[2], extent[5] - extent[4]);
        e.voro.N = N; e.voro.nb = nb;
        e.gridKind = GRID_VORO; e.Ncells = N;
        e.sync();
    });
}

int skg_medium(skg_engine* eh, int Ncells, int Ncomp, int Nlambda, const double* rho, const double* kext, const double* ksca, const double* g)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (Ncells < 1 || Ncomp < 1 || Nlambda < 1 || !rho || !kext) throw Error("medium tables missing");
        if (e.gridKind != GRID_NONE && Ncells != e.Ncells) throw Error("medium has " + std::to_string(Ncells) + " cells but the grid has " + std::to_string(e.Ncells));
        // the density table is read in aligned 32-byte sectors by the shooting stages (RhoSector): one spare sector at the end
        e.rho.ensure(sizeof(double) * (size_t)Ncells * Ncomp + 64);
        SKG_CUDA(cudaMemsetAsync(static_cast<char*>(e.rho.p) + sizeof(double) * (size_t)Ncells * Ncomp, 0, 64, e.stream));
        e.rho.upload(rho, sizeof(double) * (size_t)Ncells * Ncomp, e.stream);
        e.kext.upload(kext, sizeof(double) * (size_t)Ncomp * Nlambda, e.stream);
        std::vector<double> zeros((size_t)Ncomp * Nlambda, 0.0);
        e.ksca.upload(ksca ? ksca : zeros.data(), sizeof(double) * (size_t)Ncomp * Nlambda, e.stream);
        e.gasym.upload(g ? g : zeros.data(), sizeof(double) * (size_t)Ncomp * Nlambda, e.stream);
        e.med.rho = e.rho.as<double>(); e.med.kext = e.kext.as<double>(); e.med.ksca = e.ksca.as<double>(); e.med.g = e.gasym.as<double>();
        e.med.Ncells = Ncells; e.med.Ncomp = Ncomp; e.med.Nlambda = Nlambda;
        e.med.Ntheta = 0; e.med.S11 = e.med.S12 = e.med.S33 = e.med.S34 = e.med.thetaX = e.med.pfnorm = nullptr;    // a new medium starts unpolarised
        e.haveDustLib = false;       // the dust library tables belong to the previous medium
        e.sync();
    });
}

// ---- deterministic geometry ---------------------------------------------------------------------------
int skg_path_count(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, int64_t* offsets, int64_t* total)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (n < 0 || (n > 0 && (!r || !k)) || !offsets) throw Error("skg_path_count: bad arguments");
        const double* d_r = r; const double* d_k = k; int64_t* d_off = offsets;
        if (mem == SKG_HOST)
        {
            e.scratchR.upload(r, sizeof(double) * 3 * (size_t)n, e.stream); e.scratchK.upload(k, sizeof(double) * 3 * (size_t)n, e.stream);
            e.scratchOffsets.ensure(sizeof(int64_t) * ((size_t)n + 1));
            d_r = e.scratchR.as<double>(); d_k = e.scratchK.as<double>(); d_off = e.scratchOffsets.as<int64_t>();
        }
        e.scratchCounts.ensure(sizeof(int) * (size_t)std::max<int64_t>(n, 1));
        launchPathCount(e, n, d_r, d_k, e.scratchCounts.as<int>());
        exclusiveScan(e, n, e.scratchCounts.as<int>(), d_off);
        int64_t tot = 0;
        if (mem == SKG_HOST) SKG_CUDA(cudaMemcpyAsync(offsets, d_off, sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, e.stream));
        if (total || mem == SKG_HOST) SKG_CUDA(cudaMemcpyAsync(&tot, d_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, e.stream));
        e.sync();
        if (total) *total = tot;
    });
}

int skg_path_fill(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, const int* ell, int ellStride,
                  const int64_t* offsets, skg_segment* segments)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (n < 0 || (n > 0 && (!r || !k)) || !offsets) throw Error("skg_path_fill: bad arguments");
        if (ellStride != 0 && ellStride != 1) throw Error("ell_stride must be 0 or 1");
        if (mem == SKG_DEVICE)
        {
            if (!segments) throw Error("skg_path_fill: null output");
            if (reinterpret_cast<uintptr_t>(segments) & 31) throw Error("skg_path_fill: the segment array must be 32-byte aligned");
            launchPathFill(e, n, r, k, ell, ellStride, offsets, segments);
            e.sync();
            return;
        }
        int64_t total = n > 0 ? offsets[n] : 0;
        if (total > 0 && !segments) throw Error("skg_path_fill: null output");
        e.scratchR.upload(r, sizeof(double) * 3 * (size_t)n, e.stream); e.scratchK.upload(k, sizeof(double) * 3 * (size_t)n, e.stream);
        e.scratchOffsets.upload(offsets, sizeof(int64_t) * ((size_t)n + 1), e.stream);
        const int* d_ell = nullptr;
        if (ell) for (int64_t i = 0; i < (ellStride ? n : 1); i++) if (ell[i] < 0 || ell[i] >= e.med.Nlambda) throw Error("wavelength index out of range");
        if (ell) { e.scratchEll.upload(ell, sizeof(int) * (size_t)(ellStride ? n : 1), e.stream); d_ell = e.scratchEll.as<int>(); }
        size_t t = (size_t)std::max<int64_t>(total, 1);
        e.scratchOut[0].ensure(sizeof(skg_segment) * t);
        launchPathFill(e, n, e.scratchR.as<double>(), e.scratchK.as<double>(), d_ell, ellStride, e.scratchOffsets.as<int64_t>(),
                       e.scratchOut[0].as<skg_segment>());
        if (total > 0) SKG_CUDA(cudaMemcpyAsync(segments, e.scratchOut[0].p, sizeof(skg_segment) * total, cudaMemcpyDeviceToHost, e.stream));
        e.sync();
    });
}

static int opticalDepthImpl(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, const int* ell, int ellStride,
                            const double* distance, double* tau, bool mcWalker)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (n < 0 || (n > 0 && (!r || !k || !tau || !ell))) throw Error("skg_opticaldepth: bad arguments");
        if (ellStride != 0 && ellStride != 1) throw Error("ell_stride must be 0 or 1");
        if (mem == SKG_DEVICE) { launchOpticalDepth(e, n, r, k, ell, ellStride, distance, tau, mcWalker); e.sync(); return; }
        for (int64_t i = 0; i < (ellStride ? n : 1); i++) if (ell[i] < 0 || ell[i] >= e.med.Nlambda) throw Error("wavelength index out of range");
        e.scratchR.upload(r, sizeof(double) * 3 * (size_t)n, e.stream); e.scratchK.upload(k, sizeof(double) * 3 * (size_t)n, e.stream);
        e.scratchEll.upload(ell, sizeof(int) * (size_t)(ellStride ? n : 1), e.stream);
        const double* d_dist = nullptr;
        if (distance) { e.scratchDist.upload(distance, sizeof(double) * (size_t)n, e.stream); d_dist = e.scratchDist.as<double>(); }
        e.scratchTau.ensure(sizeof(double) * (size_t)std::max<int64_t>(n, 1));
        launchOpticalDepth(e, n, e.scratchR.as<double>(), e.scratchK.as<double>(), e.scratchEll.as<int>(), ellStride, d_dist, e.scratchTau.as<double>(), mcWalker);
        if (n > 0) SKG_CUDA(cudaMemcpyAsync(tau, e.scratchTau.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
        e.sync();
    });
}

// one-pass batched path(): slabs from analytic capacities (Cartesian) or exact counts (other grids), then the record kernel
int skg_path_batch(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, const int* ell, int ellStride,
                   int64_t* starts, int32_t* lengths, skg_segment* segments, int64_t capacity, int64_t* needed)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (n < 0 || (n > 0 && (!r || !k)) || !starts || !lengths || !needed) throw Error("skg_path_batch: bad arguments");
        if (ellStride != 0 && ellStride != 1) throw Error("ell_stride must be 0 or 1");
        const bool host = mem == SKG_HOST;
        const double* d_r = r; const double* d_k = k; const int* d_ell = ell; int64_t* d_starts = starts; int* d_len = lengths;
        if (host)
        {
            e.scratchR.upload(r, sizeof(double) * 3 * (size_t)n, e.stream); e.scratchK.upload(k, sizeof(double) * 3 * (size_t)n, e.stream);
            d_r = e.scratchR.as<double>(); d_k = e.scratchK.as<double>();
            if (ell)
            {
                for (int64_t i = 0; i < (ellStride ? n : 1); i++) if (ell[i] < 0 || ell[i] >= e.med.Nlambda) throw Error("wavelength index out of range");
                e.scratchEll.upload(ell, sizeof(int) * (size_t)(ellStride ? n : 1), e.stream); d_ell = e.scratchEll.as<int>();
            }
            e.scratchOffsets.ensure(sizeof(int64_t) * ((size_t)n + 1)); d_starts = e.scratchOffsets.as<int64_t>();
            e.scratchM.ensure(sizeof(int) * (size_t)std::max<int64_t>(n, 1)); d_len = e.scratchM.as<int>();
        }
        e.scratchCounts.ensure(sizeof(int) * (size_t)std::max<int64_t>(n, 1));
        const bool analytic = e.gridKind == GRID_CART;
        if (analytic) launchPathCapacity(e, n, d_r, d_k, e.scratchCounts.as<int>());
        else launchPathCount(e, n, d_r, d_k, e.scratchCounts.as<int>());
        exclusiveScan(e, n, e.scratchCounts.as<int>(), d_starts);
        int64_t total = 0;
        SKG_CUDA(cudaMemcpyAsync(&total, d_starts + n, sizeof(int64_t), cudaMemcpyDeviceToHost, e.stream));
        e.sync();
        *needed = total + 8;                    // a few spare records behind the last slab
        if (capacity < *needed || (!segments && total > 0))
        { if (segments) throw Error("skg_path_batch: the segment array holds " + std::to_string(capacity) + " records, " + std::to_string(*needed) + " are needed"); return; }
        skg_segment* d_seg = segments;
        if (host) { e.scratchOut[0].ensure(sizeof(skg_segment) * (size_t)*needed); d_seg = e.scratchOut[0].as<skg_segment>(); }
        else if (reinterpret_cast<uintptr_t>(segments) & 31) throw Error("skg_path_batch: the segment array must be 32-byte aligned");
        const unsigned long long errorsBefore = e.readCounters().errors;
        launchPathFill(e, n, d_r, d_k, d_ell, ellStride, d_starts, d_seg, d_len);
        if (host)
        {
            SKG_CUDA(cudaMemcpyAsync(starts, d_starts, sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, e.stream));
            if (n > 0) SKG_CUDA(cudaMemcpyAsync(lengths, d_len, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
            if (total > 0) SKG_CUDA(cudaMemcpyAsync(segments, d_seg, sizeof(skg_segment) * (size_t)total, cudaMemcpyDeviceToHost, e.stream));
        }
        if (e.readCounters().errors != errorsBefore) throw Error("skg_path_batch: a path was longer than its slab (the analytic capacity was too small); use skg_path_count + skg_path_fill");
    });
}

int skg_opticaldepth(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, const int* ell, int ellStride,
                     const double* distance, double* tau)
{ return opticalDepthImpl(eh, mem, n, r, k, ell, ellStride, distance, tau, false); }
int skg_opticaldepth_mc(skg_engine* eh, int mem, int64_t n, const double* r, const double* k, const int* ell, int ellStride,
                        const double* distance, double* tau)
{ return opticalDepthImpl(eh, mem, n, r, k, ell, ellStride, distance, tau, true); }

int skg_whichcell(skg_engine* eh, int mem, int64_t n, const double* r, int* m)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (n < 0 || (n > 0 && (!r || !m))) throw Error("skg_whichcell: bad arguments");
        if (mem == SKG_DEVICE) { launchWhichCell(e, n, r, m); e.sync(); return; }
        e.scratchR.upload(r, sizeof(double) * 3 * (size_t)n, e.stream);
        e.scratchM.ensure(sizeof(int) * (size_t)std::max<int64_t>(n, 1));
        launchWhichCell(e, n, e.scratchR.as<double>(), e.scratchM.as<int>());
        if (n > 0) SKG_CUDA(cudaMemcpyAsync(m, e.scratchM.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e.stream));
        e.sync();
    });
}

int skg_selftest_division(skg_engine* eh, uint64_t n, uint64_t seed, uint64_t* mismatches)
{ return guarded([&]{ if (!mismatches) throw Error("null output"); *mismatches = runDivisionSelfTest(E(eh), n, seed); }); }

int skg_selftest_atomics(skg_engine* eh, uint64_t n, int cells, double* atomicsPerSecond)
{ return guarded([&]{ if (!atomicsPerSecond) throw Error("null output"); *atomicsPerSecond = mcAtomicRate(E(eh), n, cells); }); }

int skg_stuck_counts(skg_engine* eh, int64_t* escaped, int64_t* terminated)
{
    return guarded([&]{
        Engine& e = E(eh);
        Counters c = e.readCounters();
        if (escaped) *escaped = (int64_t)c.stuckEscaped;
        if (terminated) *terminated = (int64_t)c.stuckTerminated;
    });
}

// ---- Monte Carlo -----------------------------------------------------------------------------------------
int skg_sources(skg_engine* eh, int Ncomp, const skg_source* comps, int Nlambda, const double* L, double emissionBias)
{ return guarded([&]{ mcSetSources(E(eh), Ncomp, comps, Nlambda, L, emissionBias); }); }
int skg_medium_polarization(skg_engine* eh, int Ntheta, const double* S11, const double* S12, const double* S33, const double* S34)
{ return guarded([&]{ mcSetPolarization(E(eh), Ntheta, S11, S12, S33, S34); }); }
int skg_instruments(skg_engine* eh, int n, const skg_instrument* instr)
{ return guarded([&]{ mcSetInstruments(E(eh), n, instr); }); }
int skg_run_stellar(skg_engine* eh, const skg_mc_params* p, skg_mc_stats* stats)
{ return guarded([&]{ if (!p) throw Error("null parameters"); mcRunStellar(E(eh), *p, stats); }); }
int skg_run_dust(skg_engine* eh, const skg_mc_params* p, int phase, double emissionBias, int mem, const double* Lcell, skg_mc_stats* stats)
{ return guarded([&]{ if (!p) throw Error("null parameters"); mcRunDust(E(eh), *p, phase, emissionBias, mem, Lcell, stats); }); }
int skg_sample_launch(skg_engine* eh, int ell, int n, uint64_t seed, double* r, double* k, double* L)
{ return guarded([&]{ mcSampleLaunch(E(eh), ell, n, seed, r, k, L); }); }
int skg_sample_density(skg_engine* eh, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* rho)
{ return guarded([&]{ mcSampleDensity(E(eh), Ncomp, geoms, norm, sampleCount, seed, rho); }); }
int skg_sample_boxes(skg_engine* eh, int64_t n, const double* box, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount, uint64_t seed, double* mass)
{ return guarded([&]{ mcSampleBoxes(E(eh), n, box, Ncomp, geoms, norm, sampleCount, seed, mass, nullptr); }); }
int skg_sample_boxes_dispersion(skg_engine* eh, int64_t n, const double* box, int Ncomp, const skg_source* geoms, const double* norm, int sampleCount,
                                uint64_t seed, double* mass, double* dispersion)
{ return guarded([&]{ if (n > 0 && !dispersion) throw Error("skg_sample_boxes_dispersion: bad arguments");
                      mcSampleBoxes(E(eh), n, box, Ncomp, geoms, norm, sampleCount, seed, mass, dispersion); }); }
int skg_reset_results(skg_engine* eh) { return guarded([&]{ mcResetResults(E(eh)); }); }
int skg_dust_library(skg_engine* eh, const double* volumes, const double* kappaabs, const double* lambda, const double* dlambda)
{ return guarded([&]{ if (!volumes || !kappaabs || !lambda || !dlambda) throw Error("skg_dust_library: null table"); mcDustLibrary(E(eh), volumes, kappaabs, lambda, dlambda); }); }
int skg_dust_cell_luminosities(skg_engine* eh, double** d_Lcell)
{ return guarded([&]{ if (!d_Lcell) throw Error("null output"); *d_Lcell = mcDustCellLuminosities(E(eh)); }); }
int skg_reset_labs_dust(skg_engine* eh)
{ return guarded([&]{ Engine& e = E(eh); if (e.labsDust.p && e.labsCount) SKG_CUDA(cudaMemsetAsync(e.labsDust.p, 0, sizeof(double) * e.labsCount, e.stream)); e.accLabsDust = Engine::ACC_ZERO; e.sync(); }); }
int skg_fetch_labs_dust(skg_engine* eh, double* labs, int add)
{ return guarded([&]{ if (!labs) throw Error("null host array"); mcFetchLabs(E(eh), labs, add, 1); }); }
int skg_labs_bolometric(skg_engine* eh, double* Labsbol)
{ return guarded([&]{ if (!Labsbol) throw Error("null host array"); mcLabsBolometric(E(eh), Labsbol); }); }

static void fetchArray(Engine& e, const double* d_src, int64_t count, double* host, int add)
{
    if (!host) throw Error("null host array");
    if (!add) { SKG_CUDA(cudaMemcpyAsync(host, d_src, sizeof(double) * count, cudaMemcpyDeviceToHost, e.stream)); e.sync(); return; }
    std::vector<double> tmp(count);
    SKG_CUDA(cudaMemcpyAsync(tmp.data(), d_src, sizeof(double) * count, cudaMemcpyDeviceToHost, e.stream)); e.sync();
    for (int64_t i = 0; i < count; i++) host[i] += tmp[i];
}

int skg_fetch_frame(skg_engine* eh, int i, double* frame, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (i < 0 || i >= (int)e.instr.size() || !e.instr[i].frame) throw Error("instrument has no frame");
        fetchArray(e, e.instr[i].frame, (int64_t)e.instr[i].frameCount, frame, add);      // (a MultiFrameInstrument: all its slabs; see skg_fetch_multiframe)
    });
}
int skg_fetch_sed(skg_engine* eh, int i, double* sed, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (i < 0 || i >= (int)e.instr.size() || !e.instr[i].sed) throw Error("instrument has no SED");
        fetchArray(e, e.instr[i].sed, e.instrNlambda, sed, add);
    });
}
int skg_fetch_multiframe(skg_engine* eh, int i, int which, int ell, double* frame, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (i < 0 || i >= (int)e.instr.size() || e.instr[i].kind != SKG_INSTR_MULTIFRAME) throw Error("not a multi-frame instrument");
        const InstrDev& d = e.instr[i];
        if (ell < 0 || ell >= e.instrNlambda) throw Error("wavelength index out of range");
        const int slab = which < 0 ? d.mfTotal : (which < d.mfNcomp && d.mfComp0 >= 0 ? d.mfComp0 + which : -1);
        if (slab < 0) throw Error(which < 0 ? "the instrument does not record the total flux" : "the instrument does not record this stellar component");
        FrameDev f;
        SKG_CUDA(cudaMemcpyAsync(&f, d.frames + ell, sizeof(FrameDev), cudaMemcpyDeviceToHost, e.stream)); e.sync();
        fetchArray(e, d.frame + (size_t)slab * d.mfPixels + f.offset, (int64_t)f.Nxp * f.Nyp, frame, add);
    });
}
int skg_fetch_frame_channel(skg_engine* eh, int i, int c, double* frame, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (i < 0 || i >= (int)e.instr.size() || !e.instr[i].chanFrame) throw Error("instrument has no channels");
        if (c < 0 || c >= e.instr[i].Nchan) throw Error("channel out of range");
        const int64_t n = (int64_t)e.instr[i].Nxp * e.instr[i].Nyp * e.instrNlambda;
        fetchArray(e, e.instr[i].chanFrame + c * n, n, frame, add);
    });
}
int skg_fetch_sed_channel(skg_engine* eh, int i, int c, double* sed, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (i < 0 || i >= (int)e.instr.size() || !e.instr[i].chanSed) throw Error("instrument has no channels");
        if (c < 0 || c >= e.instr[i].Nchan) throw Error("channel out of range");
        fetchArray(e, e.instr[i].chanSed + (int64_t)c * e.instrNlambda, e.instrNlambda, sed, add);
    });
}
int skg_fetch_labs(skg_engine* eh, double* labs, int add)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (!labs) throw Error("null host array");
        mcFetchLabs(e, labs, add, 0);
    });
}
int skg_device_accumulators(skg_engine* eh, int which, int part, double** d_ptr, int64_t* count)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (!d_ptr || !count) throw Error("null output");
        if (which == 0) { *d_ptr = e.labs.as<double>(); *count = e.labsCount; return; }
        if (which == -1) { *d_ptr = e.labsDust.as<double>(); *count = e.labsDust.p ? e.labsCount : 0; return; }
        int i = which - 1;
        if (i < 0 || i >= (int)e.instr.size()) throw Error("instrument index out of range");
        const InstrDev& d = e.instr[i];
        if (d.kind == SKG_INSTR_FULL)
        {
            if (part == 0) { *d_ptr = d.chanFrame; *count = (int64_t)d.Nxp * d.Nyp * e.instrNlambda * d.Nchan; }
            else { *d_ptr = d.chanSed; *count = (int64_t)e.instrNlambda * d.Nchan; }
            return;
        }
        if (part == 0) { *d_ptr = d.frame; *count = d.frame ? (int64_t)d.frameCount : 0; }
        else { *d_ptr = d.sed; *count = d.sed ? e.instrNlambda : 0; }
    });
}

// ---- results to the host while the next phase runs ------------------------------------------------------------------------
// skg_results_snapshot copies every accumulator into a shadow array on the engine's stream (the absorption tables straight
// into the reference's (m, ell) layout); skg_fetch_snapshot_async moves a shadow to (page-locked) host memory on a second
// stream, so that the transfer overlaps whatever the engine does next; skg_fetch_snapshot_wait completes the transfers.
int skg_results_snapshot(skg_engine* eh)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (!e.copyStream) SKG_CUDA(cudaStreamCreateWithFlags(&e.copyStream, cudaStreamNonBlocking));
        if (!e.snapEvent) SKG_CUDA(cudaEventCreateWithFlags(&e.snapEvent, cudaEventDisableTiming));
        SKG_CUDA(cudaStreamSynchronize(e.copyStream));          // transfers of the previous snapshot must not be overwritten
        size_t used = 0;
        auto shadow = [&](int which, int part, int64_t count) -> DevBuf*
        {
            if (used == e.shadows.size()) e.shadows.push_back({which, part, count, new DevBuf()});
            Engine::Shadow& sh = e.shadows[used++]; sh.which = which; sh.part = part; sh.count = count;
            sh.buf->ensure(sizeof(double) * (size_t)std::max<int64_t>(count, 1));
            return sh.buf;
        };
        for (int which : {0, -1})
        {
            DevBuf& src = which ? e.labsDust : e.labs;
            if (!src.p || !e.labsCount) continue;
            DevBuf* dst = shadow(which, 0, e.labsCount);
            mcTransposeLabs(e, src.as<double>(), dst->as<double>());
        }
        for (size_t i = 0; i < e.instr.size(); i++)
        {
            const InstrDev& d = e.instr[i]; const int64_t Nl = e.instrNlambda, Nf = (int64_t)d.Nxp * d.Nyp;
            const double* srcs[2] = {d.kind == SKG_INSTR_FULL ? d.chanFrame : d.frame, d.kind == SKG_INSTR_FULL ? d.chanSed : d.sed};
            const int64_t counts[2] = {d.kind == SKG_INSTR_FULL ? Nf * Nl * d.Nchan : (int64_t)d.frameCount, Nl * (d.kind == SKG_INSTR_FULL ? d.Nchan : 1)};
            for (int part = 0; part < 2; part++)
                if (srcs[part]) SKG_CUDA(cudaMemcpyAsync(shadow((int)i + 1, part, counts[part])->p, srcs[part], sizeof(double) * counts[part], cudaMemcpyDeviceToDevice, e.stream));
        }
        for (size_t q = used; q < e.shadows.size(); q++) e.shadows[q].count = 0;
        SKG_CUDA(cudaEventRecord(e.snapEvent, e.stream));
    });
}
int skg_fetch_snapshot_async(skg_engine* eh, int which, int part, double* host, int64_t* count)
{
    return guarded([&]{
        Engine& e = E(eh);
        if (!e.snapEvent) throw Error("skg_results_snapshot has not been called");
        for (const Engine::Shadow& sh : e.shadows)
            if (sh.which == which && sh.part == part && sh.count > 0)
            {
                if (count) *count = sh.count;
                if (!host) return;
                SKG_CUDA(cudaStreamWaitEvent(e.copyStream, e.snapEvent, 0));
                SKG_CUDA(cudaMemcpyAsync(host, sh.buf->p, sizeof(double) * (size_t)sh.count, cudaMemcpyDeviceToHost, e.copyStream));
                return;
            }
        throw Error("no such accumulator in the snapshot");
    });
}
int skg_fetch_snapshot_wait(skg_engine* eh)
{ return guarded([&]{ Engine& e = E(eh); if (e.copyStream) SKG_CUDA(cudaStreamSynchronize(e.copyStream)); }); }

}   // extern "C"
