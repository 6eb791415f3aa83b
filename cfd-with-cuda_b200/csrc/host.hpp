// Host side of the solver (product code): mesh input, node generation, sparsity, step0 assembly,
// output.  Re-designed from the reference's serial O(NE*26*12) / O(NN*216) preprocessing
// (blascoCodinaHuerta.cpp:732-2159) as hashed / sort-based O(NE) passes with OpenMP, producing the
// same numbering and the same CSR arrays.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "common.hpp"

namespace cfdb {

struct Mesh {
   // .inp header (blascoCodinaHuerta.cpp:545-562)
   int eType = 1, NE = 0, NCN = 0, NENv = 27, NENp = 8, NGP = 8;
   double alpha = 1, dt = 0, t_ini = 0, t_final = 0;
   int maxIter = 1;
   double tolerance = 0, converge = 0;
   int isRestart = 0;
   double density = 1, viscosity = 0, fx = 0, fy = 0;
   std::vector<double> coord;   // [node][3]; NCN rows after reading, NN rows after generate_midnodes
   std::vector<int> LtoGnode;   // [e][NENv]
   int nBC = 0;
   std::vector<double> BCtype, BCstr;   // BCstr [nBC][3]
   std::vector<int> velFaces, outFaces; // [i][3] = elem, face, bc (0-based)
   int zeroPressureNode = -1;
   double monCoord[3] = {0, 0, 0};
   // derived
   int NN = 0, NNp = 0;
   std::vector<int> BCvelNodes;   // [i][2]
   int BCnVelNodes = 0;
   int monPoint = 0;
   // Partitioned (multi-GPU) sub-meshes number their pressure DOFs separately from the nodes (owned first):
   // presIndex[node] = pressure DOF of a corner node, -1 otherwise.  Empty = identity (the reference's convention).
   std::vector<int> presIndex;
   std::vector<int> meshColors, NmeshColors, elementsOfColor;   // setupMeshColoring (kept for API parity; kernels use atomics)
   int nActiveColors = 0;
};

// std::vector whose resize() leaves new elements uninitialised, so that multi-GB arrays can be first-touched
// (zeroed) in parallel instead of serially inside the constructor.
template <class T>
struct default_init_alloc : std::allocator<T> {
   template <class U> struct rebind { typedef default_init_alloc<U> other; };
   template <class U> void construct(U* p) { ::new ((void*)p) U; }
   template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
};
typedef std::vector<double, default_init_alloc<double> > big_dvec;
inline void par_zero(big_dvec& v, size_t n) {
   v.resize(n);
   double* p = v.data();
#pragma omp parallel for schedule(static)
   for (long long i = 0; i < (long long)n; i++) p[i] = 0.0;
}

struct CsrPattern {
   int nrows = 0, ncols = 0;
   std::vector<int> rowStarts, col;
   long long nnz() const { return (long long)col.size(); }
};

struct Assembled {
   // Q1/Q1 on a whole mesh: M, G and the FE-Laplacian Z have one and the same pattern and element map; they are then built
   // once and shared (at 256^3 one pattern + map is 6 GB of host memory and seconds of copying)
   CsrPattern Mp, GpOwn, ZpOwn;
   std::vector<int> mapM, mapGOwn, mapZOwn;   // [e][i][j] CSR positions
   bool sharedG = false, sharedZ = false;
   CsrPattern& Gp() { return sharedG ? Mp : GpOwn; }
   CsrPattern& Zp() { return sharedZ ? Mp : ZpOwn; }
   const CsrPattern& Gp() const { return sharedG ? Mp : GpOwn; }
   const CsrPattern& Zp() const { return sharedZ ? Mp : ZpOwn; }
   const std::vector<int>& mapG() const { return sharedG ? mapM : mapGOwn; }
   const std::vector<int>& mapZ() const { return sharedZ ? mapM : mapZOwn; }
   big_dvec Mval, Kval, G1, G2, G3, Zval;
   std::vector<double> MdInv, MdOrigInv;
};

bool read_inp(const std::string& path, Mesh& m);
void make_cavity(int nx, int ny, int nz, double cluster, int NENv, Mesh& m);
void generate_midnodes(Mesh& m);
void find_bc_nodes(Mesh& m);
void find_monitor(Mesh& m);
void color_mesh(Mesh& m);
void build_pattern(const Mesh& m, int nRowLoc, int nColLoc, int nrows, int ncols, CsrPattern& P, std::vector<int>* map,
                   const std::vector<int>* rowIndex = nullptr, const std::vector<int>* colIndex = nullptr);
void assemble_step0(const Mesh& m, int scheme, Assembled& A, bool patternsOnly = false);
bool write_tecplot(const Mesh& m, const double* U, const double* P, const std::string& path, const std::string& title);
bool read_restart(const Mesh& m, const std::string& path, double* U, double* P);

// ---- element-block partition for multi-GPU runs -------------------------------------
struct PeerPlan {
   int rank = 0;
   std::vector<int> sendV, sendP;   // my OWNED local velocity-node / pressure-DOF indices the peer holds as ghosts
   int recvVoff = 0, recvVcnt = 0;  // my ghost range [off, off+cnt) filled by this peer (velocity nodes)
   int recvPoff = 0, recvPcnt = 0;  // same for pressure DOFs
   int remoteVoff = 0, remotePoff = 0;   // where MY values start inside the PEER's ghost ranges (its recvVoff / recvPoff for me)
};
struct Partition {
   int rank = 0, world = 1;
   int NEa = 0;                     // local elements that touch an owned node (the EBE kernels loop over these)
   int NNo = 0, NPo = 0;            // owned velocity nodes / pressure DOFs (numbered first)
   std::vector<int> elemGlobal;     // local element -> global element
   std::vector<int> nodeGlobal;     // local node -> global node
   std::vector<int> presGlobal;     // local pressure DOF -> global corner node
   std::vector<PeerPlan> peers;
   int pinLocal = -1;               // local pressure DOF of the pinned node (reference rule: global id > 0), -1 if absent
   int monLocal = -1;               // local node of the monitor point if owned here, else -1
};
// Extracts rank's sub-mesh (own elements + `layers` halo layers) from the global mesh with generated nodes.
void partition_mesh(const Mesh& global, int world, int rank, int layers, Mesh& local, Partition& part);

}   // namespace cfdb
