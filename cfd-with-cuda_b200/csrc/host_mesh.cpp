// Mesh input and topology preprocessing (product host code).
//
// Replaces readInputFile / findElemsOfPresNodes / findElemNeighbors / setupMeshColoring /
// setupNonCornerNodes / determineVelBCnodes / findMonitorPoint / setupSparseM / setupSparseG
// of fractionalStep/explicit/Cpp/blascoCodinaHuerta.cpp:528-2159 (identical in
// guermondQuartapelle.cpp:537-2647).  Same results (node numbering, sorted CSR columns, element ->
// CSR maps), different algorithms: one-pass text cursor instead of iostream extraction, hashing on
// corner ids instead of coordinate matching against 26 neighbours, counting-sort adjacency +
// per-row sort/unique instead of the NN x 216 NZcolsInARow table, OpenMP throughout.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <unordered_map>

#include "host.hpp"

namespace cfdb {

namespace {

// local corner ids of edges / faces, and face -> non-corner nodes
const int kEdgeCorner[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
const int kFaceCorner[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {1, 2, 5, 6}, {2, 3, 6, 7}, {0, 3, 4, 7}, {4, 5, 6, 7}};
const int kFaceInner[6][5] = {{8, 9, 10, 11, 20}, {8, 12, 13, 16, 21}, {9, 13, 14, 17, 22},
                              {10, 14, 15, 18, 23}, {11, 12, 15, 19, 24}, {16, 17, 18, 19, 25}};

// Text cursor reproducing the reference reader's grammar: `ignore(256, ':') >> value; ignore(256,'\n')`.
struct Cursor {
   const char* p;
   const char* end;
   bool ok = true;
   void skipLine() {
      while (p < end && *p != '\n') p++;
      if (p < end) p++;
   }
   void afterColon() {
      while (p < end && *p != ':') p++;
      if (p < end) p++; else ok = false;
   }
   long readInt() {
      char* q;
      long v = strtol(p, &q, 10);
      if (q == p) ok = false;
      p = q;
      return v;
   }
   double readDouble() {
      char* q;
      double v = strtod(p, &q);
      if (q == p) ok = false;
      p = q;
      return v;
   }
   void skipToken() {   // one whitespace-delimited token (the ':' separators inside BC lines)
      while (p < end && isspace((unsigned char)*p)) p++;
      while (p < end && !isspace((unsigned char)*p)) p++;
   }
};

struct Key3 {
   int a, b, c;
   bool operator==(const Key3& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct Key3Hash {
   size_t operator()(const Key3& k) const {
      uint64_t h = (uint64_t)(uint32_t)k.a * 0x9E3779B97F4A7C15ull;
      h ^= ((uint64_t)(uint32_t)k.b + 0x7F4A7C159E3779B9ull) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
      h ^= ((uint64_t)(uint32_t)k.c + 0x165667B19E3779F9ull) * 0x27D4EB2F165667C5ull + (h << 6) + (h >> 2);
      return (size_t)h;
   }
};

}   // namespace

bool read_inp(const std::string& path, Mesh& m) {
   std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
   if (!f) { set_error("cannot open %s", path.c_str()); return false; }
   const std::streamsize sz = f.tellg();
   f.seekg(0);
   std::string buf((size_t)sz + 1, '\0');
   if (sz > 0 && !f.read(&buf[0], sz)) { set_error("cannot read %s", path.c_str()); return false; }
   Cursor c{buf.data(), buf.data() + sz};
   c.skipLine(); c.skipLine();
   auto hdrInt = [&]() { c.afterColon(); long v = c.readInt(); c.skipLine(); return (int)v; };
   auto hdrDbl = [&]() { c.afterColon(); double v = c.readDouble(); c.skipLine(); return v; };
   m.eType = hdrInt(); m.NE = hdrInt(); m.NCN = hdrInt(); m.NENv = hdrInt(); m.NENp = hdrInt(); m.NGP = hdrInt();
   m.alpha = hdrDbl(); m.dt = hdrDbl(); m.t_ini = hdrDbl(); m.t_final = hdrDbl();
   m.maxIter = hdrInt(); m.tolerance = hdrDbl(); m.converge = hdrDbl(); m.isRestart = hdrInt();
   m.density = hdrDbl(); m.viscosity = hdrDbl(); m.fx = hdrDbl(); m.fy = hdrDbl();
   if (!c.ok) { set_error("%s: malformed header", path.c_str()); return false; }
   if (m.eType != 1 || m.NENp != 8 || (m.NENv != 27 && m.NENv != 8) || m.NGP != 8 || m.NE <= 0 || m.NCN <= 0) {
      set_error("%s: unsupported element (eType=%d NENv=%d NENp=%d NGP=%d); hexahedra 27/8/8 and 8/8/8 only", path.c_str(), m.eType,
                m.NENv, m.NENp, m.NGP);
      return false;
   }
   m.coord.assign((size_t)m.NCN * 3, 0.0);
   c.skipLine(); c.skipLine();
   for (int i = 0; i < m.NCN; i++) {
      c.readInt();
      m.coord[3 * (size_t)i] = c.readDouble(); m.coord[3 * (size_t)i + 1] = c.readDouble(); m.coord[3 * (size_t)i + 2] = c.readDouble();
      c.skipLine();
   }
   m.LtoGnode.assign((size_t)m.NE * m.NENv, -1);
   c.skipLine(); c.skipLine();
   for (int e = 0; e < m.NE; e++) {
      c.readInt();
      for (int i = 0; i < 8; i++) {
         const int n = (int)c.readInt() - 1;
         if (n < 0 || n >= m.NCN) c.ok = false;
         m.LtoGnode[(size_t)e * m.NENv + i] = n;
      }
      c.skipLine();
   }
   if (!c.ok) { set_error("%s: malformed node or element table", path.c_str()); return false; }
   c.skipLine(); c.skipLine();
   m.nBC = hdrInt();
   if (!c.ok || m.nBC < 0 || m.nBC > 1000000) { set_error("%s: malformed nBC", path.c_str()); return false; }
   m.BCtype.assign(m.nBC, 0.0); m.BCstr.assign((size_t)3 * m.nBC, 0.0);
   for (int i = 0; i < m.nBC; i++) {
      c.afterColon();
      m.BCtype[i] = c.readDouble();
      m.BCstr[3 * i] = c.readDouble(); c.skipToken();
      m.BCstr[3 * i + 1] = c.readDouble(); c.skipToken();
      m.BCstr[3 * i + 2] = c.readDouble();
      c.skipLine();
   }
   c.skipLine();
   const int nVel = hdrInt();
   const int nOut = hdrInt();
   if (!c.ok || nVel < 0 || nOut < 0) { set_error("%s: malformed BC counts", path.c_str()); return false; }
   auto readFaces = [&](int n, std::vector<int>& out) {
      out.resize((size_t)3 * n);
      for (int i = 0; i < n; i++) {
         const int e = (int)c.readInt() - 1, fc = (int)c.readInt() - 1, b = (int)c.readInt() - 1;
         if (e < 0 || e >= m.NE || fc < 0 || fc >= 6 || b < 0 || b >= m.nBC) c.ok = false;
         out[3 * (size_t)i] = e; out[3 * (size_t)i + 1] = fc; out[3 * (size_t)i + 2] = b;
         c.skipLine();
      }
   };
   c.skipLine(); c.skipLine();
   readFaces(nVel, m.velFaces);
   c.skipLine(); c.skipLine();
   readFaces(nOut, m.outFaces);
   if (!c.ok) { set_error("%s: malformed BC face list", path.c_str()); return false; }
   c.skipLine(); c.skipLine();
   m.zeroPressureNode = (int)c.readInt() - 1;
   c.skipLine();
   c.skipLine(); c.skipLine();
   m.monCoord[0] = c.readDouble(); m.monCoord[1] = c.readDouble(); m.monCoord[2] = c.readDouble();
   if (!c.ok) { set_error("%s: malformed pressure pin / monitor point", path.c_str()); return false; }
   if (m.zeroPressureNode >= m.NCN) { set_error("%s: pressure pin node out of range", path.c_str()); return false; }
   m.NNp = m.NCN;
   return true;
}

// Structured cavity in memory, numbered like the reference's cube generator
// (oldFiles/meshGenerators&Converters/HexaMeshGeneratorInACube_GeneratesCornerNodes.m:185-285):
// nodes x-fastest, elements x-fastest, BC faces bottom,left,right,front,back,top(lid).
void make_cavity(int nx, int ny, int nz, double cluster, int NENv, Mesh& m) {
   m = Mesh();
   m.NENv = NENv;
   m.NE = nx * ny * nz;
   m.NCN = (nx + 1) * (ny + 1) * (nz + 1);
   m.NNp = m.NCN;
   auto axis = [&](int n, double L) {
      const int nc = n + 1;
      std::vector<double> d(nc);
      if (cluster == 0.0) {
         for (int i = 0; i < nc; i++) d[i] = i * L / (nc - 1);
      } else {
         const int half = (nc + 1) / 2;
         const double mx = std::sinh(cluster);
         for (int i = 1; i <= half; i++) {
            const double xx = (1.0 * (i - 1)) / (half - 1);
            d[i - 1] = L / 2.0 / mx * std::sinh(cluster * xx);
         }
         for (int i = half; i <= nc; i++) d[i - 1] = L - d[nc + 1 - i - 1];
      }
      return d;
   };
   const std::vector<double> xs = axis(nx, 1.0), ys = axis(ny, 1.0), zs = axis(nz, 1.0);
   m.coord.resize((size_t)m.NCN * 3);
   size_t n = 0;
   for (int k = 0; k <= nz; k++)
      for (int j = 0; j <= ny; j++)
         for (int i = 0; i <= nx; i++, n++) { m.coord[3 * n] = xs[i]; m.coord[3 * n + 1] = ys[j]; m.coord[3 * n + 2] = zs[k]; }
   m.LtoGnode.assign((size_t)m.NE * NENv, -1);
   const int pl = (nx + 1) * (ny + 1);
   size_t e = 0;
   for (int k = 0; k < nz; k++)
      for (int j = 0; j < ny; j++)
         for (int i = 0; i < nx; i++, e++) {
            const int n1 = k * pl + j * (nx + 1) + i;
            int* L = &m.LtoGnode[e * NENv];
            L[0] = n1; L[1] = n1 + 1; L[2] = n1 + nx + 2; L[3] = n1 + nx + 1;
            L[4] = L[0] + pl; L[5] = L[1] + pl; L[6] = L[2] + pl; L[7] = L[3] + pl;
         }
   m.nBC = 2;
   m.BCtype.assign(2, 1.0);
   m.BCstr = {0.0, 0.0, 0.0, 1.0, 0.0, 0.0};
   auto el = [&](int i, int j, int k) { return k * nx * ny + j * nx + i; };
   auto add = [&](int elem, int face, int bc) { m.velFaces.push_back(elem); m.velFaces.push_back(face); m.velFaces.push_back(bc); };
   for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) add(el(i, j, 0), 0, 0);
   for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) add(el(0, j, k), 4, 0);
   for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) add(el(nx - 1, j, k), 2, 0);
   for (int k = 0; k < nz; k++) for (int i = 0; i < nx; i++) add(el(i, 0, k), 1, 0);
   for (int k = 0; k < nz; k++) for (int i = 0; i < nx; i++) add(el(i, ny - 1, k), 3, 0);
   for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) add(el(i, j, nz - 1), 5, 1);
   m.zeroPressureNode = (ny / 2) * (nx + 1) + (nx + 2) / 2 - 1;
   m.monCoord[0] = m.monCoord[1] = m.monCoord[2] = 0.5;
}

// setupNonCornerNodes (blascoCodinaHuerta.cpp:954-1320): ids 0..NCN-1 are the corner nodes; then
// mid-edge nodes in element/edge discovery order, then mid-face nodes, then one centre node per
// element.  An entity is identified by its sorted corner ids (conforming mesh), which yields the
// ids the reference finds by coordinate matching among neighbouring elements.
void generate_midnodes(Mesh& m) {
   const int nv = m.NENv;
   if (nv == 8) { m.NN = m.NCN; return; }
   const int NE = m.NE;
   m.coord.reserve((size_t)m.NCN * 3 * 8);
   int count = m.NCN;
   auto X = [&](int node, int d) { return m.coord[3 * (size_t)node + d]; };
   {
      std::unordered_map<uint64_t, int> edges;
      edges.reserve((size_t)NE * 4);
      for (int e = 0; e < NE; e++) {
         int* L = &m.LtoGnode[(size_t)e * nv];
         for (int ed = 0; ed < 12; ed++) {
            const int n1 = L[kEdgeCorner[ed][0]], n2 = L[kEdgeCorner[ed][1]];
            const uint64_t key = ((uint64_t)(uint32_t)std::min(n1, n2) << 32) | (uint32_t)std::max(n1, n2);
            auto it = edges.find(key);
            if (it == edges.end()) {
               edges.emplace(key, count);
               for (int d = 0; d < 3; d++) m.coord.push_back(0.5 * (X(n1, d) + X(n2, d)));
               L[8 + ed] = count++;
            } else {
               L[8 + ed] = it->second;
            }
         }
      }
   }
   {
      std::unordered_map<Key3, int, Key3Hash> faces;
      faces.reserve((size_t)NE * 4);
      for (int e = 0; e < NE; e++) {
         int* L = &m.LtoGnode[(size_t)e * nv];
         for (int fc = 0; fc < 6; fc++) {
            int q[4] = {L[kFaceCorner[fc][0]], L[kFaceCorner[fc][1]], L[kFaceCorner[fc][2]], L[kFaceCorner[fc][3]]};
            int s[4] = {q[0], q[1], q[2], q[3]};
            std::sort(s, s + 4);
            const Key3 key{s[0], s[1], s[2]};
            auto it = faces.find(key);
            if (it == faces.end()) {
               faces.emplace(key, count);
               for (int d = 0; d < 3; d++) m.coord.push_back(0.25 * (X(q[0], d) + X(q[1], d) + X(q[2], d) + X(q[3], d)));
               L[20 + fc] = count++;
            } else {
               L[20 + fc] = it->second;
            }
         }
      }
   }
   m.coord.resize((size_t)(count + NE) * 3);
   for (int e = 0; e < NE; e++) {
      int* L = &m.LtoGnode[(size_t)e * nv];
      for (int d = 0; d < 3; d++)
         m.coord[3 * (size_t)(count + e) + d] =
             0.125 * (X(L[0], d) + X(L[1], d) + X(L[2], d) + X(L[3], d) + X(L[4], d) + X(L[5], d) + X(L[6], d) + X(L[7], d));
      L[26] = count + e;
   }
   m.NN = count + NE;
}

// determineVelBCnodes (blascoCodinaHuerta.cpp:1410-1580): faces in file order, later entries win.
void find_bc_nodes(Mesh& m) {
   std::vector<int> which((size_t)m.NN, -1);
   const int nv = m.NENv;
   const size_t nf = m.velFaces.size() / 3;
   for (size_t i = 0; i < nf; i++) {
      const int e = m.velFaces[3 * i], fc = m.velFaces[3 * i + 1], b = m.velFaces[3 * i + 2];
      const int* L = &m.LtoGnode[(size_t)e * nv];
      for (int q = 0; q < 4; q++) which[L[kFaceCorner[fc][q]]] = b;
      if (nv == 27)
         for (int q = 0; q < 5; q++) which[L[kFaceInner[fc][q]]] = b;
   }
   m.BCvelNodes.clear();
   for (int i = 0; i < m.NN; i++)
      if (which[i] >= 0) { m.BCvelNodes.push_back(i); m.BCvelNodes.push_back(which[i]); }
   m.BCnVelNodes = (int)m.BCvelNodes.size() / 2;
}

// findMonitorPoint (blascoCodinaHuerta.cpp:1646-1668): nearest corner node, lowest id on ties.
void find_monitor(Mesh& m) {
   double best = 1e6;
   m.monPoint = 0;
   for (int i = 0; i < m.NCN; i++) {
      const double dx = m.coord[3 * (size_t)i] - m.monCoord[0], dy = m.coord[3 * (size_t)i + 1] - m.monCoord[1],
                   dz = m.coord[3 * (size_t)i + 2] - m.monCoord[2];
      const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
      if (d < best) { best = d; m.monPoint = i; }
   }
}

// setupMeshColoring (blascoCodinaHuerta.cpp:853-947): greedy colouring over the corner-sharing
// neighbours.  Neighbour order (corner 0..7, then ascending element id) and the scan-restart
// behaviour at :880-886 are reproduced so that the colour of every element matches the reference.
// The CUDA kernels scatter with atomics; colours are exposed for callers that want them.
void color_mesh(Mesh& m) {
   const int NE = m.NE, nv = m.NENv, NNp = m.NNp;
   std::vector<int> start((size_t)NNp + 1, 0), items((size_t)NE * 8);
   for (int e = 0; e < NE; e++)
      for (int i = 0; i < 8; i++) start[m.LtoGnode[(size_t)e * nv + i] + 1]++;
   for (int i = 0; i < NNp; i++) start[i + 1] += start[i];
   {
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int e = 0; e < NE; e++)
         for (int i = 0; i < 8; i++) items[fill[m.LtoGnode[(size_t)e * nv + i]]++] = e;
   }
   const int LARGE = 30;
   m.meshColors.assign(NE, -1);
   m.NmeshColors.assign(LARGE, 0);
   std::vector<int> nb;
   for (int e = 0; e < NE; e++) {
      nb.clear();
      for (int i = 0; i < 8; i++) {
         const int node = m.LtoGnode[(size_t)e * nv + i];
         for (int t = start[node]; t < start[node + 1]; t++) {
            const int o = items[t];
            if (o != e && std::find(nb.begin(), nb.end(), o) == nb.end()) nb.push_back(o);
         }
      }
      int cand = 0;
      for (int i = 0; i < (int)nb.size(); i++)
         if (cand == m.meshColors[nb[i]]) { cand++; i = 0; }
      cand = std::min(cand, LARGE - 1);
      m.meshColors[e] = cand;
      m.NmeshColors[cand]++;
   }
   m.elementsOfColor.assign(NE, 0);
   m.nActiveColors = 0;
   int pos = 0;
   for (int c = 0; c < LARGE; c++) {
      if (!m.NmeshColors[c]) continue;
      m.nActiveColors++;
      for (int e = 0; e < NE; e++)
         if (m.meshColors[e] == c) m.elementsOfColor[pos++] = e;
   }
}

// CSR pattern of (row nodes = first nRowLoc local nodes) x (column nodes = first nColLoc local nodes)
// with ascending columns, plus the [e][i][j] -> CSR position map (setupSparseM / setupSparseG,
// blascoCodinaHuerta.cpp:1675-2159; setupSparseZGPU, guermondQuartapelle.cpp:2418-2647).
void build_pattern(const Mesh& m, int nRowLoc, int nColLoc, int nrows, int ncols, CsrPattern& P, std::vector<int>* map,
                   const std::vector<int>* rowIndex, const std::vector<int>* colIndex) {
   const int NE = m.NE, nv = m.NENv;
   auto R = [&](int node) { return rowIndex ? (*rowIndex)[node] : node; };
   auto C = [&](int node) { return colIndex ? (*colIndex)[node] : node; };
   std::vector<int> start((size_t)nrows + 1, 0), items((size_t)NE * nRowLoc);
   for (int e = 0; e < NE; e++)
      for (int i = 0; i < nRowLoc; i++) start[R(m.LtoGnode[(size_t)e * nv + i]) + 1]++;
   for (int i = 0; i < nrows; i++) start[i + 1] += start[i];
   {
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int e = 0; e < NE; e++)
         for (int i = 0; i < nRowLoc; i++) items[fill[R(m.LtoGnode[(size_t)e * nv + i])]++] = e;
   }
   P.nrows = nrows; P.ncols = ncols;
   P.rowStarts.assign((size_t)nrows + 1, 0);
   std::vector<int> cnt(nrows, 0);
   auto gather = [&](int r, std::vector<int>& tmp) {
      tmp.clear();
      for (int t = start[r]; t < start[r + 1]; t++) {
         const int* L = &m.LtoGnode[(size_t)items[t] * nv];
         for (int j = 0; j < nColLoc; j++) tmp.push_back(C(L[j]));
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
   };
#pragma omp parallel
   {
      std::vector<int> tmp;
#pragma omp for schedule(static)
      for (int r = 0; r < nrows; r++) {
         gather(r, tmp);
         cnt[r] = (int)tmp.size();
      }
   }
   for (int r = 0; r < nrows; r++) P.rowStarts[r + 1] = P.rowStarts[r] + cnt[r];
   P.col.resize((size_t)P.rowStarts[nrows]);
#pragma omp parallel
   {
      std::vector<int> tmp;
#pragma omp for schedule(static)
      for (int r = 0; r < nrows; r++) {
         gather(r, tmp);
         std::copy(tmp.begin(), tmp.end(), P.col.begin() + P.rowStarts[r]);
      }
   }
   if (map) {
      map->resize((size_t)NE * nRowLoc * nColLoc);
#pragma omp parallel for schedule(static)
      for (int e = 0; e < NE; e++) {
         const int* L = &m.LtoGnode[(size_t)e * nv];
         for (int i = 0; i < nRowLoc; i++) {
            const int r = R(L[i]);
            const int* b = P.col.data() + P.rowStarts[r];
            const int* en = P.col.data() + P.rowStarts[r + 1];
            for (int j = 0; j < nColLoc; j++)
               (*map)[((size_t)e * nRowLoc + i) * nColLoc + j] = (int)(std::lower_bound(b, en, C(L[j])) - P.col.data());
         }
      }
   }
}

// Element-block partition: rank r owns the contiguous element range [r*NE/world, (r+1)*NE/world) (z-slabs on
// the generators' x-fastest numbering) and every node whose lowest-ranked touching element it owns.  Its
// sub-mesh holds all elements within `layers` element layers of an owned node, so that every owned ROW of
// every operator (and, with 2 layers, of Z = G^T Md^-1 G) is assembled completely from local data.  All ranks
// derive the same plan from the same global mesh, so no set-up communication is needed.
void partition_mesh(const Mesh& g, int world, int rank, int layers, Mesh& loc, Partition& part) {
   const int NE = g.NE, nv = g.NENv, NN = g.NN;
   auto eOwner = [&](int e) { return (int)(((long long)e * world) / NE); };
   // rank r owns element e iff eOwner(e) == r  (monotone, contiguous)
   std::vector<int> nOwner((size_t)NN, world);
   for (int e = 0; e < NE; e++) {
      const int o = eOwner(e);
      const int* L = &g.LtoGnode[(size_t)e * nv];
      for (int i = 0; i < nv; i++) nOwner[L[i]] = std::min(nOwner[L[i]], o);
   }
   typedef unsigned long long mask_t;
   // ranks whose sub-mesh contains element e, grown layer by layer
   std::vector<mask_t> eMask((size_t)NE, 0), nMask((size_t)NN, 0);
   for (int e = 0; e < NE; e++) {
      const int* L = &g.LtoGnode[(size_t)e * nv];
      mask_t mk = 0;
      for (int i = 0; i < nv; i++) mk |= (mask_t)1 << nOwner[L[i]];
      eMask[e] = mk;
   }
   std::vector<mask_t> eMaskA = eMask;   // layer 1: elements touching an owned node
   for (int l = 1; l < layers; l++) {
      std::fill(nMask.begin(), nMask.end(), 0);
      for (int e = 0; e < NE; e++) {
         const int* L = &g.LtoGnode[(size_t)e * nv];
         for (int i = 0; i < nv; i++) nMask[L[i]] |= eMask[e];
      }
      for (int e = 0; e < NE; e++) {
         const int* L = &g.LtoGnode[(size_t)e * nv];
         mask_t mk = eMask[e];
         for (int i = 0; i < nv; i++) mk |= nMask[L[i]];
         eMask[e] = mk;
      }
   }
   std::fill(nMask.begin(), nMask.end(), 0);   // ranks that hold node n (owned or ghost)
   for (int e = 0; e < NE; e++) {
      const int* L = &g.LtoGnode[(size_t)e * nv];
      for (int i = 0; i < nv; i++) nMask[L[i]] |= eMask[e];
   }
   const mask_t me = (mask_t)1 << rank;
   part = Partition();
   part.rank = rank; part.world = world;
   // ---- local elements: those touching an owned node first (EBE range), then the outer layer(s)
   for (int e = 0; e < NE; e++)
      if (eMaskA[e] & me) part.elemGlobal.push_back(e);
   part.NEa = (int)part.elemGlobal.size();
   for (int e = 0; e < NE; e++)
      if ((eMask[e] & me) && !(eMaskA[e] & me)) part.elemGlobal.push_back(e);
   const int NEl = (int)part.elemGlobal.size();
   // ---- local nodes: owned (ascending global id), then ghosts grouped by owner (ascending global id inside)
   std::vector<std::vector<int> > ghostsOf(world);
   for (int n = 0; n < NN; n++) {
      if (!(nMask[n] & me)) continue;
      if (nOwner[n] == rank) part.nodeGlobal.push_back(n);
      else ghostsOf[nOwner[n]].push_back(n);
   }
   part.NNo = (int)part.nodeGlobal.size();
   std::vector<int> g2l((size_t)NN, -1);
   for (int q = 0; q < world; q++) {
      if (ghostsOf[q].empty()) continue;
      PeerPlan pp;
      pp.rank = q;
      pp.recvVoff = (int)part.nodeGlobal.size();
      pp.recvVcnt = (int)ghostsOf[q].size();
      part.nodeGlobal.insert(part.nodeGlobal.end(), ghostsOf[q].begin(), ghostsOf[q].end());
      part.peers.push_back(pp);
   }
   const int NNl = (int)part.nodeGlobal.size();
   for (int i = 0; i < NNl; i++) g2l[part.nodeGlobal[i]] = i;
   // ---- pressure DOFs = corner nodes (global id < NCN), same ordering rule
   std::vector<int> presLocalOfNode((size_t)NNl, -1);
   for (int i = 0; i < part.NNo; i++)
      if (part.nodeGlobal[i] < g.NCN) { presLocalOfNode[i] = (int)part.presGlobal.size(); part.presGlobal.push_back(part.nodeGlobal[i]); }
   part.NPo = (int)part.presGlobal.size();
   for (PeerPlan& pp : part.peers) {
      pp.recvPoff = (int)part.presGlobal.size();
      for (int i = pp.recvVoff; i < pp.recvVoff + pp.recvVcnt; i++)
         if (part.nodeGlobal[i] < g.NCN) { presLocalOfNode[i] = (int)part.presGlobal.size(); part.presGlobal.push_back(part.nodeGlobal[i]); }
      pp.recvPcnt = (int)part.presGlobal.size() - pp.recvPoff;
   }
   // ---- what I send: my owned nodes that rank q holds, ascending global id (= q's ghost order for owner me)
   auto peerOf = [&](int q) -> PeerPlan& {
      for (PeerPlan& pp : part.peers)
         if (pp.rank == q) return pp;
      PeerPlan pp;
      pp.rank = q;
      pp.recvVoff = NNl; pp.recvPoff = (int)part.presGlobal.size();
      part.peers.push_back(pp);
      return part.peers.back();
   };
   for (int i = 0; i < part.NNo; i++) {
      const int n = part.nodeGlobal[i];
      mask_t others = nMask[n] & ~me;
      while (others) {
         const int q = __builtin_ctzll(others);
         others &= others - 1;
         PeerPlan& pp = peerOf(q);
         pp.sendV.push_back(i);
         if (n < g.NCN) pp.sendP.push_back(presLocalOfNode[i]);
      }
   }
   std::sort(part.peers.begin(), part.peers.end(), [](const PeerPlan& a, const PeerPlan& b) { return a.rank < b.rank; });
   // remote offsets (for direct peer-memory pushes): rank q numbers its owned entries first, then its ghosts
   // grouped by owner in ascending owner order -> my block starts at owned(q) + sum_{o < me} ghosts(q, o)
   {
      std::vector<long long> ownedV(world, 0), ownedP(world, 0), cntV((size_t)world * world, 0), cntP((size_t)world * world, 0);
      for (int n = 0; n < NN; n++) {
         const int o = nOwner[n];
         mask_t mk = nMask[n];
         while (mk) {
            const int q = __builtin_ctzll(mk);
            mk &= mk - 1;
            if (q == o) { ownedV[q]++; if (n < g.NCN) ownedP[q]++; }
            else { cntV[(size_t)q * world + o]++; if (n < g.NCN) cntP[(size_t)q * world + o]++; }
         }
      }
      for (PeerPlan& pp : part.peers) {
         long long ov = ownedV[pp.rank], op = ownedP[pp.rank];
         for (int o = 0; o < rank; o++) { ov += cntV[(size_t)pp.rank * world + o]; op += cntP[(size_t)pp.rank * world + o]; }
         pp.remoteVoff = (int)ov; pp.remotePoff = (int)op;
      }
   }
   // ---- the local mesh
   loc = Mesh();
   loc.eType = g.eType; loc.NENv = nv; loc.NENp = g.NENp; loc.NGP = g.NGP;
   loc.alpha = g.alpha; loc.dt = g.dt; loc.t_ini = g.t_ini; loc.t_final = g.t_final; loc.maxIter = g.maxIter;
   loc.tolerance = g.tolerance; loc.converge = g.converge; loc.isRestart = g.isRestart; loc.density = g.density;
   loc.viscosity = g.viscosity; loc.fx = g.fx; loc.fy = g.fy;
   loc.NE = NEl; loc.NN = NNl; loc.NNp = (int)part.presGlobal.size(); loc.NCN = loc.NNp;
   loc.coord.resize((size_t)NNl * 3);
   for (int i = 0; i < NNl; i++)
      for (int d = 0; d < 3; d++) loc.coord[3 * (size_t)i + d] = g.coord[3 * (size_t)part.nodeGlobal[i] + d];
   loc.LtoGnode.resize((size_t)NEl * nv);
   std::vector<int> e2l((size_t)NE, -1);
   for (int el = 0; el < NEl; el++) {
      const int e = part.elemGlobal[el];
      e2l[e] = el;
      for (int i = 0; i < nv; i++) loc.LtoGnode[(size_t)el * nv + i] = g2l[g.LtoGnode[(size_t)e * nv + i]];
   }
   loc.presIndex = presLocalOfNode;
   loc.nBC = g.nBC; loc.BCtype = g.BCtype; loc.BCstr = g.BCstr;
   for (size_t f = 0; f < g.velFaces.size() / 3; f++) {   // file order kept: later faces still win
      const int el = e2l[g.velFaces[3 * f]];
      if (el >= 0) { loc.velFaces.push_back(el); loc.velFaces.push_back(g.velFaces[3 * f + 1]); loc.velFaces.push_back(g.velFaces[3 * f + 2]); }
   }
   find_bc_nodes(loc);
   const int pin = g.zeroPressureNode;
   part.pinLocal = (pin > 0 && g2l[pin] >= 0) ? presLocalOfNode[g2l[pin]] : -1;
   loc.zeroPressureNode = part.pinLocal;
   part.monLocal = (g2l[g.monPoint] >= 0 && g2l[g.monPoint] < part.NNo) ? g2l[g.monPoint] : -1;
   loc.monPoint = std::max(part.monLocal, 0);
   for (int d = 0; d < 3; d++) loc.monCoord[d] = g.monCoord[d];
}

}   // namespace cfdb
