// Device-side data model of the fractional-step time loop (product code, sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "common.hpp"

namespace cfdb {

#define CFD_CK(call)                                                                          \
   do {                                                                                       \
      cudaError_t e_ = (call);                                                                \
      if (e_ != cudaSuccess) {                                                                \
         cfdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
         return CFD_ERR_CUDA;                                                                 \
      }                                                                                       \
   } while (0)

// ---------------------------------------------------------------------------------------------
// Sliced CSR ("SELL-32"): the device layout every CSR matrix of the reference is converted to.
//   * rows are grouped in slices of R = 32/T rows; one warp owns one slice, T lanes share a row
//   * every lane of a slice holds the same even number W of (col,val) entries, stored as PAIRS
//     interleaved across the 32 lanes: entry j of lane t lives at  off + (j/2)*64 + 2*t + (j&1)
//     -> one warp-wide load fetches 32 double2 (512 B) / 32 int2 (256 B), fully coalesced, 16-byte
//        aligned: the 128-bit path the byte model in DESIGN.md assumes
//   * padding entries carry val = 0 and a valid column of the same row
//   * several value arrays (K/M/A, G1/G2/G3) share one pattern, like the reference's CSR triplets
// ---------------------------------------------------------------------------------------------
struct Sell {
   int nrows = 0, ncols = 0;
   int T = 1;              // lanes per row (1,2,4,8)
   int R = 32;             // rows per slice
   int nslices = 0;
   int nval = 0;
   long long nnz = 0;      // true non-zeros (algorithmic bytes use this)
   long long padded = 0;   // stored entries
   long long* sliceOff = nullptr;   // [nslices+1] entry offsets of the value arrays (multiples of 64)
   // Column indices.  colOff[s] = start of slice s inside `cols` (even), with bit 0 = 1 when the slice is
   // OFFSET-COMPRESSED: all its rows share one list of W column offsets (col - row), stored once (W ints) instead
   // of 32*W explicit columns -- lossless, detected per slice at build time (structured numbering).  With T lanes
   // per row the list holds W*T offsets, [slot][lane-in-row].
   long long* colOff = nullptr;     // [nslices]
   long long nindex = 0;            // ints stored in `cols`
   int anyComp = 0;                 // 0: no compressed slice -> cols are laid out exactly like the values (colOff unused)
   int* cols = nullptr;
   double* vals[3] = {nullptr, nullptr, nullptr};
};

struct SellHost {   // result of the host-side conversion
   Sell meta;       // device pointers unset
   std::vector<long long> sliceOff, colOff;
   std::vector<int> cols;
   long long compressedSlices = 0;
   std::vector<long long> perm;   // CSR position -> SELL position
};

int choose_T(int nrows, long long nnz);
void sell_build_host(int nrows, int ncols, const int* rowStarts, const int* col, int T, SellHost& out);
int sell_upload(const SellHost& h, int nval, const double* const* csrVals, Sell& dev);
void sell_free(Sell& s);

struct Comm;   // NCCL communicator wrapper (dist.cuh)

// per kernel-class profile (CUDA events on the context stream)
struct ProfEntry {
   double ms = 0;
   long long launches = 0;
   double bytes = 0;    // algorithmic bytes (CSR model)
   double lbytes = 0;   // bytes of the device layout as stored (padding in, compressed indices out): what the kernel has to move
};
// the two byte counts of one launch
struct Bytes {
   double a, l;
   Bytes(double v = 0.0) : a(v), l(v) {}
   Bytes(double a_, double l_) : a(a_), l(l_) {}
};
inline Bytes operator+(const Bytes& x, const Bytes& y) { return Bytes(x.a + y.a, x.l + y.l); }

}   // namespace cfdb

// The opaque context of include/cfdb200.h
struct cfd_ctx {
   cfd_problem P;   // scalar fields only are valid after init (pointers nulled)
   cfd_options opt;
   int device = 0;
   cudaStream_t stream = nullptr;
   int numSMs = 148;
   // matrices
   cfdb::Sell K;    // vals[0] = K (explicit) | vals[0]=K+M ("MK"), vals[1]=M, vals[2]=A (implicit)
   cfdb::Sell G;    // vals[0..2] = G1,G2,G3           rows = velocity nodes
   cfdb::Sell GT;   // vals[0..2] = G1^T,G2^T,G3^T     rows = pressure nodes
   cfdb::Sell Z;    // vals[0]
   double* Zdinv = nullptr;     // 1/diag(Z)
   // mesh / EBE data
   int ldE = 0;                 // padded NE (leading dimension of the transposed element tables)
   int* LtoGT = nullptr;        // [NENv][ldE] node ids, transposed for coalesced per-element access
   double* coord = nullptr;     // [NN][3]
   int* mapT = nullptr;         // implicit: [NENv*NENv][ldE] SELL positions of A entries
   int* diagPosA = nullptr;     // implicit: [NN] SELL position of the diagonal
   int* elemOfColor = nullptr;  // [NEa] elements grouped by colour (ebe_scatter = 1)
   std::vector<int> colorStart; // [ncolors+1]
   int* bcIndex = nullptr;      // [NN] whichBC or -1
   int* bcNodes = nullptr;      // [BCnVelNodes]
   double* BCstr = nullptr;     // [nBC][3]
   double *MdInv = nullptr, *MdOrigInv = nullptr;
   // named work vectors
   std::map<std::string, std::pair<double**, long long> > vecs;
   // explicit
   double *Un = nullptr, *Unp1 = nullptr, *Unp1_prev = nullptr, *UnpHalf = nullptr, *UnpHalf_prev = nullptr;
   double *Acc = nullptr, *Acc_prev = nullptr, *KtimesAcc_prev = nullptr, *R1 = nullptr, *R3 = nullptr;
   double *Pn = nullptr, *Pnp1 = nullptr, *Pnp1_prev = nullptr, *Pdot = nullptr, *R2 = nullptr;
   // implicit
   double *Uk = nullptr, *Uk_prev = nullptr, *Pk = nullptr, *Pk_prev = nullptr, *Pk_prevprev = nullptr, *Pdiff = nullptr;
   double* R1i = nullptr;       // [3][NN] the three momentum right-hand sides
   double* Adinv = nullptr;     // [NN]
   double* bw = nullptr;        // BiCGStab work: 8 x [3][NN] (partitioned: 6, ph and shv live in the exchange window)
   double *bw_ph = nullptr, *bw_shv = nullptr;
   // Krylov work
   double *cg_r = nullptr, *cg_p = nullptr, *cg_q = nullptr, *cg_s = nullptr, *cg_pp = nullptr;   // (cg_s, cg_pp: single-reduction form only)
   int cg_variant = 1;          // 1 classical, 2 Chronopoulos-Gear (resolved from cfd_options.cg_variant at init)
   int pcgcg_blocks = 0;
   double* partials = nullptr;  // grid-reduction scratch
   double* kinfo = nullptr;     // device: Krylov results (iterations, residuals)
   double* scal = nullptr;      // device: [0..3] convergence sums, [4] maxAcc bits
   double* h_pinned = nullptr;  // pinned host mirror for read-backs (256 doubles; [128..191] = Krylov results)
   int pcg_blocks = 0, pcg_threads = 512, bicg_blocks = 0, bicg_threads = 512;
   int pcg_minb = 3;
   int pcgDynStart = 0;             // slices >= this are scheduled dynamically inside k_pcg
   int* pcgDynCounter = nullptr;
   int bicgDynStart = 0;            // same for the momentum matrix inside k_bicgstab3
   int* bicgDynCounter = nullptr;
   int *pcgPart = nullptr, *bicgPart = nullptr;   // [blocks+1] slice ranges per block, balanced by stored entries
   int* pcgcgPart = nullptr;                      // (= pcgPart: the context runs one of the two PCG kernels)
   int* pcgSmemPart = nullptr;                    // [numSMs+1] slice ranges of the shared-memory-resident PCG (small systems)
   size_t pcgSmemBytes = 0;                       // its dynamic shared memory per CTA; 0 = not eligible
   bool pcgSmemSingle = false;                    // resident kernel runs the single-reduction recurrences
   // k_pcg_res (row blocks resident per CTA, staged vector window): device arrays of its private layout; resBytes = 0: not eligible
   void* resCta = nullptr;
   double* resVals = nullptr;
   unsigned short* resIdx = nullptr;
   int *resCols = nullptr, *resGrp = nullptr;
   int resT = 0;
   size_t resBytes = 0;
   std::vector<long long> hostOffZ, hostOffK;     // host copies of the slice offsets (partitioning)
   // bookkeeping
   int timeN = 0;
   int last_cg = 0, last_bicg[3] = {0, 0, 0};
   int last_cg_conv = 1, last_bicg_conv[3] = {1, 1, 1}, last_bicg_restarts[3] = {0, 0, 0}, last_bicg_passes = 0;
   double last_cg_res = 0.0, last_bicg_true[3] = {0, 0, 0};
   long long launches = 0;
   bool profile = false;
   std::map<std::string, cfdb::ProfEntry> prof;
   struct ProfRec { const char* cls; cudaEvent_t a, b; double bytes, lbytes; };
   std::vector<ProfRec> recs;   // event-pair pool (grown on demand, reused)
   size_t recsUsed = 0;
   int lastPcgRec = -1, lastBicgRec = -1;
   bool ownStream = true;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   std::vector<void*> allocs;
   // ---- partition (multi-GPU).  NN / NNp in P are LOCAL counts (owned + ghost); rows = owned.
   int NNo = 0, NPo = 0, NEa = 0;
   int pin = -1;                    // pressure DOF to pin (local), -1 = none
   bool monOwned = true;
   cfdb::Comm* comm = nullptr;      // NCCL communicator (world > 1)
   std::vector<int> peerRank, sendVptr, recvVoff, recvVcnt, sendPptr, recvPoff, recvPcnt;
   int *sendVidx = nullptr, *sendPidx = nullptr;   // device
   double* sendBuf = nullptr;                        // device staging for packed interface values
   double* dacc = nullptr;                           // device: distributed-Krylov accumulators + state
   // ---- peer-memory (CUDA IPC over NVLink) exchange window, optional
   bool p2p = false;
   char* win = nullptr;             // this rank's window: flags | mailboxes | p
   size_t winBytes = 0;
   long long winOffP = 0, winOffV = 0;   // byte offsets of the pressure vector and of ph | shv inside the window
   std::vector<char*> peerWin;      // [world] mapped peer windows (own included)
   std::vector<long long> peerOffP, peerOffV; // [world]
   std::vector<int> peerNN;         // [world] local velocity-node counts (component strides) of the peers
   char** dPeerWin = nullptr;       // device copy
   double** dPeerP = nullptr;       // [nPeers] start of my block inside each peer's p ghosts
   int* dPeerRank = nullptr;
   int* pushPtrP = nullptr;         // [NPo+1] CSR of (peer, slot) pushes per owned pressure DOF
   int* pushPtrV = nullptr;         // same for the owned velocity nodes
   int2* pushDstV = nullptr;
   double** dPeerV = nullptr;       // [nPeers] start of my block inside each peer's ph ghosts
   int* dPeerLd = nullptr;          // [nPeers]
   int* ifaceP = nullptr;           // owned pressure DOFs with at least one push (done first, so the transfer overlaps the rest)
   int nIfaceP = 0;
   unsigned long long p2pTimeoutNs = 4000000000ull;
   int2* pushDstP = nullptr;
   unsigned long long* xstate = nullptr;   // device: [0] allreduce epoch [1] halo epoch [2] error [3] reduction ticket [4] interface ticket
   std::vector<int> remoteVoff, remotePoff;
};
