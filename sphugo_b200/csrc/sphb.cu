// sphb.cu — host side of libsphb.so: the C ABI declared in include/sphb.h.
//
// One sphb_sim = one GPU = one CUDA stream.  All particle state is device resident SoA; the
// entry points only enqueue kernels (sphb_step is asynchronous) and nothing on the step path
// needs a host round trip: the cell grid lives in device memory and is rebuilt by a kernel.
// There is no CPU fallback: every numeric result comes out of the kernels in sphb_kernels.cuh.
//
// Reference being replaced: (*Simulation).Step / CalculateForces, sim/sph.go:64-198,403-435.
#include "sphb_kernels.cuh"
#include "../../include/sphb.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct Soa {  // one copy of the per-particle state (two copies: the reorder gathers from one into the other)
  double2 *pos = nullptr, *vel = nullptr, *vdot = nullptr, *vpred = nullptr;
  double *e = nullptr, *edot = nullptr, *epred = nullptr;
  int64_t* id = nullptr;
  double4* pc = nullptr;  // {rho, c, h, P/rho^2}
  uint8_t* ghost = nullptr;
};

}  // namespace

// the slab ring inside the library (sphb_ring.inc)
#define RING_SLACK 0.5  // an owned particle may sit this many h_max outside its slab before it must migrate
struct RingState {
  bool on = false;
  int rank = 0, nranks = 1;
  bool periodic = false;
  int left = -1, right = -1;             // neighbour ranks (-1: none)
  double x_lo = 0.0, x_hi = 0.0, safety = 1.15;
  int64_t halo_cap = 0;                  // records per side
  double h_max = 0.0, v_max = 0.0;       // all-reduced maxima of the last cycle
  double h_max_run = 0.0, v_max_run = 0.0; // the same, running (read every evaluation)
  double excursion = 0.0;                // bound of the drift out of the slab since the last migration
  int64_t n_global = 0;
  int migrate_every = 0, migrations = 0;
  int cycle_len = 0;                     // evaluations since the last rebuild (inclusive)
  int pending_kind = 0, pending_age = 0; // feedback on its way: 1 = of a rebuild, 2 = of a reuse evaluation (and its age)
  bool want_idx = false;                 // the halo pack in progress records the source indices
  double* sbuf[2] = {nullptr, nullptr};  // send: low side, high side
  double* rbuf[2] = {nullptr, nullptr};  // receive: from the left, from the right
  int* sidx[2] = {nullptr, nullptr};     // pre-sort index of every packed particle
  uint32_t* hsrc[2] = {nullptr, nullptr};// sorted index of every packed particle
  uint32_t* inv = nullptr;               // inverse permutation of the last reorder
  int64_t inv_cap = 0;
  int64_t nsend[2] = {0, 0}, nrecv[2] = {0, 0}, ghost0 = 0;
  double* red_dev = nullptr;             // small device scratch of the host all-reduces
  struct RingMsg* msg = nullptr;         // device block of the single-wait rebuild (sphb_ring.cuh)
  int* refused_host = nullptr;           // pinned
  void* comm = nullptr;                  // ncclComm_t
  int comm_rank = 0, comm_size = 1;
  struct sphb_sim* peer[2] = {nullptr, nullptr};  // slabs in one process
};

struct sphb_sim {
  sphb_params prm{};
  int device = 0;
  int64_t n = 0;        // owned particles
  int64_t nghost = 0;   // slab mode: ghosts appended behind the owned ones until step_end
  int64_t cap = 0;      // capacity (owned + ghosts)
  int64_t cur_step = 0;
  cudaStream_t st = nullptr;
  Soa a, b;             // a = current
  double2* spos = nullptr;
  double* hguess = nullptr;
  uint32_t *keys = nullptr, *keysSorted = nullptr, *rank = nullptr, *perm = nullptr;
  uint32_t *cellCount = nullptr, *tileSum = nullptr;
  int ntiles_cap = 0;
  uint32_t* cellStart = nullptr;
  uint32_t* nn = nullptr;
  int* failList = nullptr;
  int* failCount = nullptr;   // [0] = count; [1] = cumulative fallback particles
  uint32_t* dflags = nullptr;
  double* statPart = nullptr;
  double* stats = nullptr;    // STAT_N doubles
  GridP* grid = nullptr;
  GridP* grid_next = nullptr;  // the next evaluation's grid, built right after the kNN of a fused step
  bool grid_next_ready = false;
  bool fuse_keys = false;      // the force launch in progress emits the next step's keys
  bool keys_ready = false;     // keys / rank / cellCount already hold the next step's cell keys (force epilogue)
  // slab mode (ring fast path): the slab the NEXT evaluation will set is known before this evaluation's force kernel (its
  // ghost widths come from maxima that lag one evaluation), so the fusion applies there too: owned particles get their
  // keys in the epilogue, the ghosts theirs when they arrive
  bool next_slab_valid = false;
  sphb_slab next_slab{}, keys_slab{};
  double keys_dtH = 0.0, next_hor[2] = {0, 0}, next_ver[2] = {0, 0};
  int keys_n = 0;
  void* scratch = nullptr;    // download / upload staging
  size_t scratchBytes = 0;
  // split upload (sphb_upload_by_id_begin / _end): the host-to-device copies run on a copy stream into a staging area of
  // their own, next to whatever steps are enqueued on the compute stream
  cudaStream_t st_copy = nullptr;
  cudaEvent_t up_done = nullptr, up_free = nullptr;
  void* up_stage = nullptr;
  size_t up_stage_bytes = 0;
  uint32_t up_mask = 0;       // fields of the upload in flight (0: none)
  int64_t up_n = 0;
  int ncell_max = 0;
  bool stats_dirty = true;
  bool have_list = false;     // nn/spos/grid describe the current particle order
  bool have_h = false;        // pc.z holds smoothing lengths of a previous evaluation (search radius guess)
  double knn_hor[2] = {0, 0}, knn_ver[2] = {0, 0};
  // slab mode
  bool slab_on = false;
  sphb_slab slab{};
  int slab_mode = 0;           // mode of the evaluation in progress (sphb_slab_step_begin)
  int* packCount = nullptr;    // device counter of the pack kernels
  int64_t nleaving = 0;        // particles packed for migration and not yet compacted away
  cudaEvent_t ev[SPHB_PH_COUNT + 1] = {};
  bool ev_valid = false;
  int64_t counters[SPHB_CNT_COUNT] = {};
  unsigned long long* hacc = nullptr;  // [HACC_N][2] smoothing-length accumulator written by the kNN kernels
  uint32_t* qmax = nullptr;   // {max h, max |v|^2} of the owned particles as float bits (kNN / force kernels)
  bool qmax_valid = false;    // the statistics pass was skipped: sphb_max_h / sphb_max_speed read qmax
  bool hacc_valid = false;    // it describes the current particles (no upload / append since the evaluation that filled it)
  double hscale = 0.0;        // fixed-point scale of the evaluation in progress (0: accumulator off)
  int ids_dense = -1;         // -1 unknown, 0 no, 1 the ids are a permutation of 0..n-1 (by-id upload / frame)
  GridTune gtune{};
  KnnTune ktune{};
  int force_nrec = 672;       // staged neighbour records per force block (shared memory)
  // certified reuse of the neighbour lists (sphb_kernels.cuh, ReuseState)
  uint32_t* nx = nullptr;     // further candidates [tile][SPHB_KX][lane]
  TileInfo* tinfo = nullptr;  // [tile]: extent of the staged block (npc = 0: none)
  uint2* ptab = nullptr;      // [tile][32]: its pieces
  double* dexcl = nullptr;    // exclusion radius per particle (0: no further candidates known)
  float2* ucum = nullptr;     // cumulative displacement since the rebuild (local displacement bound)
  float4* ubox = nullptr;     // its bounding box per coarse cell
  int ubox_cap = 0;
  uint32_t* ubox_ok = nullptr;
  ReuseState* rs = nullptr;   // device bookkeeping
  ReuseStat* stat_dev = nullptr;
  ReuseStat* stat_host = nullptr;  // pinned ring of REUSE_RING records (non-blocking feedback for the schedule)
  bool reuse_on = false;      // SPHB_FLAG_REUSE_LISTS in sphb_params.flags, or SPHB_REUSE=1 / SPHB_REUSE_PERIOD in the environment
  int reuse_period = 2;       // (unused by the budget schedule; reported): one rebuild + (period - 1) reuse evaluations; 1 = never reuse
  int reuse_period_max = 12, reuse_period_fixed = 0;
  double reuse_skin = 0.25;   // extended candidates are collected up to h (1 + skin)
  int reuse_ncw = 416;       // staged slots per tile of a rebuild that starts a cycle (<= 512: slots are 9-bit in the annulus pass)
  ReuseState* force_rs = nullptr;  // arguments of the force launch in progress
  bool force_stale = false;
  bool cell_per_h_fixed = false;   // SPHB_CELL_PER_H given: no adjustment for extended searches
  bool lists_ext = false;     // nn / nx / dexcl / rs describe the current particles (order and displacement chain)
  int reuse_age = 0;          // reuse evaluations since the rebuild
  sphb_params list_prm{};     // parameters of the rebuild (a change invalidates the displacement bookkeeping)
  RingState ring;
  bool interleaved = false;   // ring, inside a cycle: ghosts sit between the owned particles (sorted order)
  uint32_t stat_enq = 0, stat_seen = 0;  // records enqueued / read back
  cudaEvent_t stat_event = nullptr;      // recorded behind the last record's copy
  bool stat_event_valid = false;
  bool reuse_local = false;              // SPHB_REUSE_LOCAL=1: certificates may use the local displacement bound (measured: the
                                         // cycles get 40 % longer, the bookkeeping costs as much - sphb_reuse.cuh)
  bool reuse_abort = false;              // the policy ended the current cycle: the next evaluation rebuilds
  int search_level = 0, search_calm = 0; // width of the tile search (margin / column capacity), adapted from its refusals
  bool search_level_fixed = false;       // SPHB_GUESS_MARGIN / SPHB_KNN_CAP given
  bool no_record = false;                // SPHB_NO_RECORD=1: no feedback records (diagnostic)
  bool touched = false;                  // the caller changed the state (upload, append, parameters) since the last step
  int touched_streak = 0;                // consecutive steps that were preceded by such a change
  int calm_steps = 0;                    // consecutive rebuilds whose tile search refused < 0.1 % of the particles
  double reuse_kappa = 0.75;             // share of the skin the displacement bound may use up (adapts)
  double fb_D = 0.0, fb_D_prev = 0.0, fb_dy = 0.0, fb_last_frac = 0.0;  // last feedback record
  bool fb_valid = false;
  int reuse_cooldown = 0;
  std::string err;
};

#define REUSE_RING 64
#define REUSE_MIN_N (1 << 21)  // below this a step is short enough for the schedule's host wait to show

namespace {

int fail(sphb_sim* s, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (s) s->err = buf; else g_create_error = buf;
  return code;
}

#define CK(s, call)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail((s), SPHB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define CKL(s) CK(s, cudaGetLastError())

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

bool axis_open(const double a[2]) { return a[0] == SPHB_OPEN_LO; }

int check_params(sphb_sim* s, const sphb_params* p) {
  if (!p) return fail(s, SPHB_E_INVALID, "params is NULL");
  if (p->kernel < 0 || p->kernel > 2) return fail(s, SPHB_E_INVALID, "unknown kernel %d", p->kernel);
  if (p->precision != 64 && p->precision != 32) return fail(s, SPHB_E_INVALID, "precision %d: 64 (reference arithmetic) or 32 (fp32 pair arithmetic)", p->precision);
  // nearest-neighbour.go:44,53: an axis is either open on both ends or periodic
  if ((p->hor[0] == SPHB_OPEN_LO) != (p->hor[1] == SPHB_OPEN_HI) && p->hor[0] == SPHB_OPEN_LO)
    return fail(s, SPHB_E_INVALID, "cannot have open and periodic boundary in horizontal at same time");
  if ((p->ver[0] == SPHB_OPEN_LO) != (p->ver[1] == SPHB_OPEN_HI) && p->ver[0] == SPHB_OPEN_LO)
    return fail(s, SPHB_E_INVALID, "cannot have open and periodic boundary in vertical at same time");
  if (!axis_open(p->hor) && !(p->hor[1] > p->hor[0])) return fail(s, SPHB_E_INVALID, "empty horizontal period");
  if (!axis_open(p->ver) && !(p->ver[1] > p->ver[0])) return fail(s, SPHB_E_INVALID, "empty vertical period");
  return SPHB_OK;
}

PhysP make_phys(const sphb_params& p, int kernel) {
  PhysP ph;
  ph.dtH = p.dt_half; ph.gamma = p.gamma; ph.mass = p.particle_mass; ph.gx = p.accel[0]; ph.gy = p.accel[1];
  ph.hor0 = p.hor[0]; ph.hor1 = p.hor[1]; ph.ver0 = p.ver[0]; ph.ver1 = p.ver[1];
  ph.rL = p.refl_L; ph.rR = p.refl_R; ph.rU = p.refl_U; ph.rD = p.refl_D;
  // Go untyped-constant expressions rounded once (sph.go:249,261,275,293,303)
  const double MONAGHAN = 0x1.5d3b3e3583243p+3;  // 6*40/(pi*7)
  const double WEND_F = 0x1.1d34a60108f72p+1;    // 4*7/(pi*4)
  const double WEND_DF = 0x1.1d34a60108f72p+2;   // 8*7/(pi*4)
  const double TOPHAT = 0x1.45f306dc9c883p-2;    // 1/pi
  ph.Fpref = kernel == 0 ? TOPHAT : (kernel == 1 ? MONAGHAN : WEND_F);
  ph.DFpref = kernel == 0 ? 1.0 : (kernel == 1 ? MONAGHAN : WEND_DF);
  ph.cfac = p.gamma * (p.gamma - 1.0);
  ph.kernel = kernel;
  return ph;
}

template <typename T>
cudaError_t dalloc(T*& p, size_t count) {
  return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
}

cudaError_t alloc_soa(Soa& x, size_t cap) {
  cudaError_t e;
  if ((e = dalloc(x.pos, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.vel, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.vdot, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.vpred, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.e, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.edot, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.epred, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.id, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.pc, cap)) != cudaSuccess) return e;
  if ((e = dalloc(x.ghost, cap)) != cudaSuccess) return e;
  return cudaSuccess;
}
void free_soa(Soa& x) {
  cudaFree(x.pos); cudaFree(x.vel); cudaFree(x.vdot); cudaFree(x.vpred);
  cudaFree(x.e); cudaFree(x.edot); cudaFree(x.epred); cudaFree(x.id); cudaFree(x.pc); cudaFree(x.ghost);
  x = Soa{};
}

void zero_soa_range(sphb_sim* s, Soa& x, int64_t off, int64_t cnt) {
  cudaMemsetAsync(x.pos + off, 0, cnt * sizeof(double2), s->st);
  cudaMemsetAsync(x.vel + off, 0, cnt * sizeof(double2), s->st);
  cudaMemsetAsync(x.vdot + off, 0, cnt * sizeof(double2), s->st);
  cudaMemsetAsync(x.vpred + off, 0, cnt * sizeof(double2), s->st);
  cudaMemsetAsync(x.e + off, 0, cnt * sizeof(double), s->st);
  cudaMemsetAsync(x.edot + off, 0, cnt * sizeof(double), s->st);
  cudaMemsetAsync(x.epred + off, 0, cnt * sizeof(double), s->st);
  cudaMemsetAsync(x.pc + off, 0, cnt * sizeof(double4), s->st);
  cudaMemsetAsync(x.ghost + off, 0, cnt * sizeof(uint8_t), s->st);
}

int ensure_scratch(sphb_sim* s, size_t bytes) {
  if (bytes <= s->scratchBytes) return SPHB_OK;
  CK(s, cudaStreamSynchronize(s->st));
  cudaFree(s->scratch);
  s->scratch = nullptr; s->scratchBytes = 0;
  CK(s, cudaMalloc(&s->scratch, bytes));
  s->scratchBytes = bytes;
  return SPHB_OK;
}

int enter(sphb_sim* s) {
  if (!s) return SPHB_E_INVALID;
  CK(s, cudaSetDevice(s->device));
  return SPHB_OK;
}

// the state changed under the lists (upload, append, parameters): the next evaluation is a rebuild
void invalidate_reuse(sphb_sim* s) { s->lists_ext = false; s->reuse_age = 0; }

// grid parameters of an evaluation; ext: its tile search collects candidates up to h (1 + skin), and three rows of
// cells must cover that reach
GridTune grid_tune(const sphb_sim* s, bool ext) {
  GridTune t = s->gtune;
  if (ext && !s->cell_per_h_fixed) t.cell_per_h = std::max(t.cell_per_h, (1.0 + s->reuse_skin) * 1.016);
  return t;
}

// the force epilogue left the next step's cell counts in cellCount: forget them (the state changed under them)
void drop_ready_keys(sphb_sim* s) {
  if (s->keys_ready) cudaMemsetAsync(s->cellCount, 0, ((size_t)s->ntiles_cap * SC_TILE + 8) * sizeof(uint32_t), s->st);
  s->keys_ready = false;
}

// ---- one force evaluation -----------------------------------------------------------------------
enum { MODE_ASIS = 0, MODE_INIT = 1, MODE_DRIFT = 2 };

int refresh_stats(sphb_sim* s) {
  const int ntot = (int)(s->n + s->nghost);
  const int nb = ntot > 0 ? std::min(STAT_BLOCKS, cdiv(ntot, 256)) : 1;
  k_stats_partial<<<nb, 256, 0, s->st>>>(s->a.pos, s->a.vel, s->a.pc, s->a.e, s->interleaved ? ntot : (int)s->n, s->statPart,
                                         s->interleaved ? s->a.ghost : nullptr);
  k_stats_final<<<1, 32 * STAT_N, 0, s->st>>>(s->statPart, nb, s->stats);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 2;
  CKL(s);
  s->stats_dirty = false;
  return SPHB_OK;
}

// search parameters at the current width level (reuse_policy adapts it from the refusals of the previous rebuilds)
KnnTune search_tune(const sphb_sim* s) {
  KnnTune kt = s->ktune;
  if (s->search_level >= 1) { kt.guess_margin = s->search_level == 1 ? 0.04 : 0.06; kt.cap = std::max(kt.cap, 56); }
  return kt;
}

KnnExt make_ext(sphb_sim* s, bool on) {
  return KnnExt{on ? s->nx : nullptr, on ? s->tinfo : nullptr, on ? s->ptab : nullptr, on ? s->dexcl : nullptr, s->reuse_skin};
}

template <int KERNEL, bool F32, bool EXT>
void launch_knn_p(sphb_sim* s, int ntot, const PhysP& ph) {
  KnnOut out{s->hacc, s->hscale, s->qmax, s->a.pc, s->nn, s->failList, s->failCount};
  const int tiles = cdiv(ntot, 32);
  KnnTune kt = search_tune(s);
  kt.cap = s->have_h ? kt.cap : s->ktune.cap0;  // first evaluation: radius from a density estimate, wider spread
  kt.ncw = s->have_h ? (EXT ? s->reuse_ncw : s->ktune.ncw) : s->ktune.ncw0;
  const size_t smem = (size_t)KNN_WARPS * knn_smem_bytes_per_warp(kt.cap, kt.ncw, F32);
  static bool attr_done[64] = {};  // function attributes are per device (one handle per GPU, maybe several per process)
  if (!attr_done[s->device & 63]) {
    cudaFuncSetAttribute(k_knn_tile<KERNEL, F32, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done[s->device & 63] = true;
  }
  k_knn_tile<KERNEL, F32, EXT><<<cdiv(tiles, KNN_WARPS), KNN_THREADS, smem, s->st>>>(s->spos, s->keysSorted, s->cellStart,
                                                                                   s->hguess, s->a.epred, ntot, s->grid, ph, kt, out,
                                                                                   s->slab_on ? s->a.ghost : nullptr, s->dflags, make_ext(s, EXT));
}

// ext: the evaluation starts a reuse cycle: the tile search keeps its staged blocks and the candidates' slots, the
// annulus pass adds the further candidates and the exclusion radii
template <int KERNEL>
void launch_knn(sphb_sim* s, int ntot, const PhysP& ph, bool ext) {
  ext = ext && s->have_h;  // (the first evaluation has no previous h to take the skin from)
  if (s->prm.precision == 32) { if (ext) launch_knn_p<KERNEL, true, true>(s, ntot, ph); else launch_knn_p<KERNEL, true, false>(s, ntot, ph); }
  else { if (ext) launch_knn_p<KERNEL, false, true>(s, ntot, ph); else launch_knn_p<KERNEL, false, false>(s, ntot, ph); }
  KnnOut out{s->hacc, s->hscale, s->qmax, s->a.pc, s->nn, s->failList, s->failCount};
  FbExt fx{ext ? s->dexcl : nullptr, nullptr};
  const uint8_t* gf = s->slab_on ? s->a.ghost : nullptr;
  k_knn_fallback<KERNEL, false><<<148 * 4, 128, 0, s->st>>>(s->spos, s->keysSorted, s->cellStart, s->hguess,
                                                          s->a.epred, ntot, s->grid, ph, out, gf, s->dflags, fx);
  if (ext) {
    KnnTune kt = search_tune(s);
    kt.ncw = s->reuse_ncw;
    static bool attr_done[64] = {};
    if (!attr_done[s->device & 63]) {
      cudaFuncSetAttribute(k_knn_annulus, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      attr_done[s->device & 63] = true;
    }
    k_knn_annulus<<<cdiv(cdiv(ntot, 32), KNN_WARPS), KNN_THREADS, (size_t)KNN_WARPS * annulus_smem_bytes_per_warp(kt.ncw), s->st>>>(
        s->spos, s->keysSorted, s->cellStart, s->hguess, s->a.pc, ntot, s->grid, kt, make_ext(s, true), gf, s->dflags);
    s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  }
  s->have_h = true;
  s->lists_ext = ext;  // (forces_plan completes the bookkeeping; any other caller leaves plain lists)
}

// REUSE evaluation: exact kNN from the stored candidates, refused particles to the ring-expansion search on the stale cells
template <int KERNEL>
void launch_knn_reuse(sphb_sim* s, int ntot, const PhysP& ph) {
  KnnOut out{s->hacc, s->hscale, s->qmax, s->a.pc, s->nn, s->failList, s->failCount};
  const bool f32 = s->prm.precision == 32;
  const size_t smem = (size_t)REUSE_NC * REUSE_THREADS * (f32 ? 4 : 8);
  static bool attr_done[64] = {};
  if (!attr_done[s->device & 63]) {
    cudaFuncSetAttribute(k_knn_reuse<KERNEL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_knn_reuse<KERNEL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_done[s->device & 63] = true;
  }
  const uint8_t* gf = s->slab_on ? s->a.ghost : nullptr;
  const float2* uc = (s->slab_on || !s->reuse_local) ? nullptr : s->ucum;  // (a ring has no displacement sums for its ghosts)
  if (f32) k_knn_reuse<KERNEL, true><<<cdiv(ntot, REUSE_THREADS), REUSE_THREADS, smem, s->st>>>(s->spos, s->a.epred, ntot, s->grid, ph, out, s->nx, s->dexcl, s->rs, gf, s->dflags, s->keysSorted, uc, s->ubox, s->ubox_ok);
  else k_knn_reuse<KERNEL, false><<<cdiv(ntot, REUSE_THREADS), REUSE_THREADS, smem, s->st>>>(s->spos, s->a.epred, ntot, s->grid, ph, out, s->nx, s->dexcl, s->rs, gf, s->dflags, s->keysSorted, uc, s->ubox, s->ubox_ok);
  FbExt fx{s->dexcl, s->rs};
  k_knn_fallback<KERNEL, true><<<148 * 4, 128, 0, s->st>>>(s->spos, s->keysSorted, s->cellStart, s->hguess,
                                                         s->a.epred, ntot, s->grid, ph, out, gf, s->dflags, fx);
}

SlabP make_slabp_of(const sphb_sim* s, const sphb_slab& b) {
  SlabP sl{};
  sl.x_lo = b.x_lo; sl.x_hi = b.x_hi; sl.ghost_w = b.ghost_w; sl.inner_w = b.inner_w;
  sl.has_left = b.has_left; sl.has_right = b.has_right;
  if (s->slab_on && !axis_open(s->prm.hor)) {  // image frame centred on the slab
    sl.Lx = s->prm.hor[1] - s->prm.hor[0];
    sl.frame_lo = 0.5 * (b.x_lo + b.x_hi) - 0.5 * sl.Lx;
  }
  return sl;
}
SlabP make_slabp(const sphb_sim* s) { return make_slabp_of(s, s->slab); }
bool same_slab(const sphb_slab& a, const sphb_slab& b) {
  return a.x_lo == b.x_lo && a.x_hi == b.x_hi && a.ghost_w == b.ghost_w && a.inner_w == b.inner_w && a.has_left == b.has_left &&
         a.has_right == b.has_right;
}

template <int KERNEL, bool INTEGRATE, bool SLAB, typename R>
void launch_force_st(sphb_sim* s, const ForceIO& io, int ntot, const PhysP& ph) {
  constexpr bool F32 = sizeof(R) == 4;
  static bool attr_done[64] = {};  // per device
  if (!attr_done[s->device & 63]) {
    if (F32) cudaFuncSetAttribute(k_force_st32<KERNEL, INTEGRATE, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    else cudaFuncSetAttribute(k_force_st<KERNEL, INTEGRATE, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done[s->device & 63] = true;
  }
  const int nrec = s->force_nrec;
  const size_t smem = (size_t)nrec * 8 * sizeof(R);  // four arrays of two reals per staged record
  const int nb = cdiv(ntot, FORCE_THREADS);
  if (F32) k_force_st32<KERNEL, INTEGRATE, SLAB><<<nb, FORCE_THREADS, smem, s->st>>>(io, ntot, s->grid, ph, nrec, s->dflags);
  else k_force_st<KERNEL, INTEGRATE, SLAB><<<nb, FORCE_THREADS, smem, s->st>>>(io, ntot, s->grid, ph, nrec, s->dflags);
}

template <int KERNEL, bool INTEGRATE, bool SLAB>
void launch_force_p(sphb_sim* s, const ForceIO& io, int ntot, const PhysP& ph) {
  if (s->prm.precision == 32) launch_force_st<KERNEL, INTEGRATE, SLAB, float>(s, io, ntot, ph);
  else launch_force_st<KERNEL, INTEGRATE, SLAB, double>(s, io, ntot, ph);
}

template <int KERNEL>
void launch_force(sphb_sim* s, int ntot, const PhysP& ph, bool integrate) {
  ForceIO io{};
  io.spos = s->spos; io.vpred = s->a.vpred; io.pc = s->a.pc; io.nn = s->nn;
  io.keys = s->keysSorted; io.cellStart = s->cellStart; io.qmax = s->qmax;
  io.next_grid = s->fuse_keys ? s->grid_next : nullptr;
  io.next_keys = s->keys; io.next_rank = s->rank; io.next_count = s->cellCount;
  io.rs = s->force_rs; io.stale = s->force_stale ? s->rs : nullptr;
  io.ucum = (s->force_rs && !s->slab_on && s->reuse_local) ? s->ucum : nullptr; io.ucum_reset = s->force_stale ? 0 : 1;
  if (integrate && !(s->slab_on && s->force_stale)) cudaMemsetAsync(s->qmax + 1, 0, sizeof(uint32_t), s->st);  // max |v|^2 after this kick (ring: of the cycle)
  io.pos = s->a.pos; io.vel = s->a.vel; io.e = s->a.e; io.vdot = s->a.vdot; io.edot = s->a.edot;
  if (!s->slab_on) {
    if (integrate) launch_force_p<KERNEL, true, false>(s, io, ntot, ph);
    else launch_force_p<KERNEL, false, false>(s, io, ntot, ph);
    return;
  }
  io.gflag = s->a.ghost;
  if (integrate) launch_force_p<KERNEL, true, true>(s, io, ntot, ph);
  else launch_force_p<KERNEL, false, true>(s, io, ntot, ph);
}

// keep the entries of [0, nslots) flagged GF_OWNED, in place, in [0, nkeep) (nkeep = their number); failList / perm
// are free outside the kNN / sort phases and serve as the hole / filler lists.  Cell keys the force epilogue has already
// produced for the next step (keys_ready) move with their particles.
void compact_in_place(sphb_sim* s, int nslots, int nkeep) {
  if (nslots <= 0) return;
  cudaMemsetAsync(s->packCount, 0, 2 * sizeof(int), s->st);
  k_find_holes<<<cdiv(nslots, 256), 256, 0, s->st>>>(s->a.ghost, nslots, nkeep, s->failList, s->perm, s->packCount);
  SoaPtr a{s->a.pos, s->a.vel, s->a.vdot, s->a.vpred, s->a.e, s->a.edot, s->a.epred, s->a.id, s->a.pc, s->a.ghost};
  const int moved_max = std::max(1, nslots - nkeep);  // fillers sit in [nkeep, nslots)
  k_fill_holes<<<std::min(cdiv(moved_max, 256), 148 * 8), 256, 0, s->st>>>(a, s->failList, s->perm, s->packCount, s->dflags,
                                                                          s->keys_ready ? s->keys : nullptr, s->rank);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 2;
}

// sort + reorder + kNN (+ density with `kernel`); the neighbour list, spos and grid then describe s->a
// ext: start a reuse cycle (extended lists); want_hacc: the smoothing-length accumulator will be consumed by the next
// evaluation's grid (not when a reuse evaluation follows: it needs no grid)
int build_neighbours(sphb_sim* s, int mode, const double hor[2], const double ver[2], int kernel, bool timed, bool ext = false,
                     bool want_hacc = true, uint32_t* inv = nullptr) {
  const int ntot = (int)(s->n + s->nghost);
  if (ntot <= 0) return fail(s, SPHB_E_STATE, "Simulation not initialized: no particles (sph.go:92-94)");
  // Fully periodic runs (single handle, or a slab of a periodic ring) need no statistics pass per evaluation: the box
  // comes from hor / ver / the slab edges, the mean smoothing length from the accumulator the previous kNN filled,
  // max h and max speed (slab driver) from qmax.
  const bool periodic = !axis_open(ver) && !axis_open(hor) && (!s->slab_on || (s->slab.has_left && s->slab.has_right));
  invalidate_reuse(s);
  // a fused step already built this evaluation's grid (and consumed the accumulator for it)
  const bool same_box = s->grid_next_ready && periodic && hor[0] == s->next_hor[0] && hor[1] == s->next_hor[1] &&
                        ver[0] == s->next_ver[0] && ver[1] == s->next_ver[1] &&
                        (s->slab_on ? (s->keys_slab.has_left && same_slab(s->slab, s->keys_slab)) : !s->keys_slab.has_left);
  if (s->grid_next_ready && !same_box) s->hacc_valid = false;  // the accumulator went into a grid that does not apply
  s->grid_next_ready = false;
  const bool use_hacc = periodic && s->hacc_valid;
  if (s->stats_dirty && !use_hacc) { int rc = refresh_stats(s); if (rc) return rc; }
  const double hscale_prev = s->hscale;
  s->hscale = periodic && want_hacc ? 16777216.0 / std::max(hor[1] - hor[0], ver[1] - ver[0]) : 0.0;  // h < L: 24-bit fixed point
  const PhysP ph = make_phys(s->prm, kernel);
  const double dtH = s->prm.dt_half;
  if (timed) cudaEventRecord(s->ev[SPHB_PH_KEYS], s->st);
  if (same_box) std::swap(s->grid, s->grid_next);
  else k_make_grid<<<1, 32, 0, s->st>>>(s->stats, ntot, hor[0], hor[1], ver[0], ver[1], make_slabp(s), s->slab_on ? 1 : 0,
                                        grid_tune(s, ext && s->have_h), s->grid, s->hacc, hscale_prev > 0.0 ? hscale_prev : 1.0, use_hacc ? 1 : 0);
  s->hacc_valid = periodic && want_hacc;  // the kNN below refills the accumulator
  cudaMemsetAsync(s->qmax, 0, sizeof(uint32_t), s->st);  // max h of this evaluation
  const bool keys_ok = same_box && s->keys_ready && mode == MODE_DRIFT && dtH == s->keys_dtH && (int)s->n == s->keys_n;
  if (keys_ok) {
    s->keys_ready = false;  // consumed: the scan below zeroes cellCount again
    if (s->nghost) {        // slab mode: the owned particles have their keys, the ghosts have just arrived
      const size_t o = (size_t)s->n;
      k_keys<true><<<cdiv((int)s->nghost, 256), 256, 0, s->st>>>(s->a.pos + o, s->a.vel + o, (int)s->nghost, s->grid, dtH, s->keys + o, s->rank + o, s->cellCount);
      s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
    }
  } else {
    drop_ready_keys(s);
    if (mode == MODE_DRIFT) k_keys<true><<<cdiv(ntot, 256), 256, 0, s->st>>>(s->a.pos, s->a.vel, ntot, s->grid, dtH, s->keys, s->rank, s->cellCount);
    else k_keys<false><<<cdiv(ntot, 256), 256, 0, s->st>>>(s->a.pos, s->a.vel, ntot, s->grid, dtH, s->keys, s->rank, s->cellCount);
  }
  if (timed) cudaEventRecord(s->ev[SPHB_PH_SORT], s->st);
  // counting sort: scan of the cell counts, then the permutation
  k_scan_tiles<<<s->ntiles_cap, SC_THREADS, 0, s->st>>>(s->cellCount, s->grid, s->tileSum);
  k_excl_scan<<<1, 1024, 0, s->st>>>(s->tileSum, s->ntiles_cap, s->grid);
  k_scan_apply<<<s->ntiles_cap, SC_THREADS, 0, s->st>>>(s->cellCount, s->grid, s->tileSum, s->cellStart);
  k_scatter_perm<<<cdiv(ntot, 256), 256, 0, s->st>>>(s->keys, s->rank, s->cellStart, ntot, s->perm);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 4 + (same_box ? 0 : 1) + (keys_ok ? 0 : 1);
  if (timed) cudaEventRecord(s->ev[SPHB_PH_REORDER], s->st);
  StateIn in{s->a.pos, s->a.vel, s->a.vdot, s->a.vpred, s->a.e, s->a.edot, s->a.epred, s->a.id, s->a.pc, s->a.ghost};
  StateOut out{s->b.pos, s->b.vel, s->b.vdot, s->b.vpred, s->b.e, s->b.edot, s->b.epred, s->b.id, s->b.pc, s->b.ghost, s->spos, s->hguess};
  const int rb = cdiv(ntot, 256);
#define REORDER(MODE, LEAN) k_reorder<MODE, LEAN><<<rb, 256, 0, s->st>>>(in, out, s->keys, s->perm, ntot, s->grid, dtH, s->cellStart, s->keysSorted, inv)
  // `timed` = called from forces(): a force evaluation follows and overwrites VDot, EDot; kNN rewrites {rho, c, h, P}
  if (timed) {
    if (mode == MODE_DRIFT) REORDER(2, true); else if (mode == MODE_INIT) REORDER(1, true); else REORDER(0, true);
  } else {
    if (mode == MODE_DRIFT) REORDER(2, false); else if (mode == MODE_INIT) REORDER(1, false); else REORDER(0, false);
  }
#undef REORDER
  std::swap(s->a, s->b);
  cudaMemsetAsync(s->failCount, 0, sizeof(int), s->st);
  if (timed) cudaEventRecord(s->ev[SPHB_PH_KNN], s->st);
  switch (kernel) {
    case 0: launch_knn<0>(s, ntot, ph, ext); break;
    case 1: launch_knn<1>(s, ntot, ph, ext); break;
    default: launch_knn<2>(s, ntot, ph, ext); break;
  }
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 3;
  CKL(s);
  s->have_list = true;
  s->knn_hor[0] = hor[0]; s->knn_hor[1] = hor[1]; s->knn_ver[0] = ver[0]; s->knn_ver[1] = ver[1];
  return SPHB_OK;
}

// ---- schedule of the list reuse ------------------------------------------------------------------
// Every step ends with k_reuse_update, which leaves a record {age, refused particles, D, grid row height} in a pinned
// host ring.  Before it plans an evaluation the host waits for the record of the previous one (one kernel-launch
// latency per step: the device has nothing else queued) and decides:
//   - a reuse evaluation is planned only while D, the displacement bound the certificate will see (it is in the
//     record: exact, not predicted), stays below  kappa x skin x mean h  - all certificates fail together once D eats
//     the skin, and a refused particle costs ~50 accepted ones, so the cycle must end before that cliff;
//   - kappa adapts: a cycle that ended on this budget with < 0.15 % refused in its last evaluation raises it by 0.05
//     (to 0.9 at most), an evaluation that refused > 0.5 % lowers it by 0.1 (0.2 at least) and ends the cycle at once;
//   - no cycle starts unless the tile search itself refused less than 0.1 % of the particles in the last two rebuilds
//     (i.i.d. clouds, free surfaces: their smoothing lengths change by more than the search margin per step);
//   - handles below 2^21 particles do not reuse (the host wait of the schedule shows in steps that short).
// SPHB_REUSE_PERIOD = p fixes cycles of p evaluations instead (tests, sweeps).
struct ReuseFeedback { unsigned age; double frac, D, dy; bool rebuild; };

void reuse_policy(sphb_sim* s, const ReuseFeedback& f) {
  s->fb_D_prev = (!f.rebuild && f.age > 0) ? s->fb_D : 0.0;
  s->fb_D = f.D; s->fb_dy = f.dy; s->fb_valid = true;
  if (f.rebuild) {  // what the tile search itself refused: a flow that upsets even the full search is no candidate for reuse
    s->calm_steps = f.frac < 1e-3 ? s->calm_steps + 1 : 0;
    // Width of the tile search.  Its radius is h_prev (1 + margin) and its column holds cap entries: on smooth flows
    // 2 % / 47 is the fastest (fewer candidates to select from), but where h changes by several per cent per step
    // (i.i.d. clouds, fresh simulations) the search comes up short for many particles, and each of those costs a
    // warp-wide ring search.  Measured on 2^20 i.i.d. particles: 12.9 % refused at 2 % / 47 (1.55 ms per step), 1.6 % at
    // 4 % / 56 (0.89 ms), 0.14 % at 6 % / 56 (0.85 ms); the jittered lattice of the same size pays 8 % for the wide search.
    if (!s->search_level_fixed) {
      if (f.frac > 1e-2 && s->search_level < 2) { s->search_level = f.frac > 5e-2 ? 2 : s->search_level + 1; s->search_calm = 0; }
      else if (f.frac < 2e-4) { if (++s->search_calm >= 16 && s->search_level > 0) { s->search_level -= 1; s->search_calm = 0; } }
      else s->search_calm = 0;
    }
    return;
  }
  if (s->reuse_period_fixed) return;
  if (f.frac > 5e-3) {
    s->reuse_kappa = std::max(0.2, s->reuse_kappa - 0.1);
    s->reuse_abort = true;
  }
  s->fb_last_frac = f.frac;
}

// may an evaluation that sees the displacement bound D reuse the lists?
bool reuse_budget_ok(const sphb_sim* s, double D) {
  if (!s->fb_valid || !(s->fb_dy > 0.0)) return false;
  const double hmean = s->fb_dy / grid_tune(s, true).cell_per_h;  // the grid rows follow the mean h
  return D <= s->reuse_kappa * s->reuse_skin * hmean;
}

// block: wait for the record of the previous evaluation (inside a cycle the plan needs its D); otherwise take what has
// arrived - outside a cycle the records only count calm rebuilds, and steps keep being enqueued without a host wait
void reuse_poll(sphb_sim* s, bool block) {
  if (!s->stat_host) return;
  if (block && s->stat_seen < s->stat_enq && s->stat_event_valid) cudaEventSynchronize(s->stat_event);
  while (s->stat_seen < s->stat_enq) {
    const volatile ReuseStat* r = &s->stat_host[(s->stat_seen + 1) % REUSE_RING];
    if (r->seq != s->stat_seen + 1 || r->seq2 != r->seq) {
      if (s->stat_enq - s->stat_seen >= REUSE_RING) { s->stat_seen = s->stat_enq - REUSE_RING / 2; continue; }  // overwritten
      break;  // not there yet
    }
    s->stat_seen += 1;
    if (r->n == 0) continue;
    reuse_policy(s, ReuseFeedback{r->age, (double)r->refused / (double)r->n, (double)r->D, (double)r->dy, r->rebuild != 0});
  }
}

// the plan of an ordinary step from the feedback: {reuse now, may the next evaluation reuse}
void reuse_plan(sphb_sim* s, bool lists_ok, bool& reuse, bool& next_reuse) {
  const bool calm = s->reuse_period_fixed || s->calm_steps >= 2;
  if (s->reuse_period_fixed) {
    const int period = s->reuse_period_fixed;
    reuse = lists_ok && s->reuse_age + 1 < period;
    next_reuse = reuse ? s->reuse_age + 2 < period : (period > 1 && s->have_h);
    return;
  }
  const int amax = s->reuse_period_max;
  reuse = lists_ok && !s->reuse_abort && s->reuse_age + 1 < amax && reuse_budget_ok(s, s->fb_D);
  if (lists_ok && !reuse && !s->reuse_abort && s->reuse_age > 0 && s->fb_last_frac < 1.5e-3)  // the budget ended a clean cycle
    s->reuse_kappa = std::min(s->reuse_local && !s->slab_on ? 2.5 : 0.9, s->reuse_kappa + (s->reuse_local && !s->slab_on ? 0.15 : 0.05));
  s->reuse_abort = false;
  // next evaluation: the bound grows by about what it grew last time
  const double growth = s->fb_D - s->fb_D_prev;
  if (reuse) next_reuse = s->reuse_age + 2 < amax && reuse_budget_ok(s, s->fb_D + 1.1 * growth);
  else next_reuse = s->have_h && calm && amax > 1 && s->reuse_kappa > 0.0;
}

bool same_params(const sphb_params& a, const sphb_params& b) { return std::memcmp(&a, &b, sizeof(sphb_params)) == 0; }

void launch_reuse_update(sphb_sim* s, int ntot, bool rebuild) {
  const sphb_params& p = s->prm;
  double cs = 1.0;  // magnitude of the coordinates (rounding slack of the displacement bound)
  if (!axis_open(p.hor)) cs = std::max(cs, std::max(std::fabs(p.hor[0]), std::fabs(p.hor[1])));
  if (!axis_open(p.ver)) cs = std::max(cs, std::max(std::fabs(p.ver[0]), std::fabs(p.ver[1])));
  k_reuse_update<<<1, 32, 0, s->st>>>(s->rs, s->grid, ntot, rebuild ? 1 : 0, s->failCount, s->hacc, s->hscale, cs, s->stat_dev);
  s->stat_enq += 1;
  cudaMemcpyAsync(&s->stat_host[s->stat_enq % REUSE_RING], s->stat_dev, sizeof(ReuseStat), cudaMemcpyDeviceToHost, s->st);
  if (!s->stat_event_valid) s->stat_event_valid = cudaEventCreateWithFlags(&s->stat_event, cudaEventDisableTiming) == cudaSuccess;
  if (s->stat_event_valid) cudaEventRecord(s->stat_event, s->st);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
}

// what an evaluation does about the list reuse
struct EvalPlan {
  bool reuse = false;         // exact kNN from the stored candidates instead of sort + tile search
  bool next_reuse = false;    // the next evaluation may be a reuse evaluation: keep the displacement books (ring: and the ghosts)
  bool predicted = false;     // reuse: drift-1 + predict already ran (ring: before the halo is gathered)
  bool defer_update = false;  // the caller launches k_reuse_update itself (ring: after the all-reduce of the statistics)
  bool record = false;        // leave a feedback record even if no cycle is running (what the tile search refused)
};

// drop the ghosts (ring / slab mode): the few owned particles sorted behind index n move into the ghosts' slots
void drop_ghosts(sphb_sim* s) {
  if (s->nghost) compact_in_place(s, (int)(s->n + s->nghost), (int)s->n);
  s->nghost = 0;
  s->have_list = false;
  s->interleaved = false;
  invalidate_reuse(s);
}

// CalculateForces (sph.go:403-435) [+ kick, drift-2, wrap, reflections when integrate (sph.go:122-193)]
int forces_plan(sphb_sim* s, int mode, bool integrate, const EvalPlan& plan) {
  if (s->prm.kernel == SPHB_KERNEL_TOPHAT)
    return fail(s, SPHB_E_KERNEL, "TopHat2D.DF: not defined. derivative is delta distribution! (sph.go:251-253)");
  const int ntot = (int)(s->n + s->nghost);
  const PhysP ph = make_phys(s->prm, s->prm.kernel);
  const bool reuse = plan.reuse, next_reuse = plan.next_reuse;
  // the smoothing-length accumulator feeds the next rebuild's grid; a ring also takes its running max h / max speed from it
  const bool want_hacc = !next_reuse || s->slab_on;
  int rc;
  if (!reuse) {
    uint32_t* inv = nullptr;
    if (s->slab_on && s->ring.on && next_reuse) {
      if (s->ring.inv_cap < s->cap) {
        CK(s, cudaStreamSynchronize(s->st));
        cudaFree(s->ring.inv); s->ring.inv = nullptr; s->ring.inv_cap = 0;
        CK(s, dalloc(s->ring.inv, (size_t)s->cap));
        s->ring.inv_cap = s->cap;
      }
      inv = s->ring.inv;
    }
    rc = build_neighbours(s, mode, s->prm.hor, s->prm.ver, s->prm.kernel, true, next_reuse, want_hacc, inv);
    if (rc) return rc;
    s->lists_ext = s->lists_ext && next_reuse;
    s->reuse_age = 0;
    if (s->lists_ext) s->list_prm = s->prm;
  } else {
    if (ntot <= 0) return fail(s, SPHB_E_STATE, "Simulation not initialized: no particles (sph.go:92-94)");
    const bool periodic = !axis_open(s->prm.ver) && !axis_open(s->prm.hor) && (!s->slab_on || (s->slab.has_left && s->slab.has_right));
    s->grid_next_ready = false;
    s->hscale = periodic && want_hacc ? 16777216.0 / std::max(s->prm.hor[1] - s->prm.hor[0], s->prm.ver[1] - s->prm.ver[0]) : 0.0;
    s->hacc_valid = periodic && want_hacc;
    drop_ready_keys(s);
    if (!plan.predicted) cudaEventRecord(s->ev[SPHB_PH_KEYS], s->st);
    cudaEventRecord(s->ev[SPHB_PH_SORT], s->st);
    cudaEventRecord(s->ev[SPHB_PH_REORDER], s->st);
    Soa& a = s->a;
    if (!plan.predicted)
      k_predict<<<cdiv(ntot, 256), 256, 0, s->st>>>(a.pos, a.vel, a.vdot, a.e, a.edot, a.vpred, a.epred, s->spos, ntot, s->grid,
                                                    s->prm.dt_half, s->slab_on ? a.ghost : nullptr);
    cudaMemsetAsync(s->failCount, 0, sizeof(int), s->st);
    if (!s->slab_on) cudaMemsetAsync(s->qmax, 0, sizeof(uint32_t), s->st);  // (a ring keeps the max h of the whole cycle)
    cudaEventRecord(s->ev[SPHB_PH_KNN], s->st);
    if (s->prm.kernel == 1) launch_knn_reuse<1>(s, ntot, ph); else launch_knn_reuse<2>(s, ntot, ph);
    s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 3;
    s->counters[SPHB_CNT_REUSE_STEPS] += 1;
    s->reuse_age += 1;
    s->have_list = true;
    CKL(s);
  }
  cudaEventRecord(s->ev[SPHB_PH_FORCE], s->st);
  // Periodic single-handle steps: the next grid only depends on the mean h the kNN above produced, so it is built now
  // and the force epilogue emits the next step's cell keys (no k_keys pass in the next step).
  const bool slab_fuse = s->slab_on && s->next_slab_valid && s->next_slab.has_left && s->next_slab.has_right && !reuse;
  s->next_slab_valid = false;
  s->fuse_keys = integrate && (!s->slab_on || slab_fuse) && s->hacc_valid && !next_reuse && !axis_open(s->prm.hor) && !axis_open(s->prm.ver);
  if (s->fuse_keys) {
    // (the next evaluation is a rebuild; in an ordinary run of steps it starts a reuse cycle, i.e. searches with the skin)
    const bool next_ext = s->reuse_on && (s->reuse_period_fixed ? s->reuse_period_fixed > 1 : s->calm_steps >= 2);
    k_make_grid<<<1, 32, 0, s->st>>>(s->stats, ntot, s->prm.hor[0], s->prm.hor[1], s->prm.ver[0], s->prm.ver[1],
                                     s->slab_on ? make_slabp_of(s, s->next_slab) : make_slabp(s), s->slab_on ? 1 : 0,
                                     grid_tune(s, next_ext), s->grid_next, s->hacc, s->hscale, 1);
    s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  }
  s->force_rs = (next_reuse && s->lists_ext) ? s->rs : nullptr;
  s->force_stale = reuse;
  if (s->prm.kernel == 1) launch_force<1>(s, ntot, ph, integrate);
  else launch_force<2>(s, ntot, ph, integrate);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  if (s->force_rs && !s->slab_on && s->reuse_local) {  // bounding boxes of the cumulative displacements (local bound)
    k_ucum_bbox<<<148 * 4, 256, 0, s->st>>>(s->ucum, s->cellStart, s->grid, s->ubox, s->ubox_cap, s->ubox_ok);
    s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  }
  if (!plan.defer_update && (s->force_rs || reuse || plan.record)) launch_reuse_update(s, ntot, !reuse);  // (outside a cycle: for the record only)
  const bool keep = next_reuse && s->lists_ext;
  if (!keep) invalidate_reuse(s);  // the cycle ends here
  if (s->fuse_keys) {
    s->grid_next_ready = s->keys_ready = true;
    s->keys_dtH = s->prm.dt_half; s->keys_n = (int)s->n;  // (slab mode: the owned particles; ntot otherwise)
    s->keys_slab = s->slab_on ? s->next_slab : sphb_slab{};
    s->next_hor[0] = s->prm.hor[0]; s->next_hor[1] = s->prm.hor[1]; s->next_ver[0] = s->prm.ver[0]; s->next_ver[1] = s->prm.ver[1];
  }
  if (s->slab_on) {
    if (!keep) {  // drop the ghosts: the few owned particles sorted behind index n move into the ghosts' slots
      compact_in_place(s, ntot, (int)s->n);
      s->nghost = 0;
      s->have_list = false;
      s->interleaved = false;
    } else {
      s->interleaved = s->nghost > 0;  // reuse evaluations follow: the ghosts stay where the sort put them
    }
  }
  cudaEventRecord(s->ev[SPHB_PH_TOTAL], s->st);
  s->ev_valid = true;
  s->qmax_valid = s->hacc_valid;
  if (s->hacc_valid || next_reuse || reuse) s->stats_dirty = true;  // sums / bounds are refreshed on demand (sphb_reduce, open axes)
  else { rc = refresh_stats(s); if (rc) return rc; }
  CKL(s);
  return SPHB_OK;
}

// the schedule of a single handle (sphb_step); the slab protocol of include/sphb.h (sphb_slab_step_*) never reuses, the
// ring driver (sphb_ring.inc) plans for all ranks at once
int forces(sphb_sim* s, int mode, bool integrate) {
  EvalPlan plan;
  const bool cyc = s->reuse_on && !s->slab_on && mode == MODE_DRIFT && integrate &&
                   (s->reuse_period_fixed || s->n >= REUSE_MIN_N);  // an ordinary step of a handle worth the bookkeeping
  if (mode == MODE_DRIFT && integrate) {
    s->touched_streak = s->touched ? s->touched_streak + 1 : 0;
    s->touched = false;
  }
  const bool ordinary = !s->slab_on && mode == MODE_DRIFT && integrate;
  if (ordinary) reuse_poll(s, cyc && s->lists_ext);  // (outside a reuse cycle: whatever has arrived, no wait)
  if (cyc) {
    reuse_plan(s, s->lists_ext && same_params(s->prm, s->list_prm), plan.reuse, plan.next_reuse);
    // a caller that rewrites the state before every step (bench.py's e2e loop) would pay for extended lists it never uses
    if (!plan.reuse && s->touched_streak > 0) plan.next_reuse = false;
  }
  // ordinary steps leave a feedback record (search width, reuse schedule); small handles, whose steps are launch bound,
  // only every fourth step
  plan.record = ordinary && !s->no_record && (cyc || s->n >= (1 << 22) || (s->cur_step & 3) == 0);
  return forces_plan(s, mode, integrate, plan);
}

int check_async(sphb_sim* s) {
  uint32_t fl = 0;
  int fc[2] = {0, 0};
  CK(s, cudaMemcpyAsync(&fl, s->dflags, sizeof fl, cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaMemcpyAsync(fc, s->failCount, sizeof fc, cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  s->counters[SPHB_CNT_KNN_FALLBACK] = fc[1];
  if (fl) cudaMemsetAsync(s->dflags, 0, sizeof(uint32_t), s->st);
  if (fl & DFLAG_BUF_FULL) return fail(s, SPHB_E_NOMEM, "slab: a halo/migration buffer or the particle capacity is too small");
  if (fl & DFLAG_GHOST_THIN)
    return fail(s, SPHB_E_GHOST_THIN, "slab: a smoothing length reached past the ghost layer (inner_w / ghost_w too small)");
  if (fl & DFLAG_UNDERFULL) {
    return fail(s, SPHB_E_KNN_UNDERFULL,
                "kNN: fewer than 32 (particle, image) candidates exist for some particle; the reference would keep "
                "sentinel slots (nearest-neighbour.go:155-165)");
  }
  return SPHB_OK;
}

int upload_common(sphb_sim* s, int64_t off, int64_t n, const double* pos_xy, const double* vel_xy, const double* e,
                  const double* rho, const int64_t* id, cudaMemcpyKind kind, int64_t id_base) {
  zero_soa_range(s, s->a, off, n);
  CK(s, cudaMemcpyAsync(s->a.pos + off, pos_xy, n * sizeof(double2), kind, s->st));
  if (vel_xy) CK(s, cudaMemcpyAsync(s->a.vel + off, vel_xy, n * sizeof(double2), kind, s->st));
  if (e) CK(s, cudaMemcpyAsync(s->a.e + off, e, n * sizeof(double), kind, s->st));
  if (id) CK(s, cudaMemcpyAsync(s->a.id + off, id, n * sizeof(int64_t), kind, s->st));
  else k_iota64<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.id + off, (int)n, id_base);
  if (rho) {
    if (kind == cudaMemcpyHostToDevice) {
      int rc = ensure_scratch(s, n * sizeof(double)); if (rc) return rc;
      CK(s, cudaMemcpyAsync(s->scratch, rho, n * sizeof(double), kind, s->st));
      k_set_pc<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pc + off, (int)n, (const double*)s->scratch, 0);
    } else {
      k_set_pc<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pc + off, (int)n, rho, 0);
    }
  }
  CKL(s);
  CK(s, cudaStreamSynchronize(s->st));  // host buffers are borrowed for the duration of the call only
  s->stats_dirty = true;
  s->hacc_valid = false;  // new particles: the mean smoothing length must come from a statistics pass
  s->qmax_valid = false;
  drop_ready_keys(s);
  s->grid_next_ready = false;
  s->have_list = false;
  invalidate_reuse(s);
  s->touched = true;
  return SPHB_OK;
}

int create_common(const sphb_params* p, int64_t n, int64_t capacity, const double* pos_xy, const double* vel_xy,
                  const double* e, const double* rho, const int64_t* id, cudaMemcpyKind kind, sphb_sim** out) {
  if (!out) return fail(nullptr, SPHB_E_INVALID, "out is NULL");
  *out = nullptr;
  int rc = check_params(nullptr, p);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !pos_xy)) return fail(nullptr, SPHB_E_INVALID, "bad particle arrays");
  if (capacity < n) capacity = n;
  if (capacity < 1) capacity = 1;
  if (capacity >= (int64_t)IDX_MASK) return fail(nullptr, SPHB_E_NOMEM, "capacity %lld exceeds 2^28-1 particles per device", (long long)capacity);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, SPHB_E_CUDA, "no CUDA device (%s): libsphb has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (p->device < 0 || p->device >= ndev) return fail(nullptr, SPHB_E_INVALID, "device %d out of range (%d devices)", p->device, ndev);
  sphb_sim* s = new (std::nothrow) sphb_sim();
  if (!s) return fail(nullptr, SPHB_E_NOMEM, "out of host memory");
  s->prm = *p; s->device = p->device; s->n = n; s->cap = capacity;
#define CKC(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      int code = (e__ == cudaErrorMemoryAllocation) ? SPHB_E_NOMEM : SPHB_E_CUDA;                    \
      fail(nullptr, code, "%s failed: %s", #call, cudaGetErrorString(e__));                         \
      sphb_destroy(s);                                                                              \
      return code;                                                                                  \
    }                                                                                               \
  } while (0)
  CKC(cudaSetDevice(s->device));
  CKC(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
  const size_t cap = (size_t)capacity, cap32 = (cap + 31) / 32 * 32;
  CKC(alloc_soa(s->a, cap));
  CKC(alloc_soa(s->b, cap));
  CKC(dalloc(s->spos, cap));
  CKC(dalloc(s->hguess, cap));
  CKC(dalloc(s->keys, cap)); CKC(dalloc(s->keysSorted, cap)); CKC(dalloc(s->rank, cap)); CKC(dalloc(s->perm, cap));
  // cell table: about 2 particles per cell at most (the counting sort scans one counter per cell)
  int64_t ncm = std::max<int64_t>(64, std::min<int64_t>(capacity / 2 + 1, (int64_t)1 << 30));
  s->ncell_max = (int)ncm;
  s->ntiles_cap = cdiv(ncm + 1, SC_TILE);
  CKC(dalloc(s->cellStart, (size_t)s->ntiles_cap * SC_TILE + 8));
  CKC(dalloc(s->cellCount, (size_t)s->ntiles_cap * SC_TILE + 8));
  CKC(dalloc(s->tileSum, (size_t)s->ntiles_cap));
  CKC(cudaMemsetAsync(s->cellCount, 0, ((size_t)s->ntiles_cap * SC_TILE + 8) * sizeof(uint32_t), s->st));
  CKC(dalloc(s->nn, cap32 * SPHB_K));
  CKC(dalloc(s->nx, cap32 * SPHB_KX));
  CKC(dalloc(s->tinfo, cap32 / 32 + 1));
  CKC(dalloc(s->ptab, cap32 + 32));
  CKC(dalloc(s->dexcl, cap));
  CKC(dalloc(s->ucum, cap));
  s->ubox_cap = (int)(ncm / 8 + 1024);
  CKC(dalloc(s->ubox, (size_t)s->ubox_cap));
  CKC(dalloc(s->ubox_ok, 1));
  CKC(cudaMemsetAsync(s->ubox_ok, 0, sizeof(uint32_t), s->st));
  CKC(dalloc(s->rs, 1));
  CKC(cudaMemsetAsync(s->rs, 0, sizeof(ReuseState), s->st));
  CKC(dalloc(s->stat_dev, 1));
  CKC(cudaHostAlloc((void**)&s->stat_host, REUSE_RING * sizeof(ReuseStat), cudaHostAllocDefault));
  std::memset(s->stat_host, 0, REUSE_RING * sizeof(ReuseStat));
  CKC(dalloc(s->failList, cap));
  CKC(dalloc(s->failCount, 2));
  CKC(dalloc(s->packCount, 2));
  CKC(dalloc(s->qmax, 2));
  CKC(cudaMemsetAsync(s->qmax, 0, 2 * sizeof(uint32_t), s->st));
  CKC(dalloc(s->hacc, 2 * HACC_N));
  CKC(cudaMemsetAsync(s->hacc, 0, 2 * HACC_N * sizeof(unsigned long long), s->st));
  CKC(dalloc(s->dflags, 1));
  CKC(dalloc(s->statPart, (size_t)STAT_BLOCKS * STAT_N));
  CKC(dalloc(s->stats, STAT_N));
  CKC(dalloc(s->grid, 1));
  CKC(dalloc(s->grid_next, 1));
  for (auto& ev : s->ev) CKC(cudaEventCreate(&ev));
  CKC(cudaMemsetAsync(s->failCount, 0, 2 * sizeof(int), s->st));
  CKC(cudaMemsetAsync(s->dflags, 0, sizeof(uint32_t), s->st));
  CKC(cudaMemsetAsync(s->nn, 0xff, cap32 * SPHB_K * sizeof(uint32_t), s->st));
#undef CKC
  s->gtune.cell_per_h = 1.15;  // row height in units of the mean h: three rows cover +-rg
  s->gtune.aspect = 0.32;      // dx / dy: narrow cells keep the horizontal rounding waste small
  s->gtune.ppc0 = 4.0;
  s->gtune.ncell_max = s->ncell_max;
  s->gtune.force_nc = 0;
  s->ktune.guess_margin = 0.02;
  s->ktune.k_target = 46.0;
  // column of cap - 8 entries + 8 slack (overflow is checked once per 8 candidates, the compaction pads to 8); ncw staged candidates per
  // tile: a 32-particle strip of one row needs ~190 at 32 neighbours.  The fp64 build stages 28 B per candidate
  // and is occupancy-bound by shared memory: slightly tighter buffers buy two more warps per SM (measured
  // 9.6 -> 8.8 ms per evaluation at 2^25 particles for 0.05 % more fallback particles).
  s->ktune.cap = p->precision == 32 ? 50 : 47;
  s->ktune.cap0 = 80;
  s->ktune.ncw = p->precision == 32 ? 256 : 224;
  s->ktune.ncw0 = 512;
  s->reuse_on = (p->flags & SPHB_FLAG_REUSE_LISTS) != 0;
  if (const char* ev = getenv("SPHB_NO_RECORD")) s->no_record = atoi(ev) != 0;
  if (const char* ev = getenv("SPHB_REUSE_LOCAL")) s->reuse_local = atoi(ev) != 0;
  if (const char* ev = getenv("SPHB_REUSE_KAPPA")) s->reuse_kappa = atof(ev);
  if (const char* ev = getenv("SPHB_REUSE_PERIOD")) { s->reuse_period_fixed = std::max(1, std::min(64, atoi(ev))); s->reuse_on = true; }
  if (const char* ev = getenv("SPHB_REUSE")) s->reuse_on = atoi(ev) != 0;
  if (const char* ev = getenv("SPHB_REUSE_MAX")) s->reuse_period_max = std::max(1, std::min(64, atoi(ev)));
  if (const char* ev = getenv("SPHB_REUSE_SKIN")) s->reuse_skin = std::max(0.03, std::min(1.0, atof(ev)));
  if (const char* ev = getenv("SPHB_REUSE_NCW")) s->reuse_ncw = std::max(128, std::min(512, atoi(ev) / 32 * 32));
  if (const char* ev = getenv("SPHB_CELL_PER_H")) { s->gtune.cell_per_h = atof(ev); s->cell_per_h_fixed = true; }
  if (const char* ev = getenv("SPHB_PPC0")) s->gtune.ppc0 = atof(ev);
  if (const char* ev = getenv("SPHB_FORCE_NC")) s->gtune.force_nc = atoi(ev);
  if (const char* ev = getenv("SPHB_GUESS_MARGIN")) { s->ktune.guess_margin = atof(ev); s->search_level_fixed = true; }
  if (const char* ev = getenv("SPHB_K_TARGET")) s->ktune.k_target = atof(ev);
  if (const char* ev = getenv("SPHB_KNN_CAP")) { s->ktune.cap = std::max(44, std::min(96, atoi(ev))); s->search_level_fixed = true; }
  if (const char* ev = getenv("SPHB_KNN_NCW")) s->ktune.ncw = std::max(128, std::min(1024, atoi(ev) / 8 * 8));
  if (const char* ev = getenv("SPHB_CELL_ASPECT")) s->gtune.aspect = atof(ev);
  if (const char* ev = getenv("SPHB_FORCE_NREC")) s->force_nrec = std::max(64, std::min(3072, atoi(ev)));
  if (n > 0) {
    rc = upload_common(s, 0, n, pos_xy, vel_xy, e, rho, id, kind, 0);
    if (rc) { g_create_error = s->err; sphb_destroy(s); return rc; }
  }
  *out = s;
  return SPHB_OK;
}

// append beyond the capacity: Go's append reallocates (sph.go:79).  Every per-capacity array is allocated anew at twice
// the size BEFORE anything is released (a failed allocation leaves the handle as it was), the particle state is copied
// over, and everything that described the old order (neighbour list, prepared keys, next grid) is forgotten.
struct CapArrays {
  Soa a, b;
  double2* spos = nullptr;
  double* hguess = nullptr;
  uint32_t *keys = nullptr, *keysSorted = nullptr, *rank = nullptr, *perm = nullptr;
  uint32_t *cellStart = nullptr, *cellCount = nullptr, *tileSum = nullptr, *nn = nullptr;
  uint32_t* nx = nullptr;
  TileInfo* tinfo = nullptr;
  uint2* ptab = nullptr;
  double* dexcl = nullptr;
  float2* ucum = nullptr;
  float4* ubox = nullptr;
  int* failList = nullptr;
  void release() {
    free_soa(a); free_soa(b);
    cudaFree(nx); cudaFree(tinfo); cudaFree(ptab); cudaFree(dexcl); cudaFree(ucum); cudaFree(ubox);
    cudaFree(spos); cudaFree(hguess); cudaFree(keys); cudaFree(keysSorted); cudaFree(rank); cudaFree(perm);
    cudaFree(cellStart); cudaFree(cellCount); cudaFree(tileSum); cudaFree(nn); cudaFree(failList);
    *this = CapArrays{};
  }
};

int grow_capacity(sphb_sim* s, int64_t need) {
  const int64_t limit = (int64_t)IDX_MASK - 1;
  if (need > limit) return fail(s, SPHB_E_NOMEM, "append: %lld particles exceed 2^28-1 per device", (long long)need);
  const int64_t ncap = std::min<int64_t>(std::max<int64_t>(need, 2 * s->cap), limit);
  const size_t cap = (size_t)ncap, cap32 = (cap + 31) / 32 * 32;
  const int64_t ncm = std::max<int64_t>(64, std::min<int64_t>(ncap / 2 + 1, (int64_t)1 << 30));
  const int ntiles = cdiv(ncm + 1, SC_TILE);
  const size_t ncount = (size_t)ntiles * SC_TILE + 8;
  CK(s, cudaStreamSynchronize(s->st));
  CapArrays t;
#define CKG(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      t.release();                                                                                         \
      cudaGetLastError();                                                                                  \
      return fail(s, e__ == cudaErrorMemoryAllocation ? SPHB_E_NOMEM : SPHB_E_CUDA,                         \
                  "append: growing the capacity %lld -> %lld: %s failed: %s", (long long)s->cap, (long long)ncap, #call, \
                  cudaGetErrorString(e__));                                                                \
    }                                                                                                      \
  } while (0)
  CKG(alloc_soa(t.a, cap));
  CKG(alloc_soa(t.b, cap));
  CKG(dalloc(t.spos, cap));
  CKG(dalloc(t.hguess, cap));
  CKG(dalloc(t.keys, cap)); CKG(dalloc(t.keysSorted, cap)); CKG(dalloc(t.rank, cap)); CKG(dalloc(t.perm, cap));
  CKG(dalloc(t.cellStart, ncount));
  CKG(dalloc(t.cellCount, ncount));
  CKG(dalloc(t.tileSum, (size_t)ntiles));
  CKG(dalloc(t.nn, cap32 * SPHB_K));
  CKG(dalloc(t.nx, cap32 * SPHB_KX));
  CKG(dalloc(t.tinfo, cap32 / 32 + 1));
  CKG(dalloc(t.ptab, cap32 + 32));
  CKG(dalloc(t.dexcl, cap));
  CKG(dalloc(t.ucum, cap));
  CKG(dalloc(t.ubox, (size_t)(ncm / 8 + 1024)));
  CKG(dalloc(t.failList, cap));
  const size_t n = (size_t)s->n;
  const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
  CKG(cudaMemcpyAsync(t.a.pos, s->a.pos, n * sizeof(double2), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.vel, s->a.vel, n * sizeof(double2), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.vdot, s->a.vdot, n * sizeof(double2), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.vpred, s->a.vpred, n * sizeof(double2), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.e, s->a.e, n * sizeof(double), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.edot, s->a.edot, n * sizeof(double), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.epred, s->a.epred, n * sizeof(double), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.id, s->a.id, n * sizeof(int64_t), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.pc, s->a.pc, n * sizeof(double4), dd, s->st));
  CKG(cudaMemcpyAsync(t.a.ghost, s->a.ghost, n * sizeof(uint8_t), dd, s->st));
  CKG(cudaMemsetAsync(t.cellCount, 0, ncount * sizeof(uint32_t), s->st));
  CKG(cudaMemsetAsync(t.nn, 0xff, cap32 * SPHB_K * sizeof(uint32_t), s->st));
  CKG(cudaStreamSynchronize(s->st));
#undef CKG
  CapArrays old;
  old.a = s->a; old.b = s->b; old.spos = s->spos; old.hguess = s->hguess; old.keys = s->keys; old.keysSorted = s->keysSorted;
  old.rank = s->rank; old.perm = s->perm; old.cellStart = s->cellStart; old.cellCount = s->cellCount; old.tileSum = s->tileSum;
  old.nn = s->nn; old.failList = s->failList; old.nx = s->nx; old.tinfo = s->tinfo; old.ptab = s->ptab; old.dexcl = s->dexcl; old.ucum = s->ucum; old.ubox = s->ubox;
  s->a = t.a; s->b = t.b; s->spos = t.spos; s->hguess = t.hguess; s->keys = t.keys; s->keysSorted = t.keysSorted;
  s->rank = t.rank; s->perm = t.perm; s->cellStart = t.cellStart; s->cellCount = t.cellCount; s->tileSum = t.tileSum;
  s->nn = t.nn; s->failList = t.failList; s->nx = t.nx; s->tinfo = t.tinfo; s->ptab = t.ptab; s->dexcl = t.dexcl; s->ucum = t.ucum; s->ubox = t.ubox; s->ubox_cap = (int)(ncm / 8 + 1024);
  old.release();
  invalidate_reuse(s);
  cudaFree(s->ring.inv); s->ring.inv = nullptr; s->ring.inv_cap = 0;
  s->cap = ncap;
  s->ncell_max = (int)ncm;
  s->ntiles_cap = ntiles;
  s->gtune.ncell_max = s->ncell_max;
  s->keys_ready = false;  // the new cellCount is already zero
  s->grid_next_ready = false;
  s->have_list = false;
  s->hacc_valid = false;
  s->qmax_valid = false;
  s->stats_dirty = true;
  return SPHB_OK;
}

// ids a permutation of 0..n-1?  (perm / packCount are free outside an evaluation; keys / rank are not: after a fused
// step they hold the next step's cell keys)
int require_dense_ids(sphb_sim* s) {
  if (s->ids_dense < 0) {
    const int n = (int)s->n;
    int bad = 0;
    CK(s, cudaMemsetAsync(s->perm, 0, (size_t)n * sizeof(uint32_t), s->st));
    CK(s, cudaMemsetAsync(s->packCount, 0, sizeof(int), s->st));
    k_check_dense<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.id, n, s->perm, s->packCount);
    CKL(s);
    CK(s, cudaMemcpyAsync(&bad, s->packCount, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CK(s, cudaStreamSynchronize(s->st));
    s->ids_dense = bad ? 0 : 1;
  }
  if (!s->ids_dense) return fail(s, SPHB_E_STATE, "the particle ids are not a permutation of 0..N-1: by-id transfers need dense ids");
  return SPHB_OK;
}

// ring, inside a cycle: the owned particles are gathered through an index list (their order is arbitrary; join on id)
template <typename T>
int download_gathered(sphb_sim* s, const T* src, const uint32_t* list, int64_t n, void* host) {
  int rc = ensure_scratch(s, (size_t)n * sizeof(T)); if (rc) return rc;
  k_gather_list<T><<<cdiv(n, 256), 256, 0, s->st>>>(src, list, (int)n, (T*)s->scratch);
  CKL(s);
  CK(s, cudaMemcpyAsync(host, s->scratch, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  return SPHB_OK;
}

int download_interleaved(sphb_sim* s, uint32_t mask, void* const* hp) {
  const int64_t n = s->n;
  const int ntot = (int)(s->n + s->nghost);
  if (mask & (SPHB_MASK(SPHB_F_NN_IDX) | SPHB_MASK(SPHB_F_NN_DIST) | SPHB_MASK(SPHB_F_NN_POS)))
    return fail(s, SPHB_E_STATE, "ring: neighbour lists cannot be downloaded in the middle of a reuse cycle (their indices include ghosts)");
  uint32_t* list = s->perm;  // free outside a rebuild evaluation
  CK(s, cudaMemsetAsync(s->packCount, 0, sizeof(int), s->st));
  k_ring_owned_list<<<cdiv(ntot, 256), 256, 0, s->st>>>(s->a.ghost, ntot, list, s->packCount);
  CKL(s);
  int rc;
#define DG(field, src, T) \
  if (mask & SPHB_MASK(field)) { if ((rc = download_gathered<T>(s, (src), list, n, hp[field]))) return rc; }
  DG(SPHB_F_POS, s->a.pos, double2); DG(SPHB_F_VEL, s->a.vel, double2); DG(SPHB_F_E, s->a.e, double);
  DG(SPHB_F_EDOT, s->a.edot, double); DG(SPHB_F_VDOT, s->a.vdot, double2); DG(SPHB_F_EPRED, s->a.epred, double);
  DG(SPHB_F_VPRED, s->a.vpred, double2); DG(SPHB_F_ID, s->a.id, int64_t);
#undef DG
  if (mask & (SPHB_MASK(SPHB_F_RHO) | SPHB_MASK(SPHB_F_C) | SPHB_MASK(SPHB_F_H))) {
    rc = ensure_scratch(s, (size_t)n * 7 * sizeof(double)); if (rc) return rc;
    double4* pcg = (double4*)s->scratch;
    double* sc = (double*)(pcg + n);
    k_gather_list<double4><<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pc, list, (int)n, pcg);
    k_split_pc<<<cdiv(n, 256), 256, 0, s->st>>>(pcg, (int)n, sc, sc + n, sc + 2 * n);
    CKL(s);
    if (mask & SPHB_MASK(SPHB_F_RHO)) CK(s, cudaMemcpyAsync(hp[SPHB_F_RHO], sc, n * 8, cudaMemcpyDeviceToHost, s->st));
    if (mask & SPHB_MASK(SPHB_F_C)) CK(s, cudaMemcpyAsync(hp[SPHB_F_C], sc + n, n * 8, cudaMemcpyDeviceToHost, s->st));
    if (mask & SPHB_MASK(SPHB_F_H)) CK(s, cudaMemcpyAsync(hp[SPHB_F_H], sc + 2 * n, n * 8, cudaMemcpyDeviceToHost, s->st));
    CK(s, cudaStreamSynchronize(s->st));
  }
  return SPHB_OK;
}

void ring_comm_destroy(sphb_sim* s);

}  // namespace

// =================================================================================================
extern "C" {

int sphb_create(const sphb_params* p, int64_t n, int64_t capacity, const double* pos_xy, const double* vel_xy,
                const double* e, const double* rho, const int64_t* id, sphb_sim** out) {
  return create_common(p, n, capacity, pos_xy, vel_xy, e, rho, id, cudaMemcpyHostToDevice, out);
}

int sphb_create_device(const sphb_params* p, int64_t n, int64_t capacity, const double* d_pos_xy,
                       const double* d_vel_xy, const double* d_e, const int64_t* d_id, sphb_sim** out) {
  return create_common(p, n, capacity, d_pos_xy, d_vel_xy, d_e, nullptr, d_id, cudaMemcpyDeviceToDevice, out);
}

void sphb_destroy(sphb_sim* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->st) cudaStreamSynchronize(s->st);
  free_soa(s->a); free_soa(s->b);
  cudaFree(s->spos); cudaFree(s->hguess);
  cudaFree(s->keys); cudaFree(s->keysSorted); cudaFree(s->rank); cudaFree(s->perm);
  cudaFree(s->cellCount); cudaFree(s->tileSum); cudaFree(s->cellStart); cudaFree(s->nn); cudaFree(s->failList); cudaFree(s->failCount);
  cudaFree(s->packCount); cudaFree(s->hacc); cudaFree(s->qmax);
  cudaFree(s->nx); cudaFree(s->tinfo); cudaFree(s->ptab); cudaFree(s->dexcl); cudaFree(s->rs); cudaFree(s->stat_dev);
  cudaFree(s->ucum); cudaFree(s->ubox); cudaFree(s->ubox_ok);
  if (s->stat_host) cudaFreeHost(s->stat_host);
  if (s->stat_event_valid) cudaEventDestroy(s->stat_event);
  if (s->st_copy) { cudaStreamSynchronize(s->st_copy); cudaStreamDestroy(s->st_copy); }
  if (s->up_done) cudaEventDestroy(s->up_done);
  if (s->up_free) cudaEventDestroy(s->up_free);
  cudaFree(s->up_stage);
  for (int side = 0; side < 2; ++side) { cudaFree(s->ring.sbuf[side]); cudaFree(s->ring.rbuf[side]); cudaFree(s->ring.sidx[side]); cudaFree(s->ring.hsrc[side]); }
  cudaFree(s->ring.inv); cudaFree(s->ring.red_dev); cudaFree(s->ring.msg);
  if (s->ring.refused_host) cudaFreeHost(s->ring.refused_host);
  if (s->ring.comm) ring_comm_destroy(s);
  cudaFree(s->dflags); cudaFree(s->statPart); cudaFree(s->stats); cudaFree(s->grid); cudaFree(s->grid_next); cudaFree(s->scratch);
  for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

const char* sphb_last_error(const sphb_sim* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int sphb_set_params(sphb_sim* s, const sphb_params* p) {
  int rc = enter(s); if (rc) return rc;
  rc = check_params(s, p); if (rc) return rc;
  if (p->device != s->device) return fail(s, SPHB_E_INVALID, "device cannot change after create");
  if (p->precision != s->prm.precision) return fail(s, SPHB_E_INVALID, "precision cannot change after create");
  if (!same_params(s->prm, *p)) { invalidate_reuse(s); s->touched = true; }
  s->prm = *p;
  return SPHB_OK;
}
int sphb_get_params(const sphb_sim* s, sphb_params* p) {
  if (!s || !p) return SPHB_E_INVALID;
  *p = s->prm;
  return SPHB_OK;
}
int64_t sphb_count(const sphb_sim* s) { return s ? s->n : -1; }
int64_t sphb_current_step(const sphb_sim* s) { return s ? s->cur_step : -1; }
int sphb_set_current_step(sphb_sim* s, int64_t step) {
  if (!s) return SPHB_E_INVALID;
  if (step < 0) return fail(s, SPHB_E_INVALID, "negative step");
  s->cur_step = step;
  return SPHB_OK;
}

int sphb_append(sphb_sim* s, int64_t n, const double* pos_xy, const double* vel_xy, const double* e, const double* rho,
                const int64_t* id) {
  int rc = enter(s); if (rc) return rc;
  if (n < 0 || (n > 0 && !pos_xy)) return fail(s, SPHB_E_INVALID, "bad particle arrays");
  if (s->interleaved) drop_ghosts(s);  // ring, inside a cycle (state-changing calls are collective: every rank ends it)
  if (s->nghost) return fail(s, SPHB_E_STATE, "append while ghosts are attached");
  if (n == 0) return SPHB_OK;
  if (s->n + n > s->cap) {  // Go's append reallocates (sph.go:79)
    rc = grow_capacity(s, s->n + n);
    if (rc) return rc;
  }
  rc = upload_common(s, s->n, n, pos_xy, vel_xy, e, rho, id, cudaMemcpyHostToDevice, s->n);
  if (rc) return rc;
  s->n += n;
  s->ids_dense = -1;
  return SPHB_OK;
}

int sphb_step(sphb_sim* s, int32_t nsteps) {
  int rc = enter(s); if (rc) return rc;
  if (s->slab_on) return fail(s, SPHB_E_STATE, "slab mode: use sphb_slab_step_begin / _end");
  for (int32_t k = 0; k < nsteps; ++k) {
    if (s->cur_step == 0) {  // sph.go:89-103: VPred = Vel, EPred = E, forces once
      rc = forces(s, MODE_INIT, false); if (rc) return rc;
    }
    rc = forces(s, MODE_DRIFT, true); if (rc) return rc;
    s->cur_step += 1;
    s->counters[SPHB_CNT_STEPS] += 1;
  }
  return SPHB_OK;
}

int sphb_calc_forces(sphb_sim* s) {
  int rc = enter(s); if (rc) return rc;
  if (s->interleaved) drop_ghosts(s);
  if (s->slab_on) return fail(s, SPHB_E_STATE, "slab mode: use sphb_slab_step_begin / _end");
  return forces(s, MODE_ASIS, false);
}

int sphb_knn(sphb_sim* s, const double hor[2], const double ver[2]) {
  int rc = enter(s); if (rc) return rc;
  if (!hor || !ver) return fail(s, SPHB_E_INVALID, "hor/ver is NULL");
  if (hor[0] == SPHB_OPEN_LO && hor[1] != SPHB_OPEN_HI)
    return fail(s, SPHB_E_INVALID, "cannot have open and periodic boundary in horizontal at same time!");
  if (ver[0] == SPHB_OPEN_LO && ver[1] != SPHB_OPEN_HI)
    return fail(s, SPHB_E_INVALID, "cannot have open and periodic boundary in vertical at same time!");
  if ((hor[0] != SPHB_OPEN_LO && !(hor[1] > hor[0])) || (ver[0] != SPHB_OPEN_LO && !(ver[1] > ver[0])))
    return fail(s, SPHB_E_INVALID, "empty period");
  rc = build_neighbours(s, MODE_ASIS, hor, ver, s->prm.kernel, false);
  if (rc) return rc;
  s->stats_dirty = true;
  return check_async(s);
}

int sphb_density(sphb_sim* s, int32_t kernel) {
  int rc = enter(s); if (rc) return rc;
  if (kernel < 0 || kernel > 2) return fail(s, SPHB_E_INVALID, "unknown kernel %d", kernel);
  if (!s->have_list) return fail(s, SPHB_E_STATE, "density before knn: NNDists[0] is 0 (sph.go:307)");
  const int ntot = (int)(s->n + s->nghost);
  const PhysP ph = make_phys(s->prm, kernel);
  switch (kernel) {
    case 0: k_density_from_list<0><<<cdiv(ntot, 128), 128, 0, s->st>>>(s->spos, s->nn, ntot, s->grid, ph, s->a.pc); break;
    case 1: k_density_from_list<1><<<cdiv(ntot, 128), 128, 0, s->st>>>(s->spos, s->nn, ntot, s->grid, ph, s->a.pc); break;
    default: k_density_from_list<2><<<cdiv(ntot, 128), 128, 0, s->st>>>(s->spos, s->nn, ntot, s->grid, ph, s->a.pc); break;
  }
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  CKL(s);
  s->stats_dirty = true;
  return SPHB_OK;
}

void* sphb_stream(sphb_sim* s) { return s ? (void*)s->st : nullptr; }

int sphb_sync(sphb_sim* s) {
  int rc = enter(s); if (rc) return rc;
  return check_async(s);
}

int sphb_download(sphb_sim* s, uint32_t mask, void* const* hp, int64_t capacity, int64_t* n_out) {
  int rc = enter(s); if (rc) return rc;
  rc = check_async(s); if (rc) return rc;
  const int64_t n = s->n;
  if (n_out) *n_out = n;
  if (mask == 0) return SPHB_OK;
  if (!hp) return fail(s, SPHB_E_INVALID, "host_ptrs is NULL");
  if (capacity < n) return fail(s, SPHB_E_NOMEM, "download: capacity %lld < %lld particles", (long long)capacity, (long long)n);
  if (mask >> SPHB_F_COUNT) return fail(s, SPHB_E_INVALID, "unknown field bits in mask 0x%x", mask);
  for (int f = 0; f < SPHB_F_COUNT; ++f)
    if ((mask & SPHB_MASK(f)) && !hp[f]) return fail(s, SPHB_E_INVALID, "host_ptrs[%d] is NULL", f);
  if (n == 0) return SPHB_OK;
  if (s->interleaved) return download_interleaved(s, mask, hp);
#define D2H(field, src, bytes) \
  if (mask & SPHB_MASK(field)) CK(s, cudaMemcpyAsync(hp[field], (src), (size_t)(bytes), cudaMemcpyDeviceToHost, s->st))
  D2H(SPHB_F_POS, s->a.pos, n * 16);
  D2H(SPHB_F_VEL, s->a.vel, n * 16);
  D2H(SPHB_F_E, s->a.e, n * 8);
  D2H(SPHB_F_EDOT, s->a.edot, n * 8);
  D2H(SPHB_F_VDOT, s->a.vdot, n * 16);
  D2H(SPHB_F_EPRED, s->a.epred, n * 8);
  D2H(SPHB_F_VPRED, s->a.vpred, n * 16);
  D2H(SPHB_F_ID, s->a.id, n * 8);
  if (mask & (SPHB_MASK(SPHB_F_RHO) | SPHB_MASK(SPHB_F_C) | SPHB_MASK(SPHB_F_H))) {
    rc = ensure_scratch(s, (size_t)n * 3 * sizeof(double)); if (rc) return rc;
    double* sc = (double*)s->scratch;
    k_split_pc<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pc, (int)n, sc, sc + n, sc + 2 * n);
    CKL(s);
    D2H(SPHB_F_RHO, sc, n * 8);
    D2H(SPHB_F_C, sc + n, n * 8);
    D2H(SPHB_F_H, sc + 2 * n, n * 8);
    CK(s, cudaStreamSynchronize(s->st));
  }
  if (mask & (SPHB_MASK(SPHB_F_NN_IDX) | SPHB_MASK(SPHB_F_NN_DIST) | SPHB_MASK(SPHB_F_NN_POS))) {
    if (!s->have_list) return fail(s, SPHB_E_STATE, "no neighbour list for the current particle order: call knn / calc_forces first");
    // chunked so that the staging buffer stays small next to the state
    const int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 20);
    const size_t per = SPHB_K * (sizeof(int32_t) + sizeof(double) + sizeof(double2));
    rc = ensure_scratch(s, (size_t)chunk * per); if (rc) return rc;
    for (int64_t off = 0; off < n; off += chunk) {
      const int64_t m = std::min(chunk, n - off);
      // kernel indexes particles globally; run it on [off, off+m) by offsetting outputs
      double* dist = (double*)s->scratch;
      double2* npos = (double2*)(dist + (size_t)chunk * SPHB_K);
      int32_t* idx = (int32_t*)(npos + (size_t)chunk * SPHB_K);
      k_expand_list<<<cdiv(m, 128), 128, 0, s->st>>>(s->spos, s->a.pos, s->nn, (int)off, (int)m, s->grid,
                                                    (mask & SPHB_MASK(SPHB_F_NN_IDX)) ? idx : nullptr,
                                                    (mask & SPHB_MASK(SPHB_F_NN_DIST)) ? dist : nullptr,
                                                    (mask & SPHB_MASK(SPHB_F_NN_POS)) ? npos : nullptr);
      CKL(s);
      if (mask & SPHB_MASK(SPHB_F_NN_IDX))
        CK(s, cudaMemcpyAsync((int32_t*)hp[SPHB_F_NN_IDX] + off * SPHB_K, idx, (size_t)m * SPHB_K * 4, cudaMemcpyDeviceToHost, s->st));
      if (mask & SPHB_MASK(SPHB_F_NN_DIST))
        CK(s, cudaMemcpyAsync((double*)hp[SPHB_F_NN_DIST] + off * SPHB_K, dist, (size_t)m * SPHB_K * 8, cudaMemcpyDeviceToHost, s->st));
      if (mask & SPHB_MASK(SPHB_F_NN_POS))
        CK(s, cudaMemcpyAsync((double*)hp[SPHB_F_NN_POS] + off * SPHB_K * 2, npos, (size_t)m * SPHB_K * 16, cudaMemcpyDeviceToHost, s->st));
      CK(s, cudaStreamSynchronize(s->st));
    }
  }
#undef D2H
  CK(s, cudaStreamSynchronize(s->st));
  return SPHB_OK;
}

int sphb_upload(sphb_sim* s, uint32_t mask, const void* const* hp, int64_t n) {
  int rc = enter(s); if (rc) return rc;
  if (s->interleaved) drop_ghosts(s);
  if (n != s->n) return fail(s, SPHB_E_INVALID, "upload: n = %lld but the simulation holds %lld particles", (long long)n, (long long)s->n);
  const uint32_t allowed = SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL) | SPHB_MASK(SPHB_F_E) | SPHB_MASK(SPHB_F_RHO) |
                           SPHB_MASK(SPHB_F_VDOT) | SPHB_MASK(SPHB_F_EDOT) | SPHB_MASK(SPHB_F_EPRED) | SPHB_MASK(SPHB_F_VPRED);
  if (mask & ~allowed) return fail(s, SPHB_E_INVALID, "upload: field mask 0x%x has non-writable fields", mask);
  if (mask && !hp) return fail(s, SPHB_E_INVALID, "host_ptrs is NULL");
  for (int f = 0; f < SPHB_F_COUNT; ++f)
    if ((mask & SPHB_MASK(f)) && !hp[f]) return fail(s, SPHB_E_INVALID, "host_ptrs[%d] is NULL", f);
  if (n == 0) return SPHB_OK;
#define H2D(field, dst, bytes) \
  if (mask & SPHB_MASK(field)) CK(s, cudaMemcpyAsync((dst), hp[field], (size_t)(bytes), cudaMemcpyHostToDevice, s->st))
  H2D(SPHB_F_POS, s->a.pos, n * 16);
  H2D(SPHB_F_VEL, s->a.vel, n * 16);
  H2D(SPHB_F_E, s->a.e, n * 8);
  H2D(SPHB_F_EDOT, s->a.edot, n * 8);
  H2D(SPHB_F_VDOT, s->a.vdot, n * 16);
  H2D(SPHB_F_EPRED, s->a.epred, n * 8);
  H2D(SPHB_F_VPRED, s->a.vpred, n * 16);
#undef H2D
  if (mask & SPHB_MASK(SPHB_F_RHO)) {
    rc = ensure_scratch(s, (size_t)n * sizeof(double)); if (rc) return rc;
    CK(s, cudaMemcpyAsync(s->scratch, hp[SPHB_F_RHO], (size_t)n * 8, cudaMemcpyHostToDevice, s->st));
    k_set_pc<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pc, (int)n, (const double*)s->scratch, 0);
    CKL(s);
  }
  if (mask & SPHB_MASK(SPHB_F_POS)) s->have_list = false;  // the list describes the old positions; h stays a valid first guess
  if (mask & (SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL))) drop_ready_keys(s);  // they were keys of the old state
  if (mask) { invalidate_reuse(s); s->touched = true; }
  s->stats_dirty = true;
  CK(s, cudaStreamSynchronize(s->st));
  return SPHB_OK;
}

int sphb_upload_by_id(sphb_sim* s, uint32_t mask, const void* const* hp, int64_t n) {
  int rc = enter(s); if (rc) return rc;
  if (s->interleaved) drop_ghosts(s);
  if (n != s->n) return fail(s, SPHB_E_INVALID, "upload: n = %lld but the simulation holds %lld particles", (long long)n, (long long)s->n);
  const uint32_t allowed = SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL) | SPHB_MASK(SPHB_F_E);
  if (mask & ~allowed) return fail(s, SPHB_E_INVALID, "upload_by_id: POS, VEL, E only (mask 0x%x)", mask);
  if (mask && !hp) return fail(s, SPHB_E_INVALID, "host_ptrs is NULL");
  for (int f = 0; f < SPHB_F_COUNT; ++f)
    if ((mask & SPHB_MASK(f)) && !hp[f]) return fail(s, SPHB_E_INVALID, "host_ptrs[%d] is NULL", f);
  if (n == 0 || mask == 0) return SPHB_OK;
  if (s->slab_on) return fail(s, SPHB_E_STATE, "upload_by_id: not available in slab mode (a handle owns a subset of the ids)");
  rc = require_dense_ids(s); if (rc) return rc;
  rc = ensure_scratch(s, (size_t)n * 40); if (rc) return rc;
  double2* sp = (double2*)s->scratch;
  double2* sv = sp + n;
  double* se = (double*)(sv + n);
  if (mask & SPHB_MASK(SPHB_F_POS)) CK(s, cudaMemcpyAsync(sp, hp[SPHB_F_POS], (size_t)n * 16, cudaMemcpyHostToDevice, s->st));
  if (mask & SPHB_MASK(SPHB_F_VEL)) CK(s, cudaMemcpyAsync(sv, hp[SPHB_F_VEL], (size_t)n * 16, cudaMemcpyHostToDevice, s->st));
  if (mask & SPHB_MASK(SPHB_F_E)) CK(s, cudaMemcpyAsync(se, hp[SPHB_F_E], (size_t)n * 8, cudaMemcpyHostToDevice, s->st));
  k_gather_by_id<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.id, (int)n, (mask & SPHB_MASK(SPHB_F_POS)) ? sp : nullptr,
                                                 (mask & SPHB_MASK(SPHB_F_VEL)) ? sv : nullptr,
                                                 (mask & SPHB_MASK(SPHB_F_E)) ? se : nullptr, s->a.pos, s->a.vel, s->a.e);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  CKL(s);
  if (mask & SPHB_MASK(SPHB_F_POS)) s->have_list = false;
  if (mask & (SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL))) drop_ready_keys(s);
  invalidate_reuse(s);
  s->touched = true;
  s->stats_dirty = true;
  CK(s, cudaStreamSynchronize(s->st));  // host buffers are borrowed for the duration of the call only
  return SPHB_OK;
}

int sphb_upload_by_id_begin(sphb_sim* s, uint32_t mask, const void* const* hp, int64_t n) {
  int rc = enter(s); if (rc) return rc;
  if (s->up_mask) return fail(s, SPHB_E_STATE, "upload_by_id_begin: an upload is already in flight (call sphb_upload_by_id_end)");
  if (n != s->n) return fail(s, SPHB_E_INVALID, "upload: n = %lld but the simulation holds %lld particles", (long long)n, (long long)s->n);
  const uint32_t allowed = SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL) | SPHB_MASK(SPHB_F_E);
  if (mask & ~allowed) return fail(s, SPHB_E_INVALID, "upload_by_id: POS, VEL, E only (mask 0x%x)", mask);
  if (mask && !hp) return fail(s, SPHB_E_INVALID, "host_ptrs is NULL");
  for (int f = 0; f < SPHB_F_COUNT; ++f)
    if ((mask & SPHB_MASK(f)) && !hp[f]) return fail(s, SPHB_E_INVALID, "host_ptrs[%d] is NULL", f);
  if (n == 0 || mask == 0) return SPHB_OK;
  if (s->slab_on) return fail(s, SPHB_E_STATE, "upload_by_id: not available in slab mode (a handle owns a subset of the ids)");
  if (!s->st_copy) {
    CK(s, cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking));
    CK(s, cudaEventCreateWithFlags(&s->up_done, cudaEventDisableTiming));
    CK(s, cudaEventCreateWithFlags(&s->up_free, cudaEventDisableTiming));
  }
  if ((size_t)n * 40 > s->up_stage_bytes) {
    CK(s, cudaStreamSynchronize(s->st));
    CK(s, cudaStreamSynchronize(s->st_copy));
    cudaFree(s->up_stage); s->up_stage = nullptr; s->up_stage_bytes = 0;
    CK(s, cudaMalloc(&s->up_stage, (size_t)n * 40));
    s->up_stage_bytes = (size_t)n * 40;
  } else {
    CK(s, cudaStreamWaitEvent(s->st_copy, s->up_free, 0));  // the previous upload's scatter has read the staging area
  }
  double2* sp = (double2*)s->up_stage;
  double2* sv = sp + n;
  double* se = (double*)(sv + n);
  if (mask & SPHB_MASK(SPHB_F_POS)) CK(s, cudaMemcpyAsync(sp, hp[SPHB_F_POS], (size_t)n * 16, cudaMemcpyHostToDevice, s->st_copy));
  if (mask & SPHB_MASK(SPHB_F_VEL)) CK(s, cudaMemcpyAsync(sv, hp[SPHB_F_VEL], (size_t)n * 16, cudaMemcpyHostToDevice, s->st_copy));
  if (mask & SPHB_MASK(SPHB_F_E)) CK(s, cudaMemcpyAsync(se, hp[SPHB_F_E], (size_t)n * 8, cudaMemcpyHostToDevice, s->st_copy));
  CK(s, cudaEventRecord(s->up_done, s->st_copy));
  s->up_mask = mask; s->up_n = n;
  return SPHB_OK;
}

int sphb_upload_by_id_end(sphb_sim* s) {
  int rc = enter(s); if (rc) return rc;
  if (!s->up_mask) return SPHB_OK;
  const uint32_t mask = s->up_mask;
  const int64_t n = s->up_n;
  s->up_mask = 0;
  CK(s, cudaEventSynchronize(s->up_done));  // the host buffers are the caller's again
  if (n != s->n) return fail(s, SPHB_E_STATE, "upload_by_id_end: the particle count changed while the upload was in flight");
  if (s->interleaved) drop_ghosts(s);
  rc = require_dense_ids(s); if (rc) return rc;
  double2* sp = (double2*)s->up_stage;
  double2* sv = sp + n;
  double* se = (double*)(sv + n);
  k_gather_by_id<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.id, (int)n, (mask & SPHB_MASK(SPHB_F_POS)) ? sp : nullptr,
                                                 (mask & SPHB_MASK(SPHB_F_VEL)) ? sv : nullptr,
                                                 (mask & SPHB_MASK(SPHB_F_E)) ? se : nullptr, s->a.pos, s->a.vel, s->a.e);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  CKL(s);
  CK(s, cudaEventRecord(s->up_free, s->st));
  if (mask & SPHB_MASK(SPHB_F_POS)) s->have_list = false;
  if (mask & (SPHB_MASK(SPHB_F_POS) | SPHB_MASK(SPHB_F_VEL))) drop_ready_keys(s);
  invalidate_reuse(s);
  s->touched = true;
  s->stats_dirty = true;
  return SPHB_OK;
}

int sphb_reduce(sphb_sim* s, int32_t which, double* out) {
  int rc = enter(s); if (rc) return rc;
  if (!out) return fail(s, SPHB_E_INVALID, "out is NULL");
  rc = check_async(s); if (rc) return rc;
  if (which == SPHB_LAST_VEL_NORM) {  // TotalMomentum assigns instead of accumulating (sph.go:457-463)
    if (s->n == 0) { *out = 0.0; return SPHB_OK; }
    double v[2];
    CK(s, cudaMemcpyAsync(v, s->a.vel + (s->n - 1), sizeof v, cudaMemcpyDeviceToHost, s->st));
    CK(s, cudaStreamSynchronize(s->st));
    *out = std::sqrt(v[0] * v[0] + v[1] * v[1]);
    return SPHB_OK;
  }
  if (which != SPHB_SUM_E && which != SPHB_SUM_RHO) return fail(s, SPHB_E_INVALID, "unknown reduction %d", which);
  if (s->n == 0) { *out = 0.0; return SPHB_OK; }
  if (s->stats_dirty) { rc = refresh_stats(s); if (rc) return rc; }
  double st[STAT_N];
  CK(s, cudaMemcpyAsync(st, s->stats, sizeof st, cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  *out = which == SPHB_SUM_E ? st[6] : st[7];
  return SPHB_OK;
}

int sphb_frame(sphb_sim* s, int32_t width, int32_t height, float* xy_out, uint8_t* colour_out, int64_t* id_out,
               int64_t capacity, int64_t* n_out) {
  int rc = enter(s); if (rc) return rc;
  rc = check_async(s); if (rc) return rc;
  if (s->interleaved) return fail(s, SPHB_E_STATE, "ring: sphb_frame in the middle of a reuse cycle is not supported (download the fields instead)");
  const int64_t n = s->n;
  if (n_out) *n_out = n;
  if (!xy_out || !colour_out) return fail(s, SPHB_E_INVALID, "xy_out / colour_out is NULL");
  if (capacity < n) return fail(s, SPHB_E_NOMEM, "frame: capacity %lld < %lld particles", (long long)capacity, (long long)n);
  if (n == 0) return SPHB_OK;
  rc = ensure_scratch(s, (size_t)n * (sizeof(float2) + 1) + 16); if (rc) return rc;
  float2* xy = (float2*)s->scratch;
  uint8_t* col = (uint8_t*)(xy + n);
  // colorFormula = Rho / (m * float64(len(Particles) * 10)) * 256: the divisor is one rounded product
  const double div = s->prm.particle_mass * (double)(n * 10);
  if (!id_out) { rc = require_dense_ids(s); if (rc) return rc; }
  k_frame<<<cdiv(n, 256), 256, 0, s->st>>>(s->a.pos, s->a.pc, id_out ? nullptr : s->a.id, (int)n, (float)width, (float)height, div, xy, col);
  s->counters[SPHB_CNT_KERNEL_LAUNCHES] += 1;
  CKL(s);
  CK(s, cudaMemcpyAsync(xy_out, xy, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaMemcpyAsync(colour_out, col, (size_t)n, cudaMemcpyDeviceToHost, s->st));
  if (id_out) CK(s, cudaMemcpyAsync(id_out, s->a.id, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  return SPHB_OK;
}

int sphb_max_h(sphb_sim* s, double* out) {
  int rc = enter(s); if (rc) return rc;
  if (!out) return fail(s, SPHB_E_INVALID, "out is NULL");
  rc = check_async(s); if (rc) return rc;  // the slab driver calls this once per evaluation: surfaces GHOST_THIN etc.
  if (s->qmax_valid) {
    float q[2];
    CK(s, cudaMemcpyAsync(q, s->qmax, sizeof q, cudaMemcpyDeviceToHost, s->st));
    CK(s, cudaStreamSynchronize(s->st));
    *out = (double)q[0];
    return SPHB_OK;
  }
  if (s->stats_dirty) { rc = refresh_stats(s); if (rc) return rc; }
  double st[STAT_N];
  CK(s, cudaMemcpyAsync(st, s->stats, sizeof st, cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  *out = st[8] > 0.0 ? st[5] : 0.0;
  return SPHB_OK;
}

int sphb_max_speed(sphb_sim* s, double* out) {
  int rc = enter(s); if (rc) return rc;
  if (!out) return fail(s, SPHB_E_INVALID, "out is NULL");
  if (s->n == 0) { *out = 0.0; return SPHB_OK; }
  if (s->qmax_valid) {
    float q[2];
    CK(s, cudaMemcpyAsync(q, s->qmax, sizeof q, cudaMemcpyDeviceToHost, s->st));
    CK(s, cudaStreamSynchronize(s->st));
    *out = std::sqrt((double)q[1]) * (1.0 + 1e-7);
    return SPHB_OK;
  }
  if (s->stats_dirty) { rc = refresh_stats(s); if (rc) return rc; }
  double st[STAT_N];
  CK(s, cudaMemcpyAsync(st, s->stats, sizeof st, cudaMemcpyDeviceToHost, s->st));
  CK(s, cudaStreamSynchronize(s->st));
  *out = st[9] > 0.0 ? std::sqrt(st[9]) : 0.0;
  return SPHB_OK;
}

int sphb_phase_times(sphb_sim* s, double* ms, int32_t n) {
  int rc = enter(s); if (rc) return rc;
  if (!ms || n < 0 || n > SPHB_PH_COUNT) return fail(s, SPHB_E_INVALID, "bad arguments");
  if (!s->ev_valid) return fail(s, SPHB_E_STATE, "no step has run yet");
  CK(s, cudaStreamSynchronize(s->st));
  double t[SPHB_PH_COUNT];
  for (int k = 0; k < SPHB_PH_TOTAL; ++k) {
    float f = 0;
    CK(s, cudaEventElapsedTime(&f, s->ev[k], s->ev[k + 1]));
    t[k] = f;
  }
  float f = 0;
  CK(s, cudaEventElapsedTime(&f, s->ev[0], s->ev[SPHB_PH_TOTAL]));
  t[SPHB_PH_TOTAL] = f;
  for (int k = 0; k < n; ++k) ms[k] = t[k];
  return SPHB_OK;
}

int sphb_counters(const sphb_sim* s, int64_t* out, int32_t n) {
  if (!s || !out || n < 0 || n > SPHB_CNT_COUNT) return SPHB_E_INVALID;
  for (int k = 0; k < n; ++k) out[k] = s->counters[k];
  return SPHB_OK;
}

// ---- slab decomposition: implemented in sphb_slab.inc (same translation unit) ----
#include "sphb_slab.inc"
// ---- the slab ring inside the library (NCCL or slabs in one process) ----
#include "sphb_ring.inc"

}  // extern "C"
