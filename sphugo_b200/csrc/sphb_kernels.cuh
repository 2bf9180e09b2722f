// sphb_kernels.cuh — device code of libsphb.so (sm_100a).
//
// The hot path of bbeni/sphugo's (*Simulation).Step() (reference sim/sph.go:64-198), re-designed for a
// B200: no tree, no per-particle priority queue in memory.  Pipeline per force evaluation
//   K1 keys      drift-1 + uniform-cell key                 (replaces Partition      core.go:126-164)
//   SORT         counting sort by cell (scan of the counts)  (replaces Treebuild      core.go:172-224)
//   K2 reorder   SoA gather + predict + cell table          (replaces BoundingSpheres core.go:229-312)
//   K3 knn       exact kNN(32) + density + sound speed      (nearest-neighbour.go:28-165, sph.go:306-323,423-429)
//   K4 force     pressure + viscosity, kick, drift-2, walls (sph.go:327-401, 122-193)
// Nothing here is a dense contraction, so there are no tensor-core instructions: the kernels are
// latency/issue and bandwidth bound and are laid out for coalesced SoA access, warp-uniform
// (broadcast) candidate loads and shared-memory staging of the per-particle candidate columns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#define SPHB_K 32
#define IMG_SHIFT 28
#define IDX_MASK 0x0FFFFFFFu

struct GridP {
  double ox, oy;        // origin of cell (0,0)
  double dx, dy;        // cell edge
  double inv_dx, inv_dy;
  double Lx, Ly;        // period of a wrapping axis (hor[1]-hor[0]), 0 if the axis is open
  double lox, loy;      // low end of the periodic interval
  int ncx, ncy;
  int wrapx, wrapy;
  int framex;  // slab mode on a periodic x axis: positions are taken in the image frame [lox, lox + Lx) centred on
               // the slab, but the grid itself does not wrap (the neighbour slabs supply ghosts instead)
  int sides;   // slab mode: bit 0 / bit 1 = the low-x / high-x grid edge is a ghost-layer boundary, not an open end
};

struct PhysP {
  double dtH, gamma, mass, gx, gy;
  double hor0, hor1, ver0, ver1;      // Step()'s wrap interval (sph.go:147-167); +-DBL_MAX when open
  double rL, rR, rU, rD;              // reflections (sph.go:170-193)
  double Fpref, DFpref, cfac;         // kernel prefactors; cfac = gamma*(gamma-1) (sph.go:426)
  int kernel;
};

// What a rebuild evaluation leaves for the reuse evaluations of its cycle, per tile (32 consecutive particles of the
// sorted order): the staged union block of the tile search - its pieces (contiguous index ranges, one per grid row and
// periodic image) and extent - taken wide enough for the skin.  The candidates of a particle are then 16-bit slots of
// that block for the annulus pass that follows the tile search (sphb_reuse.cuh).
struct TileInfo { int npc, nst, c0, c1, r0, r1; };  // npc = 0: the tile has no shared block (no reuse for its particles)
struct KnnExt {           // (nx == nullptr: off)
  uint32_t* nx;          // further candidates [tile][SPHB_KX][lane], list entries like nn; 0xffffffff = none
  TileInfo* tinfo;       // [tile]
  uint2* ptab;           // [tile][32] = {first index | image code << 28, length} of piece p
  double* dexcl;
  double skin;           // candidates are collected up to h_prev * (1 + skin)
};

struct KnnTune {
  double guess_margin;   // search radius = h_prev * (1 + margin)
  double k_target;       // expected candidates inside the first-guess radius when no h_prev exists
  int cap;               // candidate column capacity per particle (shared memory)
  int cap0;              // same for the first evaluation (no previous h)
  int ncw;               // staged candidates per tile (shared memory), multiple of 8
  int ncw0;              // same for the first evaluation
};

// -------------------------------------------------------------------------------------------------
// Certified reuse of the neighbour lists (DESIGN "list reuse").  A REBUILD evaluation (sort + tile search) keeps, next to
// the 32 neighbours, up to SPHB_KX further candidates per particle (`nx`) and an exclusion radius `dexcl`: every particle
// that is neither in nn nor in nx was at least dexcl away when the lists were built.  A REUSE evaluation moves nothing in
// memory (no sort, no reorder): it re-evaluates those <= 48 candidates at their new positions, h' = the 32nd smallest
// distance, and the result is the exact kNN whenever  h' + D < dexcl, where D bounds |u_i - u_j| over all pairs (u =
// displacement since the build): a particle outside the candidate set can have come closer by at most D.  Particles that
// fail the test go to the ring-expansion search (on the stale cells, widened by D).  Accepted results are exact kNN, so
// parity with the reference (nearest-neighbour.go:28-165) is untouched.
// -------------------------------------------------------------------------------------------------
#define SPHB_KX 16
#define RS_SLOTS 64
struct ReuseState {
  double D;                       // bound of the relative displacement of any two particles since the build
  double ubx, uby;                // mean displacement since the build: stale-cell lookups are shifted back by it
  double mrx, mry;                // reference displacement of the coming step (the mean of the previous one); any value is valid
  long long sum[RS_SLOTS][2];     // fixed-point sums of this step's displacements (integer: order independent)
  unsigned int Mbits;             // max |delta - mref| of this step as float bits, rounded up
  unsigned int age;               // reuse evaluations since the build
  unsigned int seq;               // evaluations recorded so far (host feedback)
  unsigned int pad;
};
struct ReuseStat {                // one record per evaluation, copied to pinned host memory (non-blocking feedback)
  unsigned int seq, age, refused, n;
  float D, dy;                    // D after this evaluation's displacement; row height of the grid in use (~ mean h)
  unsigned int rebuild, seq2;     // seq2 == seq marks a complete record
};

// device-side status word
#define DFLAG_UNDERFULL 1u
#define DFLAG_GHOST_THIN 2u  // slab mode: some h reached past the ghost layer
#define DFLAG_BUF_FULL 4u    // slab mode: a pack buffer or the particle capacity was exceeded

// per-particle flag (Soa.ghost)
#define GF_OWNED 0
#define GF_INNER 1    // ghost, evaluated (kNN + density) so that owned particles can read its rho, c, h
#define GF_OUTER 2    // ghost, candidate only
#define GF_LEAVING 3  // owned, packed for migration, removed by the next compaction

// -------------------------------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double wrap_coord(double x, double lo, double L) {
  // canonical image in [lo, lo+L); exact identity for lo <= x < lo+L.  Step() keeps positions within one period of
  // the box (sph.go:147-167), so the first three cases - no division - are the ones that run.
  const double hi = lo + L;
  if (x >= lo && x < hi) return x;
  if (x < lo && x >= lo - L) { const double xs = __dadd_rn(x, L); return xs >= hi ? __dsub_rn(xs, L) : xs; }
  if (x >= hi && x < hi + L) { const double xs = __dsub_rn(x, L); return xs < lo ? __dadd_rn(xs, L) : xs; }
  double w = floor((x - lo) / L);
  double xs = __dsub_rn(x, __dmul_rn(w, L));
  if (xs < lo) xs = __dadd_rn(xs, L);
  else if (xs >= hi) xs = __dsub_rn(xs, L);
  return xs;
}

__device__ __forceinline__ int cell_of(double xs, double o, double inv, int nc) {
  int c = (int)floor((xs - o) * inv);
  return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

// floor(u / nc) for u in [-nc, 2 nc): the periodic image (-1, 0, +1) an unwrapped cell index lies in.
// Every caller bounds its stencil by one period (wider stencils are refused or clamped first).
__device__ __forceinline__ int img_idx(int u, int nc) { return (u >= nc ? 1 : 0) - (u < 0 ? 1 : 0); }

// d^2 exactly as linear-algebra.go:61-64 evaluates it on amd64: two products, one sum, no FMA
__device__ __forceinline__ double dist_sq(double dx, double dy) {
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// 1/sqrt(x) and 1/x for normal x > 0: hardware seed (MUFU.RSQ64H / RCP64H, ~2^-22) + one third-order step, no
// special cases and no branches; relative error ~e^3 ~ 2^-64 before the final rounding (consumers carry a 1e-12 bar)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);  // 1 - x y^2
  const double t = fma(0.375, e, 0.5);
  return fma(y * e, t, y);                 // y (1 + e/2 + 3 e^2/8)
}
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);          // r (1 + e + e^2)
}
// sqrt(x) from fast_rsqrt with one residual correction: faithfully rounded (almost always correctly rounded)
__device__ __forceinline__ double fast_sqrt(double x, double y /* = fast_rsqrt(x) */) {
  const double s = x * y;
  return fma(0.5 * y, fma(-s, s, x), s);
}

// SPH kernels, sph.go:244-304.  q = r / h with h the distance to the 32nd neighbour.
template <int KERNEL>
__device__ __forceinline__ double kern_F(double q) {
  if (KERNEL == 0) return 1.0;
  if (KERNEL == 1) {  // both pieces evaluated, selected without a branch
    const double lo = q * q * q - q * q + (1.0 / 6.0);
    const double t = 1.0 - q;
    const double hi = t * t * t * (1.0 / 3.0);  // the reference divides by 3: <= 1 ulp apart
    return q < 0.5 ? lo : hi;
  }
  double t = 1.0 - q;
  double t2 = t * t;
  return t2 * t2 * (1.0 + 4.0 * q);
}
template <int KERNEL>
__device__ __forceinline__ double kern_DF(double q) {
  if (KERNEL == 1) {
    const double lo = 3.0 * q * q - 2.0 * q;
    const double t = 1.0 - q;
    return q < 0.5 ? lo : -t * t;
  }
  double t = 1.0 - q;
  return -10.0 * q * t * t * t;
}

// -------------------------------------------------------------------------------------------------
// K1: (optional drift-1) + cell key + per-cell count.  Reads pos, vel; writes key, rank-in-cell.
// The rank returned by the atomic is arbitrary among the particles of one cell; K2 re-ranks each cell by
// the previous index so that the final order (and every later summation order) is deterministic.
// -------------------------------------------------------------------------------------------------
template <bool DRIFT>
__global__ void __launch_bounds__(256) k_keys(const double2* __restrict__ pos, const double2* __restrict__ vel, int n,
                                             const GridP* __restrict__ gp, double dtH, uint32_t* __restrict__ keys,
                                             uint32_t* __restrict__ rank, uint32_t* __restrict__ cellCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GridP g = *gp;
  double2 p = pos[i];
  if (DRIFT) {  // Pos += Vel*dtHalf (sph.go:112-113): product then sum, unfused
    double2 v = vel[i];
    p.x = __dadd_rn(p.x, __dmul_rn(v.x, dtH));
    p.y = __dadd_rn(p.y, __dmul_rn(v.y, dtH));
  }
  double xs = (g.wrapx | g.framex) ? wrap_coord(p.x, g.lox, g.Lx) : p.x;
  double ys = g.wrapy ? wrap_coord(p.y, g.loy, g.Ly) : p.y;
  int cx = cell_of(xs, g.ox, g.inv_dx, g.ncx);
  int cy = cell_of(ys, g.oy, g.inv_dy, g.ncy);
  const uint32_t k = (uint32_t)cy * (uint32_t)g.ncx + (uint32_t)cx;
  keys[i] = k;
  rank[i] = atomicAdd(&cellCount[k], 1u);
}

// -------------------------------------------------------------------------------------------------
// SORT: counting sort by cell (replaces Treebuild, core.go:172-224).  Exclusive scan of the cell counts in
// three small kernels (tile sums, scan of the tile sums, tile scan + offset); no look-back spinning.
// cellStart[c] = first sorted index of cell c, for c in [0, ncell]; the counts are zeroed for the next step.
// -------------------------------------------------------------------------------------------------
#define SC_THREADS 256
#define SC_ITEMS 8
#define SC_TILE (SC_THREADS * SC_ITEMS)  // 2048 cells per block

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* wsum, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    const uint32_t t = wsum[w];
    if (w < warp) pre += t;
    tot += t;
  }
  total = tot;
  __syncthreads();
  return pre + x - v;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_tiles(const uint32_t* __restrict__ cnt, const GridP* __restrict__ gp,
                                                          uint32_t* __restrict__ tileSum) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const int m = gp->ncx * gp->ncy + 1;
  const int base = blockIdx.x * SC_TILE;
  if (base >= m) { if (threadIdx.x == 0) tileSum[blockIdx.x] = 0; return; }
  uint32_t sum = 0;
  const uint4* c4 = reinterpret_cast<const uint4*>(cnt + base) + threadIdx.x * 2;
  if (base + SC_TILE <= m) {
    const uint4 a = c4[0], b = c4[1];
    sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  } else {
    for (int k = 0; k < SC_ITEMS; ++k) { const int c = base + threadIdx.x * SC_ITEMS + k; if (c < m) sum += cnt[c]; }
  }
  uint32_t tot;
  block_excl_scan_256(sum, wsum, tot);
  if (threadIdx.x == 0) tileSum[blockIdx.x] = tot;
}

// exclusive scan of m uint32 in place, one block of 1024 threads (m = number of tiles, <= a few hundred thousand)
__global__ void __launch_bounds__(1024) k_excl_scan(uint32_t* __restrict__ a, int m_cap, const GridP* __restrict__ gp) {
  __shared__ uint32_t wsum[32];
  const int m = gp ? min(m_cap, (gp->ncx * gp->ncy + 1 + SC_TILE - 1) / SC_TILE) : m_cap;
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int per = (m + 1023) / 1024;
  int b = min(tid * per, m), e = min(b + per, m);
  uint32_t sum = 0;
  for (int i = b; i < e; ++i) sum += a[i];
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane];
    uint32_t xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    wsum[lane] = xs - w;  // exclusive
  }
  __syncthreads();
  uint32_t run = wsum[warp] + (x - sum);
  for (int i = b; i < e; ++i) {
    uint32_t v = a[i];
    a[i] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(uint32_t* __restrict__ cnt, const GridP* __restrict__ gp,
                                                          const uint32_t* __restrict__ tileOff,
                                                          uint32_t* __restrict__ cellStart) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const int m = gp->ncx * gp->ncy + 1;
  const int base = blockIdx.x * SC_TILE;
  if (base >= m) return;
  uint32_t v[SC_ITEMS];
  uint32_t sum = 0;
  const int c0 = base + threadIdx.x * SC_ITEMS;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    v[k] = (c0 + k < m) ? cnt[c0 + k] : 0u;
    sum += v[k];
  }
  uint32_t tot;
  uint32_t run = block_excl_scan_256(sum, wsum, tot) + tileOff[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    if (c0 + k < m) { cellStart[c0 + k] = run; cnt[c0 + k] = 0u; }
    run += v[k];
  }
}

__global__ void __launch_bounds__(256) k_scatter_perm(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                                                     const uint32_t* __restrict__ cellStart, int n,
                                                     uint32_t* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  perm[cellStart[keys[i]] + rank[i]] = (uint32_t)i;
}

// -------------------------------------------------------------------------------------------------
// K2: gather the SoA into cell order, predict, write the search positions and the cell table.
// MODE 0: CalculateForces() on the state as it is (VPred/EPred gathered)           sph.go:403
// MODE 1: step-0 initialisation: VPred = Vel, EPred = E                            sph.go:97-100
// MODE 2: drift-1 + predict                                                        sph.go:108-117
// -------------------------------------------------------------------------------------------------
struct StateIn {
  const double2 *pos, *vel, *vdot, *vpred;
  const double *e, *edot, *epred;
  const int64_t* id;
  const double4* pc;  // {rho, c, h, P}
  const uint8_t* ghost;
};
struct StateOut {
  double2 *pos, *vel, *vdot, *vpred;
  double *e, *edot, *epred;
  int64_t* id;
  double4* pc;
  uint8_t* ghost;
  double2* spos;   // wrapped search positions
  double* hguess;  // h of the previous evaluation in the new order (0 = unknown)
};

// LEAN: a force evaluation follows (it overwrites VDot / EDot, and the kNN kernel overwrites {rho, c, h, P}), so those
// fields are neither copied nor, unless the predictor needs them, read.
template <int MODE, bool LEAN>
__global__ void __launch_bounds__(256) k_reorder(StateIn in, StateOut out, const uint32_t* __restrict__ keys,
                                                const uint32_t* __restrict__ perm, int n, const GridP* __restrict__ gp, double dtH,
                                                const uint32_t* __restrict__ cellStart, uint32_t* __restrict__ keysSorted,
                                                uint32_t* __restrict__ inv) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const GridP g = *gp;
  const uint32_t i = perm[t];
  const uint32_t c = keys[i];
  // position inside the cell = number of members with a smaller previous index (deterministic order);
  // cells crowded beyond 64 (clamped border cells of an open box) keep the arbitrary atomic order
  const uint32_t s = cellStart[c], e = cellStart[c + 1];
  uint32_t j = (uint32_t)t;
  if (e - s <= 64u) {
    j = s;
    for (uint32_t u = s; u < e; ++u) j += (perm[u] < i) ? 1u : 0u;
  }
  double2 p = in.pos[i], v = in.vel[i], a = make_double2(0.0, 0.0);
  double e_ = in.e[i], ed = 0.0;
  if (MODE == 2 || !LEAN) { a = in.vdot[i]; ed = in.edot[i]; }
  double4 pc = in.pc[i];
  double2 vp;
  double ep;
  if (MODE == 2) {
    p.x = __dadd_rn(p.x, __dmul_rn(v.x, dtH));
    p.y = __dadd_rn(p.y, __dmul_rn(v.y, dtH));
    vp.x = __dadd_rn(v.x, __dmul_rn(a.x, dtH));
    vp.y = __dadd_rn(v.y, __dmul_rn(a.y, dtH));
    ep = __dadd_rn(e_, __dmul_rn(ed, dtH));
  } else if (MODE == 1) {
    vp = v;
    ep = e_;
  } else {
    vp = in.vpred[i];
    ep = in.epred[i];
  }
  out.pos[j] = p;
  out.vel[j] = v;
  out.e[j] = e_;
  if (!LEAN) { out.vdot[j] = a; out.edot[j] = ed; out.pc[j] = pc; }
  out.vpred[j] = vp;
  out.epred[j] = ep;
  out.id[j] = in.id[i];
  out.ghost[j] = in.ghost[i];
  out.hguess[j] = pc.z;
  double2 sp;
  sp.x = (g.wrapx | g.framex) ? wrap_coord(p.x, g.lox, g.Lx) : p.x;
  sp.y = g.wrapy ? wrap_coord(p.y, g.loy, g.Ly) : p.y;
  out.spos[j] = sp;
  keysSorted[j] = c;
  if (inv) inv[i] = j;  // where the particle went (ring: halo sources / ghost destinations of the reuse evaluations)
}

// -------------------------------------------------------------------------------------------------
// K3: exact kNN (k = 32) + density + sound speed     (nearest-neighbour.go:28-165, sph.go:306-323,423-429)
//
// One warp = one tile = 32 consecutive particles of the cell-sorted order (a strip of one grid row), one
// query per lane.  Every lane has a search radius rg (previous h plus a margin, or a density estimate).
// The warp walks the UNION of its lanes' stencils row by row; a row of the union is a contiguous particle
// range, which the warp stages into shared memory (coalesced) as fp32 coordinates relative to the tile.
//   phase 1 (filter, fp32): every lane tests every staged candidate (shared-memory broadcast) and appends
//           the ones with d2f < rg^2 (1 + delta) to its private shared-memory column {fp32 key, entry}.
//   select  (fp32 keys)   : the cnt - 32 largest keys are removed (cnt is 33..40 with a tight rg).
//   phase 2 (exact, fp64) : the 32 survivors get d^2 exactly as the reference computes it (unfused
//           (p + off) - b, linear-algebra.go:61-64); h^2 = max; density and sound speed follow.
// Exactness certificate per lane (else the particle goes to the ring-expansion fallback):
//   32 <= cnt <= CAP, the fp32 gap between the 32nd and 33rd key exceeds the fp32 error bound delta
//   (no rank ambiguity, exact ties included), and exact h^2 <= rg^2 (everything within h was staged).
// fp32 error bound: coordinates relative to the tile, |v| <= V: each carries <= 2^-24 V rounding, so
//   |d2f - d2| / d2 <= ~4 * 2^-24 * V / d + 3 * 2^-24; delta = 3 * (2^-22 * V / rg + 2^-21) covers it
//   with a factor > 2 to spare near d ~ rg.
// -------------------------------------------------------------------------------------------------
#define KNN_WARPS 2
#define KNN_THREADS (KNN_WARPS * 32)

#define HACC_N 256  // spread slots of the smoothing-length accumulator (keeps the atomics uncontended)
struct KnnOut {
  // fully periodic, single-handle runs: sum of h in fixed point (h * hscale, integer => order-independent, so the next
  // grid is bit-reproducible) and particle count, [HACC_N][2]; the next k_make_grid takes the mean h from here
  // instead of a statistics pass over the state.  hscale = 0: disabled.
  unsigned long long* hacc;
  double hscale;
  uint32_t* qmax;   // [0]: max h of the owned particles as float bits, rounded up (monotone for non-negative floats)
  double4* pc;      // {rho, c, h, P = c^2/(gamma rho)}
  uint32_t* nn;     // [tile][slot][lane], entry = index | image code << 28
  int* failList;
  int* failCount;
};

// shared memory per warp: the staged candidates of the tile {fp64 position (exact phase), fp32 tile-relative
// position (filter), list entry} = 28 B each, and a column of CAP slots x 32 lanes x {fp32 key, staged slot}
__host__ __device__ inline size_t knn_smem_bytes_per_warp(int cap, int ncw, bool f32) {
  return (size_t)cap * 256 + (size_t)ncw * (f32 ? 12 : 28);  // the fp32 build stages no fp64 positions
}

__device__ __forceinline__ int warp_min_i(int v, uint32_t mask) {
  return __reduce_min_sync(0xffffffffu, (mask >> (threadIdx.x & 31)) & 1u ? v : 0x7fffffff);
}
__device__ __forceinline__ int warp_max_i(int v, uint32_t mask) {
  return __reduce_max_sync(0xffffffffu, (mask >> (threadIdx.x & 31)) & 1u ? v : (int)0x80000000);
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// image code of a list entry: bits 2-3 = ix + 1, bits 0-1 = iy + 1 (ix, iy in {-1, 0, 1})
__device__ __forceinline__ uint32_t img_code(int ix, int iy) { return (uint32_t)(((ix + 1) << 2) | (iy + 1)); }

// branch-free append of one candidate to the lane's column: kp is the shared-space byte address of the next
// free slot (stride 256 B).  The caller checks the remaining room once per 8 candidates.
__device__ __forceinline__ void knn_append(uint32_t& kp, float d2f, uint32_t en, float thr) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 kb;\n\t"
      "setp.lt.f32 p, %1, %3;\n\t"
      "mov.b32 kb, %1;\n\t"
      "@p st.shared.v2.b32 [%0], {kb, %2};\n\t"
      "@p add.u32 %0, %0, 256;\n\t"
      "}\n"
      : "+r"(kp)
      : "f"(d2f), "r"(en), "f"(thr));
}

// Selection on a lane's column of fp32 keys (bit patterns): the mrem largest keys are to be dropped.  Five keys per pass
// (a max/min insertion network); T = smallest dropped key (0xffffffff: none), akey = largest kept key.  Warp-collective:
// lanes with active == false run along with zero trips.
__device__ __forceinline__ void knn_select_drop(const uint2* col, int cnt, int mrem, bool active, uint32_t& T, uint32_t& akey) {
  uint32_t bound = 0xffffffffu;
  T = 0xffffffffu; akey = 0u;
  bool sel_done = !active;
  while (__any_sync(0xffffffffu, !sel_done)) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    const int lim = sel_done ? 0 : cnt;
    for (int s = 0; s < lim; ++s) {
      uint32_t k = col[s * 32].x;
      k = k < bound ? k : 0u;
      uint32_t a;
      a = max(t0, k); k = min(t0, k); t0 = a;
      a = max(t1, k); k = min(t1, k); t1 = a;
      a = max(t2, k); k = min(t2, k); t2 = a;
      a = max(t3, k); k = min(t3, k); t3 = a;
      t4 = max(t4, k);
    }
    if (!sel_done) {
      if (mrem <= 4) {
        T = mrem == 0 ? bound : (mrem == 1 ? t0 : (mrem == 2 ? t1 : (mrem == 3 ? t2 : t3)));
        akey = mrem == 0 ? t0 : (mrem == 1 ? t1 : (mrem == 2 ? t2 : (mrem == 3 ? t3 : t4)));
        sel_done = true;
      } else {
        mrem -= 5;
        bound = t4;
      }
    }
  }
}

// warp-level contribution to the smoothing-length accumulator (all 32 lanes call): h * hscale < 2^24 per lane, so the
// warp sum fits 32 bits and is one REDUX
__device__ __forceinline__ void knn_accumulate_h(const KnnOut& out, bool ok, bool owned, double h) {
  if (out.hscale == 0.0) return;
  const unsigned q = __reduce_add_sync(0xffffffffu, ok ? (unsigned)__double2uint_rn(h * out.hscale) : 0u);
  const unsigned cnt = __popc(__ballot_sync(0xffffffffu, ok));
  const unsigned hm = __reduce_max_sync(0xffffffffu, (ok && owned) ? __float_as_uint(__double2float_ru(h)) : 0u);
  if ((threadIdx.x & 31) == 0 && cnt) {
    unsigned long long* a = out.hacc + 2 * (blockIdx.x & (HACC_N - 1));
    atomicAdd(a, (unsigned long long)q);
    atomicAdd(a + 1, (unsigned long long)cnt);
    if (hm) atomicMax(out.qmax, hm);
  }
}

// radius up to which a rebuild collects candidates: h_prev (1 + skin), from the search radius rg = h_prev (1 + margin)
__device__ __forceinline__ double knn_ext_radius(double rg, double margin, double skin) { return rg * ((1.0 + skin) / (1.0 + margin)); }

// unwrapped cell range containing every point within r of (xa, ya) (clamped to one period either side)
__device__ __forceinline__ void knn_cell_range(const GridP& g, double xa, double ya, double r, int cxa, int cya, int& clo, int& chi,
                                               int& rlo, int& rhi) {
  clo = (int)floor((xa - r - g.ox) * g.inv_dx); chi = (int)floor((xa + r - g.ox) * g.inv_dx);
  rlo = (int)floor((ya - r - g.oy) * g.inv_dy); rhi = (int)floor((ya + r - g.oy) * g.inv_dy);
  clo = max(min(clo, cxa), -g.ncx); chi = min(max(chi, cxa), 2 * g.ncx - 1);
  rlo = max(min(rlo, cya), -g.ncy); rhi = min(max(rhi, cya), 2 * g.ncy - 1);
}

template <int KERNEL, bool F32, bool EXT>
__global__ void __launch_bounds__(KNN_THREADS, 5) k_knn_tile(const double2* __restrict__ spos,
                                                         const uint32_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ cellStart,
                                                         const double* __restrict__ hguess,
                                                         const double* __restrict__ epred, int n,
                                                         const GridP* __restrict__ gp, PhysP ph, KnnTune tune,
                                                         KnnOut out, const uint8_t* __restrict__ gflag,
                                                         uint32_t* __restrict__ dflags, KnnExt ex) {
  const GridP g = *gp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CAP = tune.cap, NCW = tune.ncw;
  unsigned char* wb = smem_raw + (size_t)warp * knn_smem_bytes_per_warp(CAP, NCW, F32);
  uint2* col = reinterpret_cast<uint2*>(wb) + lane;                                  // [slot * 32] = {key, staged slot}
  const size_t cbase = (size_t)CAP * 256;
  float2* candF = reinterpret_cast<float2*>(wb + cbase);                             // fp32 tile-relative positions
  const float4* candF4 = reinterpret_cast<const float4*>(candF);                     // two staged candidates each
  uint32_t* candE = reinterpret_cast<uint32_t*>(wb + cbase + (size_t)NCW * 8);       // index | image code << 28
  double2* candD = reinterpret_cast<double2*>(wb + cbase + (size_t)NCW * 12);        // exact positions (fp64 build)
  const uint32_t kbase = (uint32_t)__cvta_generic_to_shared(col);
  const uint32_t klim = kbase + (uint32_t)(CAP - 8) * 256u;  // beyond this fewer than 8 free slots remain

  const int tile = blockIdx.x * KNN_WARPS + warp;
  if (tile * 32 >= n) return;  // whole warp out of range (warp-uniform)
  const int i = tile * 32 + lane;
  // queries: owned particles and (slab mode) inner ghosts; outer ghosts are candidates only
  const bool valid = i < n && (gflag == nullptr || gflag[i] != GF_OUTER);
  const bool owned = i < n && (gflag == nullptr || gflag[i] == GF_OWNED);
  // an outer ghost is a candidate only: its {rho, c, h, P} are never evaluated.  rho = -1 marks it where the force kernel
  // will find it (its bulk-staged records are raw copies of these rows): a pair with such a neighbour is refused
  if (i < n && !valid) out.pc[i] = make_double4(-1.0, 0.0, 0.0, 0.0);

  double xa = 0, ya = 0, rg = 0, ep = 0;
  int cxa = 0, cya = 0;
  if (valid) {
    const double2 p = spos[i];
    xa = p.x; ya = p.y;
    const uint32_t k = keys[i];
    const double hp = hguess[i];
    ep = epred[i];
    cya = (int)(k / (uint32_t)g.ncx);
    cxa = (int)(k - (uint32_t)cya * (uint32_t)g.ncx);
    if (hp > 0.0) {
      rg = hp * (1.0 + tune.guess_margin);
    } else {  // density estimate from the block of cells around the particle
      const int wx = max(1, (int)(1.5 * g.dy * g.inv_dx));
      const int x0 = max(cxa - wx, 0), x1 = min(cxa + wx, g.ncx - 1);
      const int y0 = max(cya - 1, 0), y1 = min(cya + 1, g.ncy - 1);
      uint32_t c = 0;
      for (int r = y0; r <= y1; ++r) c += cellStart[r * g.ncx + x1 + 1] - cellStart[r * g.ncx + x0];
      const double area = (double)(x1 - x0 + 1) * g.dx * (double)(y1 - y0 + 1) * g.dy;
      rg = sqrt(tune.k_target * area / (3.141592653589793 * (double)(c > 0 ? c : 1)));
    }
  }
  const double rg2 = rg * rg;
  // unwrapped cell range that contains every point within rw = rg (1 + 1e-4): the lane scans only these cells.
  // The widening matters in the fp32 build: a candidate outside them has a true d^2 > rg^2 (1 + 2e-4), so even
  // with its fp32 key error (< 1.5e-5 relative, enforced below) it cannot undercut an accepted h^2 < thr.
  const double rw = rg * (1.0 + 1e-4);
  int clo = (int)floor((xa - rw - g.ox) * g.inv_dx), chi = (int)floor((xa + rw - g.ox) * g.inv_dx);
  int rlo = (int)floor((ya - rw - g.oy) * g.inv_dy), rhi = (int)floor((ya + rw - g.oy) * g.inv_dy);
  // (clamped to one period either side: non-finite or absurd positions must not overflow the range arithmetic)
  clo = max(min(clo, cxa), -g.ncx); chi = min(max(chi, cxa), 2 * g.ncx - 1);
  rlo = max(min(rlo, cya), -g.ncy); rhi = min(max(rhi, cya), 2 * g.ncy - 1);
  // EXT: the staged block of the whole tile is taken wide enough for the skin (rgx = h_prev (1 + skin)) and kept for the
  // cycle's reuse evaluations; the lanes' own windows stay those of rg
  const double rgx = EXT ? knn_ext_radius(rg, tune.guess_margin, ex.skin) : rg;
  int xlo = clo, xhi = chi, ylo = rlo, yhi = rhi;
  if (EXT) knn_cell_range(g, xa, ya, rgx * (1.0 + 1e-4), cxa, cya, xlo, xhi, ylo, yhi);
  bool ext_try = EXT;   // first pass: all lanes of the tile as one group on the wide block

  // One pass of the whole pipeline per group of lanes that sit in the same grid row (a tile is a strip of one
  // row, so there is one group unless the tile straddles the end of a row).
  // Where the particles are sparse for the grid (h well above the mean) the union block of 32 lanes may not fit
  // the staging area: the group is then halved (lanes are ordered along the strip) and retried.
  uint32_t todo = __ballot_sync(0xffffffffu, valid);
  int maxlanes = 32;
  if (EXT && lane == 0) ex.tinfo[tile].npc = 0;  // (set again below if the tile gets a shared block)
  while (todo) {
    const int lead = __ffs(todo) - 1;
    const int grow = __shfl_sync(0xffffffffu, cya, lead);
    const uint32_t cand = ext_try ? todo : (__ballot_sync(0xffffffffu, valid && cya == grow) & todo);
    const uint32_t grp = __ballot_sync(0xffffffffu, ((cand >> lane) & 1u) && __popc(cand & ((1u << lane) - 1u)) < maxlanes);
    todo &= ~grp;
    const bool mine = (grp >> lane) & 1u;
    bool bad = false;  // stencil wider than the period / fp32 bound not applicable / staging area full: fallback
    int c0 = warp_min_i(ext_try ? xlo : clo, grp), c1 = warp_max_i(ext_try ? xhi : chi, grp);
    int r0 = warp_min_i(ext_try ? ylo : rlo, grp), r1 = warp_max_i(ext_try ? yhi : rhi, grp);
    bool gbad = false;
    if (g.wrapx) { if (c1 - c0 + 1 > g.ncx) gbad = true; }
    else { c0 = max(c0, 0); c1 = min(c1, g.ncx - 1); c0 = min(c0, g.ncx - 1); c1 = max(c1, 0); }
    if (g.wrapy) { if (r1 - r0 + 1 > g.ncy) gbad = true; }
    else { r0 = max(r0, 0); r1 = min(r1, g.ncy - 1); r0 = min(r0, g.ncy - 1); r1 = max(r1, 0); }

    // the pieces of the union block: (row, x image) -> a contiguous range of the sorted order; lane p fetches piece p
    const int ix0 = g.wrapx ? img_idx(c0, g.ncx) : 0, ix1 = g.wrapx ? img_idx(c1, g.ncx) : 0;
    const int nix = ix1 - ix0 + 1, npc = (r1 - r0 + 1) * nix;
    bool big = !gbad && npc > 32;  // does not fit: more pieces than lanes / more candidates than staging slots
    int p_s = 0, p_len = 0, p_off = 0, nst = 0;
    uint32_t p_code = 0;
    if (!gbad && !big) {
      if (lane < npc) {
        const int rr = lane / nix, ix = ix0 + (lane - rr * nix), ru = r0 + rr;
        const int iy = g.wrapy ? img_idx(ru, g.ncy) : 0;
        const int row = ru - iy * g.ncy;
        const int a = max(c0, ix * g.ncx) - ix * g.ncx, b = min(c1, ix * g.ncx + g.ncx - 1) - ix * g.ncx;
        p_s = (int)cellStart[row * g.ncx + a];
        p_len = (int)cellStart[row * g.ncx + b + 1] - p_s;
        p_code = img_code(ix, iy);
      }
      int incl = (p_len + 7) & ~7;  // every piece is padded to 8 with unreachable dummies
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      p_off = incl - ((p_len + 7) & ~7);
      nst = __shfl_sync(0xffffffffu, incl, 31);
      if (nst > NCW) big = true;
    }
    if (ext_try && (big || gbad)) {  // no shared wide block for this tile: the ordinary search, no reuse for its particles
      ext_try = false; todo |= grp;
      continue;
    }
    if (big) {
      const int nl = __popc(grp);
      if (nl > 1) { todo |= grp; maxlanes = nl >> 1; continue; }  // retry with half the lanes
      gbad = true;  // a single query whose stencil does not fit: ring-expansion fallback
    }
    if (gbad) {
      if (mine) { const int slot = atomicAdd(out.failCount, 1); out.failList[slot] = i; }
      continue;
    }

    // fp32 frame of this group: origin at the centre of the union block; V bounds every |relative coordinate|
    const double xref = g.ox + 0.5 * (double)(c0 + c1 + 1) * g.dx;
    const double yref = g.oy + 0.5 * (double)(r0 + r1 + 1) * g.dy;
    double V = fmax(0.5 * (double)(c1 - c0 + 1) * g.dx, 0.5 * (double)(r1 - r0 + 1) * g.dy);
    // clamped border cells of an open axis may hold particles beyond the block: bound by the queries' reach
    V = fmax(V, warp_max_d(mine ? fmax(fabs(xa - xref), fabs(ya - yref)) + (ext_try ? rgx : rg) : 0.0));
    const float qfx = (float)(xa - xref), qfy = (float)(ya - yref);
    const float2 nqx2 = make_float2(-qfx, -qfx), nqy2 = make_float2(-qfy, -qfy);
    const double delta = 3.0 * (2.384185791015625e-07 * V / fmax(rg, 1e-300) + 4.76837158203125e-07);
    const float deltaf = (float)delta;
    // tile far wider than this lane's radius: no useful fp32 bound (fp32 build: accuracy of h itself)
    if (mine && !(delta < (F32 ? 1.5e-5 : 1e-3))) bad = true;
    const float thr0 = (float)(rg2 * (1.0 + delta)) * 1.0000002f;
    float thrf = (mine && !bad) ? thr0 : -1.0f;

    const float Vf = (float)V * 1.000001f;

    // window of piece 0 (see the filter below): issued here so that its two lookups overlap the staging loads
    int w_lo = 0, w_hi = 0;
    {
      const int ru = r0, ix = ix0;
      const int iy = g.wrapy ? img_idx(ru, g.ncy) : 0, row = ru - iy * g.ncy, uL = ix * g.ncx;
      const int a = max(c0, uL) - uL, b = min(c1, uL + g.ncx - 1) - uL;
      const int ca = max(clo, a + uL) - uL, cb = min(chi, b + uL) - uL;
      const int ps = __shfl_sync(0xffffffffu, p_s, 0), po = __shfl_sync(0xffffffffu, p_off, 0);
      if (mine && ru >= rlo && ru <= rhi && ca <= cb) {
        w_lo = (int)cellStart[row * g.ncx + ca] - ps + po;
        w_hi = (int)cellStart[row * g.ncx + cb + 1] - ps + po;
      } else { w_lo = po; w_hi = po; }
    }
    // stage the whole union block (coalesced loads within a piece)
    bool anyimg = false;  // some staged piece is a periodic image (warp-uniform)
    int self_slot = -1;
    __syncwarp();
    for (int pc_ = 0; pc_ < npc; ++pc_) {
      const int s = __shfl_sync(0xffffffffu, p_s, pc_), len = __shfl_sync(0xffffffffu, p_len, pc_);
      const int off = __shfl_sync(0xffffffffu, p_off, pc_);
      const uint32_t code = __shfl_sync(0xffffffffu, p_code, pc_);
      const int ix = (int)(code >> 2) - 1, iy = (int)(code & 3u) - 1;
      anyimg |= (code != 5u);
      // candidate position in the query frame: b + img * L  (the reference shifts the query by -img * L)
      const double sx = (double)ix * g.Lx - xref, sy = (double)iy * g.Ly - yref;
      const int len8 = (len + 7) & ~7;
      for (int t = lane; t < len8; t += 32) {
        float fx = 3.0e18f, fy = 3.0e18f;
        if (t < len) {
          const double2 pb = spos[s + t];
          fx = (float)(pb.x + sx); fy = (float)(pb.y + sy);
          // beyond the bounded block (clamped border cell): farther than every lane's reach, drop it
          if (!(fabsf(fx) <= Vf && fabsf(fy) <= Vf)) { fx = 3.0e18f; fy = 3.0e18f; }
          if (!F32) candD[off + t] = pb;
          candE[off + t] = (uint32_t)(s + t) | (code << IMG_SHIFT);
        }
        float* cf = reinterpret_cast<float*>(candF) + (size_t)((off + t) >> 1) * 4 + ((off + t) & 1);
        cf[0] = fx; cf[2] = fy;  // pairs of candidates as {x0, x1, y0, y1}: operands of the packed fp32 filter
      }
      if (mine && code == 5u && i >= s && i < s + len) self_slot = off + (i - s);
    }
    __syncwarp();

    uint32_t kp = kbase;
    bool ovf = false;  // column (nearly) full: stop appending, the lane goes to the fallback
    // phase 1: fp32 filter, branch-free.  A lane only scans its own window of every piece: the candidates in the
    // cells its search disc touches (a contiguous slot range, from two cellStart lookups per piece, fetched one
    // piece ahead).  The trip count is the longest window of the warp; a window that would run past the end of
    // its piece is shifted back, so every staged slot is seen at most once per lane.
    for (int pc_ = 0; pc_ < npc; ++pc_) {
      const int ws = w_lo, we = w_hi;
      const int po = __shfl_sync(0xffffffffu, p_off, pc_), pl = __shfl_sync(0xffffffffu, p_len, pc_);
      if (pc_ + 1 < npc) {  // fetch the window of the next piece while this one is scanned
        const int q = pc_ + 1, rr = q / nix, ix = ix0 + (q - rr * nix), ru = r0 + rr;
        const int iy = g.wrapy ? img_idx(ru, g.ncy) : 0, row = ru - iy * g.ncy, uL = ix * g.ncx;
        const int a = max(c0, uL) - uL, b = min(c1, uL + g.ncx - 1) - uL;
        const int ca = max(clo, a + uL) - uL, cb = min(chi, b + uL) - uL;
        const int ps = __shfl_sync(0xffffffffu, p_s, q), pq = __shfl_sync(0xffffffffu, p_off, q);
        if (mine && ru >= rlo && ru <= rhi && ca <= cb) {
          w_lo = (int)cellStart[row * g.ncx + ca] - ps + pq;
          w_hi = (int)cellStart[row * g.ncx + cb + 1] - ps + pq;
        } else { w_lo = pq; w_hi = pq; }
      }
      const int wal = ws & ~1;  // two candidates per 16-byte load
      const int trips = __reduce_max_sync(0xffffffffu, (we - wal + 7) >> 3);
      const int pend8 = po + ((pl + 7) & ~7);
      int c = min(wal, pend8 - 8 * trips);  // >= po: a window lies inside its piece
      for (int k = 0; k < trips; ++k, c += 8) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = candF4[(c >> 1) + u];
        float d2f[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // two candidates per instruction (FADD2 / FMUL2 / FFMA2)
          const float2 ax = __fadd2_rn(make_float2(v[u].x, v[u].y), nqx2), ay = __fadd2_rn(make_float2(v[u].z, v[u].w), nqy2);
          const float2 d2 = __ffma2_rn(ay, ay, __fmul2_rn(ax, ax));
          d2f[2 * u] = d2.x;
          d2f[2 * u + 1] = d2.y;
        }
        if (kp > klim) { ovf = true; thrf = -1.0f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) knn_append(kp, d2f[u], (uint32_t)(c + u), thrf);
      }
    }
    __syncwarp();

    // The column holds cnt entries, one of them the lane itself (d2f = 0, centre image; self is excluded in
    // every image, nearest-neighbour.go:79).  m = cnt - 33 entries with the largest keys must be dropped.
    const int cnt = (int)((kp - kbase) >> 8);
    bool ok = mine && !bad && !ovf && cnt >= SPHB_K + 1 && cnt <= CAP - 8;  // (the compaction pads up to cnt + 7)
    // select A: the 5 largest keys below `bound` per pass (a max/min insertion network on the fp32 bit patterns;
    // the warp-wide maximum of m is 4 on average at a 2 % margin, so one pass usually does).
    // T = smallest dropped key, akey = largest kept key.
    uint32_t T, akey;
    knn_select_drop(col, cnt, ok ? cnt - (SPHB_K + 1) : 0, ok, T, akey);
    // fp64 build, rank ambiguity (includes exact ties): the smallest dropped key must exceed the largest kept one
    // by more than the fp32 error; with nothing dropped the bound is the acceptance threshold itself (checked
    // as h^2 <= rg^2).  fp32 build: the fp32 keys are the distances, ties may fall either way.
    if (!F32 && ok && T != 0xffffffffu && !(__uint_as_float(T) > __uint_as_float(akey) * (1.0f + deltaf) * 1.000001f)) ok = false;

    // select B: compact the kept entries (key < T, not self) to the head of the column, so that the next phase
    // runs over exactly 32 slots with independent loads in flight
    int kept = 0;
    bool self_seen = false;
    {
      const int lim = ok ? cnt : 0;
      // pad the column to a multiple of eight with keys that are never kept (slots up to cnt + 7 <= CAP - 1 exist:
      // the overflow test of the filter leaves that room)
      if (ok) {
#pragma unroll
        for (int u = 0; u < 8; ++u) col[(cnt + u) * 32].x = 0xffffffffu;
      }
      for (int s0 = 0; s0 < lim; s0 += 8) {  // eight loads in flight, then the (aliasing) compacted stores
        uint2 ke[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ke[u] = col[(s0 + u) * 32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool self = ke[u].y == (uint32_t)self_slot;  // (a padding slot holds a stale slot id but the lane's own
          self_seen |= self && ke[u].x != 0xffffffffu;        //  entry is always inside the first cnt)
          if (ke[u].x < T && !self) {  // at most cnt - 1 - m = 32 entries qualify; kept <= s0 + u: only slots already read
            col[kept * 32] = ke[u];
            ++kept;
          }
        }
      }
    }
    // exactly 32 kept (fails on key ties at a pass boundary; more than 32 cannot happen: at most cnt - 1 - m)
    if (ok && !(kept == SPHB_K && self_seen)) ok = false;
    uint32_t* np = out.nn + (size_t)tile * 32 * 32 + lane;  // list slots of this lane (stride 32 words)
    if (F32) {
      // ---- fp32 build: the keys are the squared distances; list, h, density straight from the column
      float h2f = 0.0f;
      if (ok) {
#pragma unroll 8
        for (int s = 0; s < SPHB_K; ++s) h2f = fmaxf(h2f, __uint_as_float(col[s * 32].x));
        if (!(h2f < thr0)) ok = false;  // everything below the acceptance threshold was seen
      }
      if (mine && !ok) {
        const int slot = atomicAdd(out.failCount, 1);
        out.failList[slot] = i;
      }
      if (ok) {
        float inv_h;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_h) : "f"(h2f));
        const float hf = h2f * inv_h;
        float acc = 0.0f;
#pragma unroll 8
        for (int s = 0; s < SPHB_K; ++s) {
          const uint2 ke = col[s * 32];
          np[s * 32] = candE[ke.y];
          const float q2 = __uint_as_float(ke.x) * (inv_h * inv_h);
          float rq;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(fmaxf(q2, 1e-30f)));
          const float q = fminf(q2 * rq, 1.0f);
          if (KERNEL == 0) acc += 1.0f;
          else if (KERNEL == 1) {
            const float lo = fmaf(q * q, q - 1.0f, 1.0f / 6.0f);
            const float t = 1.0f - q;
            acc += q < 0.5f ? lo : t * t * t * (1.0f / 3.0f);
          } else {
            const float t = 1.0f - q, t2 = t * t;
            acc += t2 * t2 * fmaf(4.0f, q, 1.0f);
          }
        }
        const double h = (double)hf;
        if (g.sides && (((g.sides & 1) && xa - g.ox < h) || ((g.sides & 2) && g.ox + (double)g.ncx * g.dx - xa < h)))
          atomicOr(dflags, DFLAG_GHOST_THIN);
        // Density2D (sph.go:322), sound speed (sph.go:426-428), pressure term c^2/(gamma rho) (sph.go:332,360)
        const float rho = (float)(ph.Fpref * ph.mass) * acc * (inv_h * inv_h);
        const float c = sqrtf((float)(ph.cfac * ep));
        out.pc[i] = make_double4((double)rho, (double)c, h, (double)(c * c / ((float)ph.gamma * rho)));
      }
      const double hacc_h = (double)(ok ? h2f * rsqrtf(fmaxf(h2f, 1e-37f)) : 0.0f);
      if (EXT && ext_try) {  // the shared block of the tile, for the annulus pass and the cycle's reuse evaluations
        // for the annulus pass: the smallest key the selection did not take (nothing dropped: the acceptance threshold)
        if (ok) ex.dexcl[i] = (double)(T != 0xffffffffu ? __uint_as_float(T) : thr0);
        if (lane < npc) ex.ptab[(size_t)tile * 32 + lane] = make_uint2((uint32_t)p_s | (p_code << IMG_SHIFT), (uint32_t)p_len);
        if (lane == 0) ex.tinfo[tile] = TileInfo{npc, nst, c0, c1, r0, r1};
      }
      knn_accumulate_h(out, ok, owned, hacc_h);
      continue;
    }
    // ---- fp64 build, exact phase: d^2 exactly as the reference computes it; the list entry goes to global
    // memory, the slot is reused for the exact d^2
    double h2 = 0.0;
    if (ok) {
      if (!anyimg) {  // warp-uniform: no periodic image in this tile's block, the query is never shifted
#pragma unroll
        for (int s0 = 0; s0 < SPHB_K; s0 += 8) {
          uint32_t en[8], sls[8];
          double2 pb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t sl = col[(s0 + u) * 32].y;
            sls[u] = sl;
            pb[u] = candD[sl];
            en[u] = candE[sl];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double d2 = dist_sq(xa - pb[u].x, ya - pb[u].y);
            h2 = fmax(h2, d2);
            np[(s0 + u) * 32] = en[u];
            *reinterpret_cast<double*>(&col[(s0 + u) * 32]) = d2;
          }
        }
      } else {
        const double qxm = __dadd_rn(xa, g.Lx), qxp = __dadd_rn(xa, -g.Lx);  // ix = -1 / +1: query + (-ix * L)
        const double qym = __dadd_rn(ya, g.Ly), qyp = __dadd_rn(ya, -g.Ly);
#pragma unroll
        for (int s0 = 0; s0 < SPHB_K; s0 += 8) {
          uint32_t en[8], sls[8];
          double2 pb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t sl = col[(s0 + u) * 32].y;
            sls[u] = sl;
            pb[u] = candD[sl];
            en[u] = candE[sl];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t cx = (en[u] >> (IMG_SHIFT + 2)) & 3u, cy = (en[u] >> IMG_SHIFT) & 3u;
            const double qx = cx == 1u ? xa : (cx == 0u ? qxm : qxp);
            const double qy = cy == 1u ? ya : (cy == 0u ? qym : qyp);
            const double d2 = dist_sq(qx - pb[u].x, qy - pb[u].y);
            h2 = fmax(h2, d2);
            np[(s0 + u) * 32] = en[u];
            *reinterpret_cast<double*>(&col[(s0 + u) * 32]) = d2;
          }
        }
      }
      // nothing within h can have been missed
      if (!(h2 <= rg2)) ok = false;
    }
    if (mine && !ok) {
      const int slot = atomicAdd(out.failCount, 1);
      out.failList[slot] = i;
    }
    if (ok) {
      const double inv_h = fast_rsqrt(h2);
      const double h = fast_sqrt(h2, inv_h);
      if (g.sides && (((g.sides & 1) && xa - g.ox < h) || ((g.sides & 2) && g.ox + (double)g.ncx * g.dx - xa < h)))
        atomicOr(dflags, DFLAG_GHOST_THIN);
      double acc = 0.0;
#pragma unroll 4
      for (int s = 0; s < SPHB_K; ++s) {
        const double d2 = *reinterpret_cast<const double*>(&col[s * 32]);
        const double d = d2 * fast_rsqrt(d2 + 1e-300);  // coincident particles: d = 0
        acc += kern_F<KERNEL>(d * inv_h);                    // d <= h up to rounding
      }
      // Density2D (sph.go:322), sound speed (sph.go:426-428), pressure term c^2/(gamma rho) (sph.go:332,360)
      const double rho = ph.Fpref * ph.mass * acc * (inv_h * inv_h);
      const double c2 = ph.cfac * ep;
      const double c = c2 > 0.0 ? fast_sqrt(c2, fast_rsqrt(c2)) : sqrt(c2);
      out.pc[i] = make_double4(rho, c, h, c * c * fast_rcp(ph.gamma * rho));
    }
    if (EXT && ext_try) {
      if (ok) ex.dexcl[i] = (double)(T != 0xffffffffu ? __uint_as_float(T) : thr0);
      if (lane < npc) ex.ptab[(size_t)tile * 32 + lane] = make_uint2((uint32_t)p_s | (p_code << IMG_SHIFT), (uint32_t)p_len);
      if (lane == 0) ex.tinfo[tile] = TileInfo{npc, nst, c0, c1, r0, r1};
    }
    knn_accumulate_h(out, ok, owned, ok ? sqrt(h2) : 0.0);
  }
}

// Fallback: one WARP per failed particle, lanes = candidates, ring expansion until the result is certified
// exact.  Handles everything the tile kernel refuses: too few / too many candidates inside the guess, fp32
// rank ambiguity and exact ties, stencils wider than the period (several images of the same particle, like
// the reference's 3x3 image loop).  The bounded top-32 is the reference's sorted queue
// (nearest-neighbour.go:139-153) distributed over the warp: lane s holds slot s, descending, lane 0 = h^2;
// an insertion is one ballot + one shuffle-down.  Candidates are inserted in scan order with the
// reference's strict comparisons, so equal keys end up in the same relative order.
// STALE (reuse evaluations): the cell table describes the positions of the last rebuild.  The block of cells is then
// taken around the query's own (stale) cell and widened by D, the bound of the relative displacement since the build:
// a particle now within d of the query was within d + D of it then, and the query was inside its cell.  A particle may
// have crossed the periodic seam since (its wrapped position jumped by a period while its cell did not): the image of a
// candidate is therefore corrected to the nearest one - unless the block spans more than half a period, in which case
// all three images of every cell are scanned (exact whatever the cells say: positions are the current ones).
struct FbExt {
  double* dexcl;           // set to 0 = "no further candidates known" for the particles searched here (nullptr: no reuse bookkeeping)
  const ReuseState* rs;    // STALE: D
};

template <int KERNEL, bool STALE>
__global__ void __launch_bounds__(128) k_knn_fallback(const double2* __restrict__ spos,
                                                     const uint32_t* __restrict__ keys,
                                                     const uint32_t* __restrict__ cellStart,
                                                     const double* __restrict__ hguess,
                                                     const double* __restrict__ epred, int n,
                                                     const GridP* __restrict__ gp, PhysP ph, KnnOut out,
                                                     const uint8_t* __restrict__ gflag, uint32_t* __restrict__ dflags, FbExt fx) {
  const GridP g = *gp;
  const int nfail = *out.failCount;
  if (blockIdx.x == 0 && threadIdx.x == 0) out.failCount[1] += nfail;  // cumulative, read by sphb_counters
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const double BIG = 1.7976931348623157e308;
  const double D = STALE ? fx.rs->D : 0.0;
  for (int f = gwarp; f < nfail; f += nwarps) {
    const int i = out.failList[f];
    const double2 pa = spos[i];
    const uint32_t k = keys[i];
    const int cya = (int)(k / (uint32_t)g.ncx), cxa = (int)(k - (uint32_t)cya * (uint32_t)g.ncx);
    double R = hguess[i] > 0.0 ? 1.5 * hguess[i] : 1.5 * fmax(g.dx, g.dy);
    double td = BIG;
    uint32_t ti = 0xffffffffu;
    int found = 0;
    for (int iter = 0; iter < 64; ++iter) {
      td = BIG; ti = 0xffffffffu; found = 0;
      // unwrapped cell block covering [pa - R, pa + R]; sides that cannot hide anything are "complete"
      int c0, c1, r0, r1;
      double spanx = 0.0, spany = 0.0;  // STALE: half width of the block around the query's cell
      bool fullx = false, fully = false;
      if (STALE) {
        const int wx = (int)fmin(ceil((R + D) * g.inv_dx), 1.0e9), wy = (int)fmin(ceil((R + D) * g.inv_dy), 1.0e9);
        fullx = g.wrapx && 2 * (2 * (long long)wx + 1) > (long long)g.ncx;
        fully = g.wrapy && 2 * (2 * (long long)wy + 1) > (long long)g.ncy;
        c0 = fullx ? -g.ncx : (int)max((long long)cxa - wx, -(long long)g.ncx);
        c1 = fullx ? 2 * g.ncx - 1 : (int)min((long long)cxa + wx, 2 * (long long)g.ncx - 1);
        r0 = fully ? -g.ncy : (int)max((long long)cya - wy, -(long long)g.ncy);
        r1 = fully ? 2 * g.ncy - 1 : (int)min((long long)cya + wy, 2 * (long long)g.ncy - 1);
        spanx = (double)wx * g.dx - D; spany = (double)wy * g.dy - D;
      } else {
        c0 = (int)floor((pa.x - R - g.ox) * g.inv_dx); c1 = (int)floor((pa.x + R - g.ox) * g.inv_dx);
        r0 = (int)floor((pa.y - R - g.oy) * g.inv_dy); r1 = (int)floor((pa.y + R - g.oy) * g.inv_dy);
        c0 = min(c0, cxa); c1 = max(c1, cxa); r0 = min(r0, cya); r1 = max(r1, cya);
      }
      bool doneL, doneR, doneD, doneU;
      if (g.wrapx) { doneL = c0 <= -g.ncx; doneR = c1 >= 2 * g.ncx - 1; c0 = max(c0, -g.ncx); c1 = min(c1, 2 * g.ncx - 1); }
      else { doneL = c0 <= 0; doneR = c1 >= g.ncx - 1; c0 = min(max(c0, 0), g.ncx - 1); c1 = max(min(c1, g.ncx - 1), 0); }
      if (g.wrapy) { doneD = r0 <= -g.ncy; doneU = r1 >= 2 * g.ncy - 1; r0 = max(r0, -g.ncy); r1 = min(r1, 2 * g.ncy - 1); }
      else { doneD = r0 <= 0; doneU = r1 >= g.ncy - 1; r0 = min(max(r0, 0), g.ncy - 1); r1 = max(min(r1, g.ncy - 1), 0); }
      // image order: the reference loops the x image outermost (nearest-neighbour.go:57-61); only the
      // relative order of exactly equal keys depends on it, which parity excludes
      for (int ru = r0; ru <= r1; ++ru) {
        const int iy = g.wrapy ? img_idx(ru, g.ncy) : 0;
        const int row = ru - iy * g.ncy;
        const int ix0 = g.wrapx ? img_idx(c0, g.ncx) : 0, ix1 = g.wrapx ? img_idx(c1, g.ncx) : 0;
        for (int ix = ix0; ix <= ix1; ++ix) {
          const int a = max(c0, ix * g.ncx) - ix * g.ncx, b = min(c1, ix * g.ncx + g.ncx - 1) - ix * g.ncx;
          const int s = (int)cellStart[row * g.ncx + a], e = (int)cellStart[row * g.ncx + b + 1];
          for (int j0 = s; j0 < e; j0 += 32) {
            const int j = j0 + lane;
            bool have = j < e;
            double d2 = BIG;
            int sx = ix, sy = iy;  // image of the candidate: it is taken at pb + (sx Lx, sy Ly)
            if (have) {
              const double2 pb = spos[j];
              if (STALE) {  // nearest image, where the block is narrow enough for that to be the only one in reach
                if (g.wrapx && !fullx) {
                  const double d0 = (pa.x - (double)sx * g.Lx) - pb.x;
                  sx += d0 > 0.5 * g.Lx ? 1 : (d0 < -0.5 * g.Lx ? -1 : 0);
                }
                if (g.wrapy && !fully) {
                  const double d0 = (pa.y - (double)sy * g.Ly) - pb.y;
                  sy += d0 > 0.5 * g.Ly ? 1 : (d0 < -0.5 * g.Ly ? -1 : 0);
                }
                if (sx < -1 || sx > 1 || sy < -1 || sy > 1) have = false;  // farther than a period: not a candidate
              }
              const double qx = (sx == 0) ? pa.x : __dadd_rn(pa.x, -(double)sx * g.Lx);
              const double qy = (sy == 0) ? pa.y : __dadd_rn(pa.y, -(double)sy * g.Ly);
              if (have) d2 = dist_sq(qx - pb.x, qy - pb.y);
            }
            const uint32_t code = img_code(sx, sy) << IMG_SHIFT;
            const double thr = __shfl_sync(0xffffffffu, td, 0);
            // strict admission, self excluded in every image (nearest-neighbour.go:79)
            uint32_t pend = __ballot_sync(0xffffffffu, have && d2 < thr && j != i);
            while (pend) {
              const int src = __ffs(pend) - 1;
              pend &= pend - 1;
              const double v = __shfl_sync(0xffffffffu, d2, src);
              const uint32_t en = __shfl_sync(0xffffffffu, (uint32_t)j | code, src);
              const double t0 = __shfl_sync(0xffffffffu, td, 0);
              if (!(v < t0)) continue;  // the threshold tightened meanwhile (warp-uniform)
              const int p = __popc(__ballot_sync(0xffffffffu, td > v));  // slots 0..p-1 shift towards 0
              const double tdn = __shfl_down_sync(0xffffffffu, td, 1);
              const uint32_t tin = __shfl_down_sync(0xffffffffu, ti, 1);
              if (lane < p - 1) { td = tdn; ti = tin; }
              else if (lane == p - 1) { td = v; ti = en; }
              ++found;
            }
          }
        }
      }
      const bool all = doneL && doneR && doneD && doneU;
      const double t0 = __shfl_sync(0xffffffffu, td, 0);
      if (found >= SPHB_K) {
        // certified if the 32nd distance does not reach past the scanned block on any open side
        const double d = sqrt(t0);
        double reachL, reachR, reachD, reachU;
        if (STALE) {
          reachL = doneL ? 1e300 : spanx; reachR = doneR ? 1e300 : spanx;
          reachD = doneD ? 1e300 : spany; reachU = doneU ? 1e300 : spany;
        } else {
          reachL = doneL ? 1e300 : pa.x - (g.ox + (double)c0 * g.dx);
          reachR = doneR ? 1e300 : (g.ox + (double)(c1 + 1) * g.dx) - pa.x;
          reachD = doneD ? 1e300 : pa.y - (g.oy + (double)r0 * g.dy);
          reachU = doneU ? 1e300 : (g.oy + (double)(r1 + 1) * g.dy) - pa.y;
        }
        const double reach = fmin(fmin(reachL, reachR), fmin(reachD, reachU));
        if (d * (1.0 + 1e-9) <= reach) break;
        R = d * (1.0 + 1e-6);  // guaranteed to suffice next time
      } else {
        if (all) break;  // fewer than 32 (particle, image) candidates exist at all
        R *= 2.0;
      }
    }
    const int tile = i >> 5, ql = i & 31;
    uint32_t* nncol = out.nn + (size_t)tile * 32 * 32 + ql;
    if (found < SPHB_K) {
      if (lane == 0) atomicOr(dflags, DFLAG_UNDERFULL);
      nncol[lane * 32] = 0xffffffffu;
      if (lane == 0) { out.pc[i] = make_double4(0.0, 0.0, 0.0, 0.0); if (fx.dexcl) fx.dexcl[i] = 0.0; }
      continue;
    }
    const double h2 = __shfl_sync(0xffffffffu, td, 0);
    const double h = sqrt(h2);
    const double inv_h = 1.0 / h;
    if (lane == 0 && fx.dexcl) fx.dexcl[i] = 0.0;  // its list has no slots in the tile's block: full search until the next rebuild
    if (!STALE && lane == 0 && g.sides && (((g.sides & 1) && pa.x - g.ox < h) || ((g.sides & 2) && g.ox + (double)g.ncx * g.dx - pa.x < h)))
      atomicOr(dflags, DFLAG_GHOST_THIN);
    double acc = kern_F<KERNEL>(fmin(sqrt(td) * inv_h, 1.0));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    nncol[lane * 32] = ti;
    if (lane == 0 && out.hscale != 0.0) {
      unsigned long long* a = out.hacc + 2 * (blockIdx.x & (HACC_N - 1));
      atomicAdd(a, (unsigned long long)__double2uint_rn(h * out.hscale));
      atomicAdd(a + 1, 1ull);
      if (gflag == nullptr || gflag[i] == GF_OWNED) atomicMax(out.qmax, __float_as_uint(__double2float_ru(h)));
    }
    if (lane == 0) {
      const double rho = ph.Fpref * ph.mass * acc / (h * h);
      const double c = sqrt(ph.cfac * epred[i]);
      out.pc[i] = make_double4(rho, c, h, c * c / (ph.gamma * rho));
    }
  }
}

// image offset of a list entry: the query is shifted by -(img)*L exactly as in K3
__device__ __forceinline__ void decode_entry(uint32_t ent, const GridP& g, int& j, double& offx, double& offy) {
  j = (int)(ent & IDX_MASK);
  const int code = (int)(ent >> IMG_SHIFT);
  const int ix = (code >> 2) - 1, iy = (code & 3) - 1;
  offx = -(double)ix * g.Lx;
  offy = -(double)iy * g.Ly;
}

// Density2D for an arbitrary kernel from the stored list (examples/density, sph.go:306-323)
template <int KERNEL>
__global__ void __launch_bounds__(128) k_density_from_list(const double2* __restrict__ spos,
                                                          const uint32_t* __restrict__ nn, int n, const GridP* __restrict__ gp, PhysP ph,
                                                          double4* __restrict__ pc) {
  const GridP g = *gp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 pa = spos[i];
  double4 q = pc[i];
  const double h = q.z;
  if (!(h > 0.0)) return;
  const double inv_h = 1.0 / h;
  const uint32_t* col = nn + (size_t)(i >> 5) * 1024 + (i & 31);
  double acc = 0.0;
  for (int s = 0; s < SPHB_K; ++s) {
    int j; double ox, oy;
    decode_entry(col[s * 32], g, j, ox, oy);
    const double2 pb = spos[j];
    const double d = sqrt(dist_sq((pa.x + ox) - pb.x, (pa.y + oy) - pb.y));
    acc += kern_F<KERNEL>(fmin(d * inv_h, 1.0));
  }
  q.x = ph.Fpref * ph.mass * acc / (h * h);
  q.w = q.y * q.y / (ph.gamma * q.x);
  pc[i] = q;
}

// -------------------------------------------------------------------------------------------------
// K4: AccelerationAndEDot2D (sph.go:327-401) over the neighbour list, fused with kick, drift-2,
// periodic wrap and reflections (sph.go:122-193) when INTEGRATE.
// -------------------------------------------------------------------------------------------------
struct ForceIO {
  const double2* spos;
  const double2* vpred;
  const double4* pc;
  const uint32_t* nn;
  const uint32_t* keys;       // cell key of every particle (sorted order)
  const uint32_t* cellStart;
  double2* pos;
  double2* vel;
  double* e;
  double2* vdot;
  double* edot;
  const uint8_t* gflag;  // slab mode: only owned particles are evaluated; ghosts are removed afterwards (k_fill_holes)
  uint32_t* qmax;        // [1]: max |Vel|^2 of the evaluated particles after the kick, as float bits rounded up
  // periodic single-handle steps: the cell keys of the NEXT step (drift-1 of the new state in the next grid, which is
  // already known: it only depends on the mean h this evaluation's kNN produced) come out of the epilogue, so
  // the next step needs no k_keys pass.  next_grid == nullptr: off.
  const GridP* next_grid;
  uint32_t *next_keys, *next_rank, *next_count;
  // list reuse (nullptr: off): the epilogue accumulates the displacement of every particle between this evaluation and
  // the next one (max deviation from the reference displacement, fixed-point sum for the mean); stale != 0: the cell
  // table is the last rebuild's, so the staging lookups are shifted back by the mean displacement and widened by D
  ReuseState* rs;
  const ReuseState* stale;
  float2* ucum;     // with rs: cumulative displacement of every particle since the rebuild (local displacement bound)
  int ucum_reset;   // this evaluation is the rebuild: the sum starts here
};

// -------------------------------------------------------------------------------------------------
// K4 (staged): the same computation with the neighbour records staged in shared memory.
//
// A block is FORCE_THREADS consecutive particles of the cell-sorted order (a strip of a grid row), one thread
// per particle.  Their neighbours lie in a few contiguous ranges of the sorted order ("pieces": the strip
// widened by h in the rows above / below, plus periodic-image pieces at the box edges).  The block finds
// the pieces (per-thread cell ranges -> cellStart -> warp + shared-memory min/max per class), copies the
// records {position, predicted velocity, rho, c, h, P} of every piece into shared memory with coalesced
// loads, and the 32 x FORCE_THREADS gathers of the pair loop become shared-memory loads.  A piece is keyed by
// (image code << 28 | sorted index), the same word as a neighbour-list entry, so the lookup of an entry is a
// few unsigned compares against the sorted piece starts.  The staging area is a cache: an entry that is
// not covered (piece table full, staging area full, rows farther than FORCE_RMAX, mixed periodic images) is
// evaluated from global memory after the main loop, so the result never depends on what was staged.
//
// Staged positions are already shifted to the periodic image of their piece: b + img * L, the same first
// operation as the reference's NNPos = b - offset (nearest-neighbour.go:80), so rAB is bit-identical.
// R = double: reference arithmetic (sph.go:327-401).
// R = float : the fp32 build.  Positions are staged as fp32 offsets from a block-local origin (the fp64
//             difference is formed before the conversion); everything per pair is fp32.  Integration stays fp64.
// The viscosity means are staged pre-scaled: rho/2, -0.375 c, h/2 (sph.go:379-386 with alpha = 0.75).
// -------------------------------------------------------------------------------------------------
#define FORCE_THREADS 128
#ifndef FORCE_MINB
#define FORCE_MINB 5
#endif
#define FORCE_NPIECE 6
#define FORCE_RMAX 2
#define FORCE_NCLS ((2 * FORCE_RMAX + 1) * 3)
#define FORCE_CODE_MIXED 15

template <typename R> struct RealV { typedef double2 T; static constexpr int N = 4; };  // vectors of 16 bytes
template <> struct RealV<float> { typedef float4 T; static constexpr int N = 2; };

__device__ __forceinline__ double pair_rsqrt(double x) { return fast_rsqrt(x); }
__device__ __forceinline__ float pair_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double pair_rcp(double x) { return fast_rcp(x); }
__device__ __forceinline__ float pair_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int KERNEL, typename R>
__device__ __forceinline__ R kern_DF_r(R q) {
  if (KERNEL == 1) {
    const R lo = q * (R(3.0) * q - R(2.0));
    const R t = R(1.0) - q;
    return q < R(0.5) ? lo : -t * t;
  }
  const R t = R(1.0) - q;
  return R(-10.0) * q * t * t * t;
}

// one neighbour record as the pair loop consumes it
template <typename R> struct NbrRec { R x, y, vx, vy, rhoh, cs, hh, P; };  // rhoh = rho/2, cs = -0.375 c, hh = h/2

__device__ __forceinline__ NbrRec<double> load_rec(const double2* __restrict__ sm, int nrec, int sl) {
  const double2 a = sm[sl], b = sm[nrec + sl], c = sm[2 * nrec + sl], d = sm[3 * nrec + sl];
  return NbrRec<double>{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
}
__device__ __forceinline__ NbrRec<float> load_rec(const float4* __restrict__ sm, int nrec, int sl) {
  const float4 a = sm[sl], b = sm[nrec + sl];
  return NbrRec<float>{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
}
__device__ __forceinline__ void store_rec(double2* __restrict__ sm, int nrec, int sl, const NbrRec<double>& r) {
  sm[sl] = make_double2(r.x, r.y); sm[nrec + sl] = make_double2(r.vx, r.vy);
  sm[2 * nrec + sl] = make_double2(r.rhoh, r.cs); sm[3 * nrec + sl] = make_double2(r.hh, r.P);
}
__device__ __forceinline__ void store_rec(float4* __restrict__ sm, int nrec, int sl, const NbrRec<float>& r) {
  sm[sl] = make_float4(r.x, r.y, r.vx, r.vy); sm[nrec + sl] = make_float4(r.rhoh, r.cs, r.hh, r.P);
}

// record of particle j as seen through periodic image `code`; the fp32 build takes position and velocity
// relative to a block-local reference `ref` = {x, y, vx, vy} before converting (the pair terms only use
// differences, so the fp32 error scales with the block's spread, not with |x| or a bulk flow velocity)
template <typename R, bool SLAB>
__device__ __forceinline__ NbrRec<R> make_rec(const ForceIO& io, const GridP& g, int j, int code, const double4& ref) {
  const double2 pb = io.spos[j];
  const double2 vb = io.vpred[j];
  const double4 qb = io.pc[j];
  double rho = 0.5 * qb.x;
  if (SLAB && io.gflag[j] == GF_OUTER) rho = -1.0;  // its rho, c, h were not evaluated
  const double x = pb.x + (double)((code >> 2) - 1) * g.Lx, y = pb.y + (double)((code & 3) - 1) * g.Ly;
  NbrRec<R> r;
  r.x = (R)(x - ref.x); r.y = (R)(y - ref.y);  // ref = 0 in the fp64 build: x - 0 is exact
  r.vx = (R)(vb.x - ref.z); r.vy = (R)(vb.y - ref.w);
  r.rhoh = (R)rho; r.cs = (R)(-0.375 * qb.y); r.hh = (R)(0.5 * qb.z); r.P = (R)qb.w;
  return r;
}

// own quantities of the particle a thread evaluates
template <typename R> struct OwnRec { R x, y, xl, yl, vx, vy, rhoh, cs, hh, P, inv_h; };  // (xl, yl): fp32 residual of (x, y)

// one pair of AccelerationAndEDot2D (sph.go:357-397): adds to (ax, ay, aed); sel = 0 discards the pair
// RAW: the neighbour's {rhoh, cs, hh} hold rho, c, h as stored (bulk-staged records): the scaling rides on the additions
template <int KERNEL, typename R, bool RAW = false>
__device__ __forceinline__ void pair_term(const OwnRec<R>& o, const NbrRec<R>& b, bool sel, R& ax, R& ay, R& aed) {
  R rx = b.x - o.x, ry = b.y - o.y;
  if (sizeof(R) == 4) { rx -= o.xl; ry -= o.yl; }  // own position to full precision: only the neighbour's rounding is left
  const R vx = b.vx - o.vx, vy = b.vy - o.vy;
  const R r2 = fma(ry, ry, rx * rx);
  const R dot = fma(vy, ry, vx * rx);
  const R rinv = pair_rsqrt(r2);  // coincident particles give Inf/NaN like the reference (sph.go:391)
  // q = d / h <= 1 up to the rounding of the reciprocals: the kernels' derivatives are O(eps^2) there, no clamp
  const R q = r2 * rinv * o.inv_h;
  R dk = kern_DF_r<KERNEL, R>(q);
  // artificial viscosity, sph.go:375-388: mu = vr hAB / (r^2 + eta^2), Pi = (-alpha cAB mu + beta mu^2) / rhoAB for
  // approaching pairs (vr < 0): min(vr, 0) makes mu, hence Pi, vanish otherwise
  R cs, rs, hs;  // -0.75 cAB, rhoAB, hAB
  if (RAW) { cs = fma(R(-0.375), b.cs, o.cs); rs = fma(R(0.5), b.rhoh, o.rhoh); hs = fma(R(0.5), b.hh, o.hh); }
  else { cs = o.cs + b.cs; rs = o.rhoh + b.rhoh; hs = o.hh + b.hh; }
  const R den = r2 + R(0.01);
  const R dneg = fmin(dot, R(0.0));
  R mu, pi;
  if (sizeof(R) == 4) {
    mu = dneg * hs * pair_rcp(den);
    pi = mu * fma(R(1.5), mu, cs) * pair_rcp(rs);
  } else {  // the two divisions share one reciprocal
    const R inv = pair_rcp(den * rs);
    mu = dneg * hs * (rs * inv);
    pi = mu * fma(R(1.5), mu, cs) * (den * inv);
  }
  R w = (pi + o.P + b.P) * dk * rinv;
  w = sel ? w : R(0.0);
  dk = sel ? dk : R(0.0);
  ax = fma(rx, w, ax);
  ay = fma(ry, w, ay);
  aed = fma(dot, dk, aed);
}

// ---- bulk-copy staging of the fp64 build (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) -------------------------------
// One elected thread per piece copies the piece's rows of spos / vpred / pc straight from global into shared memory; the
// copy engine signals an mbarrier with the byte count.  The records land RAW ({x, y}, {vx, vy}, {rho, c, h, P}); the pair
// loop folds rho/2, -0.375 c, h/2 into its additions (pair_term<RAW>), so a block away from the periodic seam never
// touches the staged bytes: no staging loop, no second barrier.  Seam blocks patch the positions of image pieces in place.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// record layout of the bulk-staged area: nrec x {x, y}, nrec x {vx, vy}, nrec x {rho, c, h, P} (32 bytes each, as stored)
__device__ __forceinline__ NbrRec<double> load_rec_bulk(const double2* __restrict__ sm, int nrec, int sl) {
  const double2 a = sm[sl], b = sm[nrec + sl], c = sm[2 * nrec + 2 * sl], d = sm[2 * nrec + 2 * sl + 1];
  return NbrRec<double>{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
}

struct ForceSt {           // block tables of the staged kernel
  int cls_s[FORCE_NCLS], cls_e[FORCE_NCLS], cls_cmin[FORCE_NCLS], cls_cmax[FORCE_NCLS];
  uint32_t stc[FORCE_NPIECE];           // piece keys (code << 28 | start), ascending; unused = 0xffffffff
  uint2 tab[FORCE_NPIECE + 1];          // [t] = {key - staged offset, key + length} of piece t - 1; [0] = {0, 0}: "not staged"
  int p_s[FORCE_NPIECE], p_len[FORCE_NPIECE], p_off[FORCE_NPIECE], p_code[FORCE_NPIECE];
  int np;
  int fix;  // bulk staging: some piece is a periodic image (its positions need the shift)
};

// the 32 pair interactions of one particle; NP = number of staged pieces the lookup distinguishes
template <int KERNEL, bool SLAB, typename R, int NP, bool BULK>
__device__ __forceinline__ void force_pairs(const ForceIO& io, int n, const GridP& g, const ForceSt& T,
                                            const typename RealV<R>::T* __restrict__ sm, int nrec, int i,
                                            const OwnRec<R>& own, const double4& ref, const uint32_t (&ent0)[8],
                                            R& ax, R& ay, R& aed, bool& thin) {
  uint32_t stc[NP];
#pragma unroll
  for (int m = 0; m < NP; ++m) stc[m] = T.stc[m];
  const uint32_t* col = io.nn + (size_t)(i >> 5) * 1024 + (i & 31);
  uint32_t cur[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) cur[u] = ent0[u];
  bool allhit = true;
#pragma unroll 1
  for (int s0 = 0; s0 < SPHB_K; s0 += 8) {
    uint32_t nxt[8];  // the next eight list entries are in flight while these eight are processed
    if (s0 + 8 < SPHB_K) {
#pragma unroll
      for (int u = 0; u < 8; ++u) nxt[u] = col[(s0 + 8 + u) * 32];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // branch-free: eight independent pairs for the scheduler to interleave
      const uint32_t ent = cur[u];
      int t = NP;  // t = #{m : ent >= stc[m]}: one borrow-propagating subtraction pair per piece
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        uint32_t tmp;
        asm("{sub.cc.u32 %1, %2, %3;\n\tsubc.u32 %0, %0, 0;}" : "+r"(t), "=r"(tmp) : "r"(ent), "r"(stc[m]));
      }
      const uint2 de = T.tab[t];
      const bool hit = ent < de.y;
      const int sl = hit ? (int)(ent - de.x) : 0;
      allhit = allhit && hit;
      NbrRec<R> b;
      if constexpr (BULK) b = load_rec_bulk(sm, nrec, sl);
      else b = load_rec(sm, nrec, sl);
      if (SLAB) thin |= hit && b.rhoh < R(0.0);
      pair_term<KERNEL, R, BULK>(own, b, hit, ax, ay, aed);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
  }
  if (allhit) return;
  // some entries were not staged (rare: box edges, very non-uniform blocks): the same pairs from global memory
#pragma unroll 1
  for (int s = 0; s < SPHB_K; ++s) {
    const uint32_t ent = col[s * 32];
    int t = 0;
#pragma unroll
    for (int m = 0; m < NP; ++m) t += (ent >= stc[m]) ? 1 : 0;
    if (ent < T.tab[t].y) continue;  // was staged
    const int j = (int)(ent & IDX_MASK);
    if ((uint32_t)j >= (uint32_t)n) continue;  // empty slot of an underfull list (reported as SPHB_E_KNN_UNDERFULL)
    const NbrRec<R> b = make_rec<R, SLAB>(io, g, j, (int)(ent >> IMG_SHIFT), ref);
    if (SLAB) thin |= b.rhoh < R(0.0);
    pair_term<KERNEL, R>(own, b, true, ax, ay, aed);
  }
}

template <int KERNEL, bool INTEGRATE, bool SLAB, typename R>
__device__ __forceinline__ void force_block(const ForceIO& io, int n, const GridP* __restrict__ gp, const PhysP& ph,
                                            int nrec, uint32_t* __restrict__ dflags) {
  typedef typename RealV<R>::T RV;
  constexpr bool F32 = sizeof(R) == 4;
  constexpr bool BULK = !F32;  // the fp32 build converts while it stages, so its records cannot be byte copies
  const GridP g = *gp;
  __shared__ ForceSt T;
  __shared__ __align__(8) unsigned long long stage_bar;
  extern __shared__ __align__(16) unsigned char fsm[];
  RV* sm = reinterpret_cast<RV*>(fsm);  // RealV<R>::N arrays of nrec vectors
  const int tid = threadIdx.x, lane = tid & 31;
  const int i0 = blockIdx.x * FORCE_THREADS;
  const int i = i0 + tid;
  const bool active = i < n && (!SLAB || io.gflag[i] == GF_OWNED);
  if (tid < FORCE_NCLS) { T.cls_s[tid] = 0x7fffffff; T.cls_e[tid] = 0; T.cls_cmin[tid] = FORCE_CODE_MIXED; T.cls_cmax[tid] = 0; }
  if (BULK && tid == 0) mbar_init(&stage_bar, 1);
  if (INTEGRATE && active && (tid & 7) == 0) {  // the epilogue's rows: in L2 by the time the pairs are done
    asm volatile("prefetch.global.L2 [%0];" ::"l"(io.pos + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(io.vel + i));
    if ((tid & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.e + i));
  }
  __syncthreads();

  double2 pa = make_double2(0.0, 0.0), va = make_double2(0.0, 0.0);
  double4 qa = make_double4(1.0, 0.0, 0.0, 0.0);
  int cxa = 0, cya = 0, clo = 0, chi = -1, rlo = 0, rhi = -1;
  uint32_t ent0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    pa = io.spos[i]; va = io.vpred[i]; qa = io.pc[i];
    const uint32_t k = io.keys[i];
    const uint32_t* col = io.nn + (size_t)(i >> 5) * 1024 + (i & 31);
#pragma unroll
    for (int u = 0; u < 8; ++u) ent0[u] = col[u * 32];  // in flight during the staging phase
    // rows 8..31 of the warp's list tile (one 128-byte line each) into L2: under the register cap ptxas sinks the pair
    // loop's look-ahead loads to the end of a batch, so their latency is what a warp waits for
    if (lane >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.nn + (size_t)(i >> 5) * 1024 + lane * 32));
    cya = (int)(k / (uint32_t)g.ncx);
    cxa = (int)(k - (uint32_t)cya * (uint32_t)g.ncx);
    double rw = qa.z * (1.0 + 1e-6), lx = pa.x, ly = pa.y;
    if (io.stale) {  // where the neighbours' (stale) cells are: only what gets staged depends on this, never the result
      rw += io.stale->D; lx -= io.stale->ubx; ly -= io.stale->uby;
    }
    clo = (int)floor((lx - rw - g.ox) * g.inv_dx); chi = (int)floor((lx + rw - g.ox) * g.inv_dx);
    rlo = (int)floor((ly - rw - g.oy) * g.inv_dy); rhi = (int)floor((ly + rw - g.oy) * g.inv_dy);
    clo = max(min(clo, cxa), -g.ncx); chi = min(max(chi, cxa), 2 * g.ncx - 1);
    rlo = max(min(rlo, cya), cya - g.ncy + 1); rhi = min(max(rhi, cya), cya + g.ncy - 1);
  }
  // piece classes: (row offset dr, x piece: centre / image -1 / image +1); every class is a contiguous range of
  // the sorted order when taken over consecutive particles (the end of one row abuts the start of the next)
#pragma unroll
  for (int dr = -FORCE_RMAX; dr <= FORCE_RMAX; ++dr) {
    const int ru = cya + dr;
    int iy = 0;
    if (g.wrapy) iy = img_idx(ru, g.ncy);
    const int row = ru - iy * g.ncy;
    const bool has = active && ru >= rlo && ru <= rhi && row >= 0 && row < g.ncy;
    if (!__any_sync(0xffffffffu, has)) continue;
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) {
      int a, b;
      bool okp;
      if (ty == 0) { a = max(clo, 0); b = min(chi, g.ncx - 1); okp = a <= b; }
      else if (ty == 1) { a = max(clo + g.ncx, 0); b = g.ncx - 1; okp = g.wrapx && clo < 0; }
      else { a = 0; b = min(chi - g.ncx, g.ncx - 1); okp = g.wrapx && chi >= g.ncx; }
      const bool hv = has && okp;
      if (!__any_sync(0xffffffffu, hv)) continue;
      int s = 0x7fffffff, e = 0;
      const int code = (int)img_code(ty == 0 ? 0 : (ty == 1 ? -1 : 1), iy);
      if (hv) { s = (int)io.cellStart[row * g.ncx + a]; e = (int)io.cellStart[row * g.ncx + b + 1]; }
      const int smin = __reduce_min_sync(0xffffffffu, s), emax = __reduce_max_sync(0xffffffffu, e);
      const int cmin = __reduce_min_sync(0xffffffffu, hv ? code : FORCE_CODE_MIXED), cmax = __reduce_max_sync(0xffffffffu, hv ? code : 0);
      if (lane == 0) {
        const int c = (dr + FORCE_RMAX) * 3 + ty;
        atomicMin(&T.cls_s[c], smin); atomicMax(&T.cls_e[c], emax);
        atomicMin(&T.cls_cmin[c], cmin); atomicMax(&T.cls_cmax[c], cmax);
      }
    }
  }
  __syncthreads();
  if (tid < 32) {  // piece table by one warp: lane = class in priority order (nearest rows first)
    bool v = false;
    int s = 0, e = 0, code = FORCE_CODE_MIXED;
    if (lane < FORCE_NCLS) {
      const int rr = lane / 3, ty = lane - rr * 3;
      const int dr = (rr == 0) ? 0 : ((rr & 1) ? -((rr + 1) >> 1) : (rr >> 1));  // 0, -1, +1, -2, +2
      const int c = (dr + FORCE_RMAX) * 3 + ty;
      s = T.cls_s[c]; e = T.cls_e[c];
      code = (T.cls_cmin[c] == T.cls_cmax[c]) ? T.cls_cmin[c] : FORCE_CODE_MIXED;
      v = e > s && code != FORCE_CODE_MIXED;
    }
    const uint32_t vm = __ballot_sync(0xffffffffu, v);
    v = v && __popc(vm & ((1u << lane) - 1u)) < FORCE_NPIECE;
    int len = v ? e - s : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int off = incl - len;
    len = min(len, nrec - off);  // clipped to the staging area
    v = v && len > 0;
    const uint32_t key = v ? (((uint32_t)code << IMG_SHIFT) | (uint32_t)s) : 0xffffffffu;
    int rk = 0;
#pragma unroll
    for (int o = 0; o < FORCE_NCLS; ++o) {
      const uint32_t ko = __shfl_sync(0xffffffffu, key, o);
      rk += (ko < key || (ko == key && o < lane)) ? 1 : 0;
    }
    const int np = __popc(__ballot_sync(0xffffffffu, v));
    if (v) {
      T.stc[rk] = key; T.tab[rk + 1] = make_uint2(key - (uint32_t)off, key + (uint32_t)len);
      T.p_s[rk] = s; T.p_len[rk] = len; T.p_off[rk] = off; T.p_code[rk] = code;
    }
    if (lane >= np && lane < FORCE_NPIECE) { T.stc[lane] = 0xffffffffu; T.tab[lane + 1] = make_uint2(0u, 0u); }
    const uint32_t shm = __ballot_sync(0xffffffffu, v && code != 5);
    if (lane == 0) { T.tab[0] = make_uint2(0u, 0u); T.np = np; T.fix = shm != 0u; }
    if constexpr (BULK) {  // the copies are in flight while the block meets at the barrier below
      const uint32_t tot = __reduce_add_sync(0xffffffffu, v ? (uint32_t)len : 0u);
      if (lane == 0) mbar_expect_tx(&stage_bar, tot * 64u);
      __syncwarp();
      if (v) {
        double2* smb = reinterpret_cast<double2*>(fsm);
        bulk_g2s(smb + off, io.spos + s, (uint32_t)len * 16u, &stage_bar);
        bulk_g2s(smb + nrec + off, io.vpred + s, (uint32_t)len * 16u, &stage_bar);
        bulk_g2s(smb + 2 * nrec + 2 * off, io.pc + s, (uint32_t)len * 32u, &stage_bar);
      }
    }
  }
  __syncthreads();
  // block-local origin of the fp32 frame: position and predicted velocity of the middle particle of the block
  double4 ref = make_double4(0.0, 0.0, 0.0, 0.0);
  if (F32) {
    const int mid = min(i0 + FORCE_THREADS / 2, n - 1);
    const double2 pr = io.spos[mid], vr = io.vpred[mid];
    ref = make_double4(pr.x, pr.y, vr.x, vr.y);
  }
  const int np = T.np;
  if constexpr (BULK) {
    mbar_wait(&stage_bar, 0u);  // every thread waits for the bytes itself: no block barrier unless something is patched
    if (T.fix) {  // (slab mode: an outer ghost's row already carries rho = -1, written by the tile search)
      double2* smb = reinterpret_cast<double2*>(fsm);
      for (int m = 0; m < np; ++m) {
        const int len = T.p_len[m], off = T.p_off[m], code = T.p_code[m];
        if (code == 5) continue;
        const double sx = (double)((code >> 2) - 1) * g.Lx, sy = (double)((code & 3) - 1) * g.Ly;
        for (int t = tid; t < len; t += FORCE_THREADS) { double2 pp = smb[off + t]; pp.x += sx; pp.y += sy; smb[off + t] = pp; }
      }
      __syncthreads();
    }
  } else {
    for (int m = 0; m < np; ++m) {  // stage (coalesced)
      const int s = T.p_s[m], len = T.p_len[m], off = T.p_off[m], code = T.p_code[m];
      for (int t = tid; t < len; t += FORCE_THREADS)
        store_rec(sm, nrec, off + t, make_rec<R, SLAB>(io, g, s + t, code, ref));
    }
    __syncthreads();
  }
  if (!active) return;

  OwnRec<R> own;
  own.x = (R)(pa.x - ref.x); own.y = (R)(pa.y - ref.y);
  own.xl = (R)((pa.x - ref.x) - (double)own.x); own.yl = (R)((pa.y - ref.y) - (double)own.y);
  own.vx = (R)(va.x - ref.z); own.vy = (R)(va.y - ref.w);
  own.rhoh = (R)(0.5 * qa.x); own.cs = (R)(-0.375 * qa.y); own.hh = (R)(0.5 * qa.z); own.P = (R)qa.w;
  own.inv_h = pair_rcp((R)qa.z);
  R ax = R(0.0), ay = R(0.0), aed = R(0.0);
  bool thin = false;
  if (np <= 3) force_pairs<KERNEL, SLAB, R, 3, BULK>(io, n, g, T, sm, nrec, i, own, ref, ent0, ax, ay, aed, thin);
  else force_pairs<KERNEL, SLAB, R, FORCE_NPIECE, BULK>(io, n, g, T, sm, nrec, i, own, ref, ent0, ax, ay, aed, thin);
  if (SLAB && thin) atomicOr(dflags, DFLAG_GHOST_THIN);
  const double h = qa.z;
  const double f = ph.mass * ph.DFpref / (h * h * h);
  double2 a = make_double2((double)ax * f + ph.gx, (double)ay * f + ph.gy);
  const double ed = qa.w * (double)aed * ph.mass;  // Benz formulation, sph.go:400
  double2 p = io.pos[i], v = io.vel[i];
  double e = io.e[i];
  if (INTEGRATE) {
    const double dt = 2.0 * ph.dtH;
    // kick (sph.go:122-127), drift 2 (sph.go:130-135): unfused like the reference
    v.x = __dadd_rn(v.x, __dmul_rn(a.x, dt));
    v.y = __dadd_rn(v.y, __dmul_rn(a.y, dt));
    e = __dadd_rn(e, __dmul_rn(__dmul_rn(ed, 2.0), ph.dtH));
    p.x = __dadd_rn(p.x, __dmul_rn(v.x, ph.dtH));
    p.y = __dadd_rn(p.y, __dmul_rn(v.y, ph.dtH));
    // periodic wrap, single shift, X shift skips the Y test (sph.go:147-167)
    if (p.x < ph.hor0) p.x = __dadd_rn(p.x, ph.hor1 - ph.hor0);
    else if (p.x > ph.hor1) p.x = __dsub_rn(p.x, ph.hor1 - ph.hor0);
    else if (p.y < ph.ver0) p.y = __dadd_rn(p.y, ph.ver1 - ph.ver0);
    else if (p.y > ph.ver1) p.y = __dsub_rn(p.y, ph.ver1 - ph.ver0);
    // reflections L, R, U, D: pos -= pos - wall (sph.go:170-193)
    if (p.x < ph.rL) { p.x = __dsub_rn(p.x, __dsub_rn(p.x, ph.rL)); v.x = -v.x; }
    if (p.x > ph.rR) { p.x = __dsub_rn(p.x, __dsub_rn(p.x, ph.rR)); v.x = -v.x; }
    if (p.y < ph.rU) { p.y = __dsub_rn(p.y, __dsub_rn(p.y, ph.rU)); v.y = -v.y; }
    if (p.y > ph.rD) { p.y = __dsub_rn(p.y, __dsub_rn(p.y, ph.rD)); v.y = -v.y; }
  }
  io.vdot[i] = a;
  io.edot[i] = ed;
  if (INTEGRATE) {
    io.pos[i] = p;
    io.vel[i] = v;
    io.e[i] = e;
    if (io.next_grid) {  // == k_keys<true> of the next step (slab mode: of the owned particles)
      const GridP gn = *io.next_grid;
      const double xd = __dadd_rn(p.x, __dmul_rn(v.x, ph.dtH)), yd = __dadd_rn(p.y, __dmul_rn(v.y, ph.dtH));
      const double xs = (gn.wrapx | gn.framex) ? wrap_coord(xd, gn.lox, gn.Lx) : xd, ys = gn.wrapy ? wrap_coord(yd, gn.loy, gn.Ly) : yd;
      const uint32_t k = (uint32_t)cell_of(ys, gn.oy, gn.inv_dy, gn.ncy) * (uint32_t)gn.ncx + (uint32_t)cell_of(xs, gn.ox, gn.inv_dx, gn.ncx);
      io.next_keys[i] = k;
      io.next_rank[i] = atomicAdd(&io.next_count[k], 1u);
    }
    // max speed for the slab driver's migration schedule (whatever subset of the warp is converged here)
    const unsigned m = __activemask();
    const unsigned vm = __reduce_max_sync(m, __float_as_uint(__double2float_ru(v.x * v.x + v.y * v.y)));
    if ((threadIdx.x & 31) == (__ffs(m) - 1)) atomicMax(io.qmax + 1, vm);
    if (io.rs) {
      // displacement from this evaluation's position (pa, wrapped) to the next one's (drift-1 of the coming step), seam
      // jumps removed; deviation from the reference displacement -> max (float bits, rounded up), fixed-point sum
      double ddx = __dadd_rn(p.x, __dmul_rn(v.x, ph.dtH)) - pa.x, ddy = __dadd_rn(p.y, __dmul_rn(v.y, ph.dtH)) - pa.y;
      if (g.Lx > 0.0) ddx -= g.Lx * rint(ddx / g.Lx);
      if (g.Ly > 0.0) ddy -= g.Ly * rint(ddy / g.Ly);
      const double ex_ = ddx - io.rs->mrx, ey_ = ddy - io.rs->mry;
      const float dev = __double2float_ru(sqrt(ex_ * ex_ + ey_ * ey_) * (1.0 + 1e-9));
      const unsigned dm = __reduce_max_sync(m, __float_as_uint(dev));
      const double sc = 1048576.0 * g.inv_dy;  // fixed point: 2^-20 of a cell row (k_reuse_update divides by the same)
      const int qx = (int)fmin(fmax(rint(ddx * sc), -1048576.0), 1048576.0), qy = (int)fmin(fmax(rint(ddy * sc), -1048576.0), 1048576.0);
      const int sx_ = __reduce_add_sync(m, qx), sy_ = __reduce_add_sync(m, qy);
      if (io.ucum) {
        float2 u = io.ucum_reset ? make_float2(0.0f, 0.0f) : io.ucum[i];
        u.x += (float)ddx; u.y += (float)ddy;
        io.ucum[i] = u;
      }
      if ((threadIdx.x & 31) == (__ffs(m) - 1)) {
        atomicMax(&io.rs->Mbits, dm);
        long long* a = io.rs->sum[blockIdx.x & (RS_SLOTS - 1)];
        atomicAdd((unsigned long long*)a, (unsigned long long)(long long)sx_);
        atomicAdd((unsigned long long*)(a + 1), (unsigned long long)(long long)sy_);
      }
    }
  }
}

// the two builds as separate kernels: the fp64 one is capped at 96 registers (5 blocks of 128 threads per SM, the
// number the staging area allows), the fp32 one is left to ptxas (64 registers, 8 blocks)
template <int KERNEL, bool INTEGRATE, bool SLAB>
__global__ void __launch_bounds__(FORCE_THREADS, FORCE_MINB) k_force_st(ForceIO io, int n, const GridP* __restrict__ gp, PhysP ph, int nrec,
                                                              uint32_t* __restrict__ dflags) {
  force_block<KERNEL, INTEGRATE, SLAB, double>(io, n, gp, ph, nrec, dflags);
}
template <int KERNEL, bool INTEGRATE, bool SLAB>
__global__ void __launch_bounds__(FORCE_THREADS) k_force_st32(ForceIO io, int n, const GridP* __restrict__ gp, PhysP ph, int nrec,
                                                             uint32_t* __restrict__ dflags) {
  force_block<KERNEL, INTEGRATE, SLAB, float>(io, n, gp, ph, nrec, dflags);
}

// -------------------------------------------------------------------------------------------------
// slab mode helpers
// -------------------------------------------------------------------------------------------------
struct SlabP {
  double x_lo, x_hi, ghost_w, inner_w;
  double Lx, frame_lo;  // periodic x: period and low end of the image frame centred on the slab; Lx = 0: none
  int has_left, has_right;
};

__device__ __forceinline__ double slab_frame_x(double x, const SlabP& sl) {
  return sl.Lx > 0.0 ? wrap_coord(x, sl.frame_lo, sl.Lx) : x;
}

// append with one atomic per warp; returns the slot of this lane or -1
__device__ __forceinline__ int warp_append_slot(bool want, int* counter) {
  const uint32_t m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return -1;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return want ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// Ghost records carry the predictor's inputs, {x, y, vx, vy, vdotx, vdoty, e, edot, id, h}: the receiving slab
// drifts and predicts its ghosts in the same reorder kernel and with the same arithmetic as its owned
// particles (sph.go:108-117), so no separate predict pass over the state is needed and a ghost is bit-identical
// to the owner's copy.  mode 0 (CalculateForces on the state as is) sends {VPred, EPred} in the velocity / energy
// slots with zero derivatives.  The width test uses the position the evaluation will see (after drift-1).
#define HALO_REC 10
__global__ void __launch_bounds__(256) k_pack_halo(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                                  const double2* __restrict__ vdot, const double2* __restrict__ vpred,
                                                  const double* __restrict__ e, const double* __restrict__ edot,
                                                  const double* __restrict__ epred, const int64_t* __restrict__ id,
                                                  const double4* __restrict__ pc, int n, SlabP sl, int mode, double dtH,
                                                  double* __restrict__ buf_lo, double* __restrict__ buf_hi, int cap,
                                                  int* __restrict__ counters, uint32_t* __restrict__ dflags,
                                                  int* __restrict__ idx_lo, int* __restrict__ idx_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool want_lo = false, want_hi = false;
  double2 p = make_double2(0, 0), v = make_double2(0, 0);
  if (i < n) {
    p = pos[i];
    v = vel[i];
    const double xe = mode == 2 ? __dadd_rn(p.x, __dmul_rn(v.x, dtH)) : p.x;
    const double xl = slab_frame_x(xe, sl);
    want_lo = sl.has_left && (xl - sl.x_lo < sl.ghost_w);
    want_hi = sl.has_right && (sl.x_hi - xl <= sl.ghost_w);
  }
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int slot = warp_append_slot(side == 0 ? want_lo : want_hi, counters + side);
    if (slot < 0) continue;
    if (slot >= cap) { atomicOr(dflags, DFLAG_BUF_FULL); continue; }
    double* r = (side == 0 ? buf_lo : buf_hi) + (size_t)slot * HALO_REC;
    if (idx_lo) (side == 0 ? idx_lo : idx_hi)[slot] = i;  // ring: the same particles are sent again by every reuse evaluation
    r[0] = p.x; r[1] = p.y;
    if (mode == 0) {
      const double2 vp = vpred[i];
      r[2] = vp.x; r[3] = vp.y; r[4] = 0.0; r[5] = 0.0; r[6] = epred[i]; r[7] = 0.0;
    } else {
      const double2 a = vdot[i];
      r[2] = v.x; r[3] = v.y; r[4] = a.x; r[5] = a.y; r[6] = e[i]; r[7] = edot[i];
    }
    r[8] = __longlong_as_double(id[i]);
    r[9] = pc[i].z;
  }
}

__global__ void __launch_bounds__(256) k_add_ghosts(const double* __restrict__ buf, int count, SlabP sl, int mode, double dtH,
                                                   double2* __restrict__ pos, double2* __restrict__ vel,
                                                   double2* __restrict__ vdot, double2* __restrict__ vpred,
                                                   double* __restrict__ e, double* __restrict__ edot,
                                                   double* __restrict__ epred, int64_t* __restrict__ id,
                                                   double4* __restrict__ pc, uint8_t* __restrict__ gflag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* r = buf + (size_t)k * HALO_REC;
  const double x = r[0];
  pos[k] = make_double2(x, r[1]);
  vel[k] = make_double2(r[2], r[3]);
  vdot[k] = make_double2(r[4], r[5]);
  vpred[k] = make_double2(r[2], r[3]);  // read by the mode-0 reorder only
  e[k] = r[6]; edot[k] = r[7];
  epred[k] = r[6];
  id[k] = __double_as_longlong(r[8]);
  pc[k] = make_double4(0.0, 0.0, r[9], 0.0);
  const double xe = mode == 2 ? __dadd_rn(x, __dmul_rn(r[2], dtH)) : x;
  const double xl = slab_frame_x(xe, sl);
  gflag[k] = (xl >= sl.x_lo - sl.inner_w && xl < sl.x_hi + sl.inner_w) ? GF_INNER : GF_OUTER;
}

// migration records {x, y, vx, vy, e, vdotx, vdoty, edot, h, id, rho, c}
__global__ void __launch_bounds__(256) k_pack_migrants(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                                      const double* __restrict__ e, const double2* __restrict__ vdot,
                                                      const double* __restrict__ edot, const int64_t* __restrict__ id,
                                                      const double4* __restrict__ pc, uint8_t* __restrict__ gflag,
                                                      int n, SlabP sl, int side, double* __restrict__ buf, int cap,
                                                      int* __restrict__ counter, uint32_t* __restrict__ dflags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool want = false;
  double2 p = make_double2(0, 0);
  if (i < n && gflag[i] == GF_OWNED) {
    p = pos[i];
    const double xl = slab_frame_x(p.x, sl);
    want = side == 0 ? (sl.has_left && xl < sl.x_lo) : (sl.has_right && xl >= sl.x_hi);
  }
  const int slot = warp_append_slot(want, counter);
  if (slot < 0) return;
  if (slot >= cap) { atomicOr(dflags, DFLAG_BUF_FULL); return; }
  gflag[i] = GF_LEAVING;
  double* r = buf + (size_t)slot * 12;
  const double2 v = vel[i], a = vdot[i];
  const double4 q = pc[i];
  r[0] = p.x; r[1] = p.y; r[2] = v.x; r[3] = v.y; r[4] = e[i]; r[5] = a.x; r[6] = a.y; r[7] = edot[i];
  r[8] = q.z; r[9] = __longlong_as_double(id[i]); r[10] = q.x; r[11] = q.y;
}

__global__ void __launch_bounds__(256) k_add_migrants(const double* __restrict__ buf, int count, double gamma,
                                                     double2* __restrict__ pos, double2* __restrict__ vel,
                                                     double2* __restrict__ vdot, double2* __restrict__ vpred,
                                                     double* __restrict__ e, double* __restrict__ edot,
                                                     double* __restrict__ epred, int64_t* __restrict__ id,
                                                     double4* __restrict__ pc, uint8_t* __restrict__ gflag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* r = buf + (size_t)k * 12;
  pos[k] = make_double2(r[0], r[1]);
  vel[k] = make_double2(r[2], r[3]);
  e[k] = r[4];
  vdot[k] = make_double2(r[5], r[6]);
  edot[k] = r[7];
  vpred[k] = make_double2(r[2], r[3]);
  epred[k] = r[4];
  id[k] = __double_as_longlong(r[9]);
  const double rho = r[10], c = r[11];
  pc[k] = make_double4(rho, c, r[8], rho > 0.0 ? c * c / (gamma * rho) : 0.0);
  gflag[k] = GF_OWNED;
}

// In-place compaction (ghost removal after an evaluation, departed particles after a migration): entries of
// [0, nslots) flagged GF_OWNED are kept and must end up in [0, nkeep), nkeep = their number.  The order of the
// particles is irrelevant (the next evaluation sorts them), so it is enough to move the owned entries that sit at
// or beyond nkeep ("fillers") into the non-owned slots below nkeep ("holes"): there are equally many of both,
// a fraction of a percent of the particles, instead of a pass over the whole state.
__global__ void __launch_bounds__(256) k_find_holes(const uint8_t* __restrict__ gflag, int nslots, int nkeep,
                                                   int* __restrict__ holes, uint32_t* __restrict__ fillers,
                                                   int* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool owned = i < nslots && gflag[i] == GF_OWNED;
  const bool hole = i < nkeep && !owned, filler = i >= nkeep && owned;
  const int sh = warp_append_slot(hole, counters), sf = warp_append_slot(filler, counters + 1);
  if (sh >= 0) holes[sh] = i;
  if (sf >= 0) fillers[sf] = (uint32_t)i;
}

struct SoaPtr {
  double2 *pos, *vel, *vdot, *vpred;
  double *e, *edot, *epred;
  int64_t* id;
  double4* pc;
  uint8_t* ghost;
};

__global__ void __launch_bounds__(256) k_fill_holes(SoaPtr a, const int* __restrict__ holes, const uint32_t* __restrict__ fillers,
                                                   const int* __restrict__ counters, uint32_t* __restrict__ dflags,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ rank) {
  const int nh = counters[0];
  if (blockIdx.x == 0 && threadIdx.x == 0 && counters[1] != nh) atomicOr(dflags, DFLAG_BUF_FULL);  // cannot happen
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nh; k += gridDim.x * blockDim.x) {
    const int d = holes[k];
    const uint32_t f = fillers[k];
    a.pos[d] = a.pos[f]; a.vel[d] = a.vel[f]; a.vdot[d] = a.vdot[f]; a.vpred[d] = a.vpred[f];
    a.e[d] = a.e[f]; a.edot[d] = a.edot[f]; a.epred[d] = a.epred[f];
    a.id[d] = a.id[f]; a.pc[d] = a.pc[f]; a.ghost[d] = GF_OWNED;
    if (keys) { keys[d] = keys[f]; rank[d] = rank[f]; }  // the next step's cell key and arrival rank (force epilogue)
  }
}

// -------------------------------------------------------------------------------------------------
// reductions and statistics
// -------------------------------------------------------------------------------------------------
// per-block partials of {min x, max x, min y, max y, sum h, max h, sum e, sum rho, #(h > 0), max |v|^2}; one block folds them
#define STAT_N 10
#define STAT_BLOCKS 592
__device__ __forceinline__ double stat_comb(int k, double a, double b) {
  return (k == 0 || k == 2) ? fmin(a, b) : ((k == 1 || k == 3 || k == 5 || k == 9) ? fmax(a, b) : a + b);
}
__device__ __forceinline__ double stat_init(int k) {
  return (k == 0 || k == 2) ? 1.7976931348623157e308 : ((k == 1 || k == 3 || k == 5 || k == 9) ? -1.7976931348623157e308 : 0.0);
}

__global__ void __launch_bounds__(256) k_stats_partial(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                                      const double4* __restrict__ pc, const double* __restrict__ e, int n,
                                                      double* __restrict__ part, const uint8_t* __restrict__ gflag) {
  __shared__ double sh[8][STAT_N];
  double v[STAT_N];
#pragma unroll
  for (int k = 0; k < STAT_N; ++k) v[k] = stat_init(k);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (gflag && gflag[i] != GF_OWNED) continue;  // ghosts interleaved with the owned particles (ring, inside a cycle)
    double2 p = pos[i];
    double4 q = pc[i];
    const double2 u = vel[i];
    v[9] = fmax(v[9], u.x * u.x + u.y * u.y);
    v[0] = fmin(v[0], p.x); v[1] = fmax(v[1], p.x);
    v[2] = fmin(v[2], p.y); v[3] = fmax(v[3], p.y);
    v[4] += q.z; v[5] = fmax(v[5], q.z);
    v[6] += e[i]; v[7] += q.x;
    v[8] += (q.z > 0.0) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < STAT_N; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = stat_comb(k, v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < STAT_N; ++k) sh[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < STAT_N) {
    double r = sh[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) r = stat_comb(threadIdx.x, r, sh[w][threadIdx.x]);
    part[blockIdx.x * STAT_N + threadIdx.x] = r;
  }
}
__global__ void __launch_bounds__(32 * STAT_N) k_stats_final(const double* __restrict__ part, int nblk,
                                                            double* __restrict__ outv) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;  // one warp per statistic
  double r = stat_init(k);
  for (int b = lane; b < nblk; b += 32) r = stat_comb(k, r, part[b * STAT_N + k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = stat_comb(k, r, __shfl_xor_sync(0xffffffffu, r, o));
  if (lane == 0) outv[k] = r;
}

// -------------------------------------------------------------------------------------------------
// grid set-up on the device (one thread): the cell edge follows the mean smoothing length of the previous
// evaluation, or a mean-density estimate when there is none.  Keeping GridP in device memory lets a
// sequence of steps be enqueued without a host round trip (open axes: the box follows the particles).
// hor/ver are the search periodicity (nearest-neighbour.go:28): {-DBL_MAX, DBL_MAX} = open.
// -------------------------------------------------------------------------------------------------
struct GridTune {
  double cell_per_h;   // cell height dy = cell_per_h * mean h
  double aspect;       // dx = aspect * dy
  double ppc0;         // particles per cell for the first (h unknown) evaluation
  int ncell_max;       // capacity of the cell table
  int force_nc;        // > 0: fixed cells per axis (tests)
};

__global__ void k_make_grid(const double* __restrict__ stats, int n, double hor0, double hor1, double ver0, double ver1,
                            SlabP sl, int slab_on, GridTune t, GridP* __restrict__ out, unsigned long long* __restrict__ hacc,
                            double hscale, int use_hacc) {
  // the smoothing-length accumulator of the previous evaluation (one warp sums the slots and clears them)
  unsigned long long hs = 0, hc = 0;
  for (int k = threadIdx.x; k < HACC_N; k += 32) { hs += hacc[2 * k]; hc += hacc[2 * k + 1]; hacc[2 * k] = 0; hacc[2 * k + 1] = 0; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { hs += __shfl_xor_sync(0xffffffffu, hs, o); hc += __shfl_xor_sync(0xffffffffu, hc, o); }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  GridP g;
  g.wrapx = !(hor0 == -1.7976931348623157e308);
  g.wrapy = !(ver0 == -1.7976931348623157e308);
  double ex, ey;
  g.framex = 0; g.sides = 0;
  if (slab_on) {  // the slab's own frame: ghosts extend the box on the sides that have a neighbour
    const int periodic = g.wrapx;
    g.wrapx = 0;
    g.framex = periodic && sl.Lx > 0.0;
    g.Lx = g.framex ? sl.Lx : 0.0;
    g.lox = g.framex ? sl.frame_lo : 0.0;
    const double lo = sl.has_left ? sl.x_lo - sl.ghost_w : stats[0];
    const double hi = sl.has_right ? sl.x_hi + sl.ghost_w : stats[1];
    g.ox = lo; ex = hi - lo;
    g.sides = (sl.has_left ? 1 : 0) | (sl.has_right ? 2 : 0);
  }
  else if (g.wrapx) { g.lox = hor0; g.Lx = hor1 - hor0; g.ox = hor0; ex = g.Lx; }
  else { g.lox = 0; g.Lx = 0; g.ox = stats[0]; ex = stats[1] - stats[0]; }
  if (g.wrapy) { g.loy = ver0; g.Ly = ver1 - ver0; g.oy = ver0; ey = g.Ly; }
  else { g.loy = 0; g.Ly = 0; g.oy = stats[2]; ey = stats[3] - stats[2]; }
  const double scale = fmax(fmax(ex, ey), 1e-300);
  if (!(ex > 1e-12 * scale)) ex = 1e-12 * scale;
  if (!(ey > 1e-12 * scale)) ey = 1e-12 * scale;
  double d;  // row height dy; dx = aspect * dy
  const double asp = (t.aspect > 0.0) ? t.aspect : 1.0;
  const double nh = use_hacc ? (double)hc : stats[8];
  const double sumh = use_hacc ? (double)hs / hscale : stats[4];
  if (nh > 0.0) d = t.cell_per_h * sumh / nh;
  else d = sqrt(t.ppc0 * ex * ey / ((double)(n > 0 ? n : 1) * asp));
  if (!(d > 0.0)) d = scale;
  double fx = fmin(fmax(floor(ex / (d * asp)), 1.0), 1.0e6), fy = fmin(fmax(floor(ey / d), 1.0), 1.0e6);
  for (int it = 0; it < 8 && fx * fy > (double)t.ncell_max; ++it) {
    d *= sqrt(fx * fy / (double)t.ncell_max) * 1.0001;
    fx = fmax(floor(ex / (d * asp)), 1.0);
    fy = fmax(floor(ey / d), 1.0);
  }
  if (fx * fy > (double)t.ncell_max) { fx = 1.0; fy = 1.0; }
  if (t.force_nc > 0) { fx = fy = (double)t.force_nc; }
  g.ncx = (int)fx; g.ncy = (int)fy;
  g.dx = ex / fx; g.dy = ey / fy;
  g.inv_dx = 1.0 / g.dx; g.inv_dy = 1.0 / g.dy;
  *out = g;
}

// expand the neighbour list for download in the reference's layout: [particle][slot], slots sorted by
// DESCENDING distance so that slot 0 is the farthest neighbour = h (nearest-neighbour.go:139-153).
// NN_IDX = index in current device order (-1 = none), NN_DIST = NNDists, NN_POS = NNPos (neighbour image
// position in the frame of the query's raw position, nearest-neighbour.go:80).
__global__ void __launch_bounds__(128) k_expand_list(const double2* __restrict__ spos, const double2* __restrict__ pos,
                                                    const uint32_t* __restrict__ nn, int off, int m_count,
                                                    const GridP* __restrict__ gp, int32_t* __restrict__ idx_out,
                                                    double* __restrict__ dist_out, double2* __restrict__ pos_out) {
  const int t_ = blockIdx.x * blockDim.x + threadIdx.x;
  if (t_ >= m_count) return;
  const int i = off + t_;
  const GridP g = *gp;
  const double2 pa = spos[i];
  const double2 praw = pos[i];
  const uint32_t* col = nn + (size_t)(i >> 5) * 1024 + (i & 31);
  double dd[SPHB_K];
  uint32_t ee[SPHB_K];
  int m = 0;
  for (int s = 0; s < SPHB_K; ++s) {
    const uint32_t ent = col[s * 32];
    if (ent == 0xffffffffu) continue;
    int j; double ox, oy;
    decode_entry(ent, g, j, ox, oy);
    const double2 pb = spos[j];
    const double d2 = dist_sq((pa.x + ox) - pb.x, (pa.y + oy) - pb.y);
    int t = m;  // insertion sort, descending; equal keys keep scan order
    for (; t > 0 && dd[t - 1] < d2; --t) { dd[t] = dd[t - 1]; ee[t] = ee[t - 1]; }
    dd[t] = d2; ee[t] = ent;
    ++m;
  }
  for (int s = 0; s < SPHB_K; ++s) {
    const size_t o = (size_t)t_ * SPHB_K + s;
    if (s >= m) {
      if (idx_out) idx_out[o] = -1;
      if (dist_out) dist_out[o] = 0.0;
      if (pos_out) pos_out[o] = make_double2(0.0, 0.0);
      continue;
    }
    int j; double ox, oy;
    decode_entry(ee[s], g, j, ox, oy);
    const double2 pb = spos[j];
    if (idx_out) idx_out[o] = j;
    if (dist_out) dist_out[o] = sqrt(dd[s]);
    if (pos_out) pos_out[o] = make_double2(praw.x + ((pb.x - ox) - pa.x), praw.y + ((pb.y - oy) - pa.y));
  }
}

__global__ void __launch_bounds__(256) k_split_pc(const double4* __restrict__ pc, int n, double* __restrict__ rho,
                                                 double* __restrict__ c, double* __restrict__ h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 q = pc[i];
  if (rho) rho[i] = q.x;
  if (c) c[i] = q.y;
  if (h) h[i] = q.z;
}

// dense-id check: every id in [0, n) exactly once (seen[] zeroed by the caller)
__global__ void __launch_bounds__(256) k_check_dense(const int64_t* __restrict__ id, int n, uint32_t* __restrict__ seen,
                                                    int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = id[i];
  if (v < 0 || v >= n || atomicExch(&seen[v], 1u) != 0u) *bad = 1;
}

// host-order (by id) staging arrays -> device order: field[i] = staged[id[i]]
__global__ void __launch_bounds__(256) k_gather_by_id(const int64_t* __restrict__ id, int n, const double2* __restrict__ spos_,
                                                     const double2* __restrict__ svel, const double* __restrict__ se,
                                                     double2* __restrict__ pos, double2* __restrict__ vel, double* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = id[i];
  if (spos_) pos[i] = spos_[k];
  if (svel) vel[i] = svel[k];
  if (se) e[i] = se[k];
}

// frame data of animator.go:75-101: pixel coordinates (float32 products like the reference) and the colour-ramp index;
// id != nullptr: element id[i] of the outputs (the caller's own particle order), else element i
__global__ void __launch_bounds__(256) k_frame(const double2* __restrict__ pos, const double4* __restrict__ pc,
                                              const int64_t* __restrict__ id, int n,
                                              float w, float h, double colour_div, float2* __restrict__ xy,
                                              uint8_t* __restrict__ colour) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = pos[i];
  const double rho = pc[i].x;
  if (id) i = (int)id[i];
  xy[i] = make_float2(__fmul_rn((float)p.x, w), __fmul_rn((float)p.y, h));
  const double cf = fmin(__dmul_rn(__ddiv_rn(rho, colour_div), 256.0), 255.0);  // Rho / (m N 10) * 256, math.Min(., 255)
  colour[i] = (uint8_t)(int)cf;                           // uint8(): truncation; negative / NaN do not occur (Rho >= 0)
}

// upload helpers: scatter host-provided rho into pc.x, invalidate h when positions are overwritten
__global__ void __launch_bounds__(256) k_set_pc(double4* __restrict__ pc, int n, const double* __restrict__ rho,
                                               int zero_h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 q = pc[i];
  if (rho) q.x = rho[i];
  if (zero_h) q.z = 0.0;
  pc[i] = q;
}

__global__ void __launch_bounds__(256) k_iota64(int64_t* __restrict__ id, int n, int64_t base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id[i] = base + i;
}

#include "sphb_reuse.cuh"
#include "sphb_ring.cuh"
