// sphb_reuse.cuh — kernels of a REUSE evaluation (certified reuse of the neighbour lists, see ReuseState in
// sphb_kernels.cuh).  Pipeline of such a step, in place of keys / sort / reorder / tile search:
//   k_predict      drift-1 + predict, streaming, particle order unchanged            (sph.go:108-117)
//   k_knn_reuse    exact kNN(32) from the <= 48 stored candidates (slots of the tile's staged block) + density + sound
//                  speed, certificate per particle;  k_knn_annulus (rebuild evaluations) collects the further candidates
//   k_knn_fallback<STALE> for the particles the certificate refused
//   k_force_st*    as in every step (staging lookups on the stale cell table)
//   k_reuse_update D += 2 max|delta - mref|, mean displacement, feedback record for the host
#pragma once

// drift-1 + predict in place (the reorder kernel does this while gathering; here nothing moves)
__global__ void __launch_bounds__(256) k_predict(double2* __restrict__ pos, const double2* __restrict__ vel,
                                                const double2* __restrict__ vdot, const double* __restrict__ e,
                                                const double* __restrict__ edot, double2* __restrict__ vpred,
                                                double* __restrict__ epred, double2* __restrict__ spos, int n,
                                                const GridP* __restrict__ gp, double dtH, const uint8_t* __restrict__ gflag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gflag && gflag[i] != GF_OWNED) return;  // ghosts arrive predicted (slab mode)
  const GridP g = *gp;
  double2 p = pos[i];
  const double2 v = vel[i], a = vdot[i];
  p.x = __dadd_rn(p.x, __dmul_rn(v.x, dtH));
  p.y = __dadd_rn(p.y, __dmul_rn(v.y, dtH));
  pos[i] = p;
  vpred[i] = make_double2(__dadd_rn(v.x, __dmul_rn(a.x, dtH)), __dadd_rn(v.y, __dmul_rn(a.y, dtH)));
  epred[i] = __dadd_rn(e[i], __dmul_rn(edot[i], dtH));
  double2 sp;
  sp.x = (g.wrapx | g.framex) ? wrap_coord(p.x, g.lox, g.Lx) : p.x;
  sp.y = g.wrapy ? wrap_coord(p.y, g.loy, g.Ly) : p.y;
  spos[i] = sp;
}

// -------------------------------------------------------------------------------------------------
// Shared by the two kernels below: the tile's staged block, rebuilt from the saved piece table.  Lane p holds piece p
// (first index, length, image code); every piece is padded to a multiple of 8 slots like in the tile search, so a slot
// number means the same particle in all three kernels for the whole cycle (the sorted order does not change).
// -------------------------------------------------------------------------------------------------
struct TilePieces { int s, len, off; uint32_t code; int nst; };

__device__ __forceinline__ TilePieces tile_pieces(const KnnExt& ex, int tile, int npc, int lane) {
  TilePieces p{0, 0, 0, 5u, 0};
  if (lane < npc) {
    const uint2 e = ex.ptab[(size_t)tile * 32 + lane];
    p.s = (int)(e.x & IDX_MASK); p.code = e.x >> IMG_SHIFT; p.len = (int)e.y;
  }
  int incl = (p.len + 7) & ~7;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  p.off = incl - ((p.len + 7) & ~7);
  p.nst = __shfl_sync(0xffffffffu, incl, 31);
  return p;
}

// -------------------------------------------------------------------------------------------------
// Rebuild evaluations, after the tile search: the further candidates of every particle - those between its new h and
// rgx = h_prev (1 + skin) - from the tile's shared block, as slots; the up to SPHB_KX nearest are kept, and the exclusion
// radius dexcl: every particle that is neither a neighbour nor kept is at least that far away.
//   fp32 keys in the tile search's own frame, computed by the same instructions from the same staged values, so they
//   are bit-identical to the keys that search ranked: it leaves, per particle, the smallest key it did NOT take (or its
//   acceptance threshold) in dexcl, and the candidates here are exactly the staged particles with a key in
//   [that key, rgx^2 (1 + delta)).  Candidates the lane saw and did not keep have a key >= T, what it did not see lies
//   beyond rgx:  dexcl^2 = min(T (1 - 2 delta), rgx^2 (1 - 1e-5)).
// Column entries are 32-bit: key bits 31..9 (truncated: only ever lowers a bound) | slot (9 bits).
// -------------------------------------------------------------------------------------------------
#define ANN_CAP 48
__host__ __device__ inline size_t annulus_smem_bytes_per_warp(int ncw) { return (size_t)ANN_CAP * 128 + (size_t)ncw * 12; }

__global__ void __launch_bounds__(KNN_THREADS) k_knn_annulus(const double2* __restrict__ spos, const uint32_t* __restrict__ keys,
                                                            const uint32_t* __restrict__ cellStart, const double* __restrict__ hguess,
                                                            const double4* __restrict__ pc, int n, const GridP* __restrict__ gp,
                                                            KnnTune tune, KnnExt ex, const uint8_t* __restrict__ gflag,
                                                            uint32_t* __restrict__ dflags) {
  const GridP g = *gp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int NCW = tune.ncw;
  unsigned char* wb = smem_raw + (size_t)warp * annulus_smem_bytes_per_warp(NCW);
  uint32_t* col = reinterpret_cast<uint32_t*>(wb) + lane;                              // [slot * 32]
  float2* candF = reinterpret_cast<float2*>(wb + (size_t)ANN_CAP * 128);                // fp32 tile-relative positions, pairs
  const float4* candF4 = reinterpret_cast<const float4*>(candF);
  uint32_t* candE = reinterpret_cast<uint32_t*>(wb + (size_t)ANN_CAP * 128 + (size_t)NCW * 8);     // index | image code << 28
  const int tile = blockIdx.x * KNN_WARPS + warp;
  if (tile * 32 >= n) return;
  const int i = tile * 32 + lane;
  const bool valid = i < n && (gflag == nullptr || gflag[i] != GF_OUTER);
  const TileInfo ti = ex.tinfo[tile];
  if (ti.npc == 0) { if (valid) ex.dexcl[i] = 0.0; return; }  // no shared block: its particles take the full search
  double xa = 0, ya = 0, h = 0, rgx = 0, tkey = 0;
  int cxa = 0, cya = 0;
  if (valid) {
    const double2 p = spos[i];
    xa = p.x; ya = p.y;
    const uint32_t k = keys[i];
    cya = (int)(k / (uint32_t)g.ncx);
    cxa = (int)(k - (uint32_t)cya * (uint32_t)g.ncx);
    h = pc[i].z;
    rgx = knn_ext_radius(hguess[i] * (1.0 + tune.guess_margin), tune.guess_margin, ex.skin);  // the radius the block was sized for
    tkey = ex.dexcl[i];  // the tile search's first key beyond the 32 neighbours (0: it refused the lane - no extended list)
  }
  bool ok = valid && h > 0.0 && rgx > h && tkey > 0.0;
  int clo, chi, rlo, rhi;
  knn_cell_range(g, xa, ya, rgx * (1.0 + 1e-4), cxa, cya, clo, chi, rlo, rhi);
  const TilePieces P = tile_pieces(ex, tile, ti.npc, lane);
  const int npc = ti.npc, c0 = ti.c0, c1 = ti.c1, r0 = ti.r0, r1 = ti.r1;
  const int ix0 = g.wrapx ? img_idx(c0, g.ncx) : 0, ix1 = g.wrapx ? img_idx(c1, g.ncx) : 0, nix = ix1 - ix0 + 1;
  // fp32 frame of the tile (as in the tile search)
  const double xref = g.ox + 0.5 * (double)(c0 + c1 + 1) * g.dx, yref = g.oy + 0.5 * (double)(r0 + r1 + 1) * g.dy;
  double V = fmax(0.5 * (double)(c1 - c0 + 1) * g.dx, 0.5 * (double)(r1 - r0 + 1) * g.dy);
  V = fmax(V, warp_max_d(valid ? fmax(fabs(xa - xref), fabs(ya - yref)) + rgx : 0.0));
  const float qfx = (float)(xa - xref), qfy = (float)(ya - yref);
  const float2 nqx2 = make_float2(-qfx, -qfx), nqy2 = make_float2(-qfy, -qfy);
  const double delta = 3.0 * (2.384185791015625e-07 * V / fmax(h, 1e-300) + 4.76837158203125e-07);
  // (below this bound a particle outside the tile search's own windows cannot have a key under its threshold)
  if (ok && !(delta < 5e-5)) ok = false;
  const float Vf = (float)V * 1.000001f;
  // stage: fp32 positions and list entries
  __syncwarp();
  for (int pc_ = 0; pc_ < npc; ++pc_) {
    const int s = __shfl_sync(0xffffffffu, P.s, pc_), len = __shfl_sync(0xffffffffu, P.len, pc_);
    const int off = __shfl_sync(0xffffffffu, P.off, pc_);
    const uint32_t code = __shfl_sync(0xffffffffu, P.code, pc_);
    const int ix = (int)(code >> 2) - 1, iy = (int)(code & 3u) - 1;
    const double sx = (double)ix * g.Lx - xref, sy = (double)iy * g.Ly - yref;
    const int len8 = (len + 7) & ~7;
    for (int t = lane; t < len8; t += 32) {
      float fx = 3.0e18f, fy = 3.0e18f;
      if (t < len) {
        const double2 pb = spos[s + t];
        fx = (float)(pb.x + sx); fy = (float)(pb.y + sy);
        if (!(fabsf(fx) <= Vf && fabsf(fy) <= Vf)) { fx = 3.0e18f; fy = 3.0e18f; }
        candE[off + t] = (uint32_t)(s + t) | (code << IMG_SHIFT);
      }
      float* cf = reinterpret_cast<float*>(candF) + (size_t)((off + t) >> 1) * 4 + ((off + t) & 1);
      cf[0] = fx; cf[2] = fy;
    }
  }
  __syncwarp();
  // filter: keys in [first key beyond the neighbours, rgx^2 (1 + delta)) over the lane's own windows
  const float lo = ok ? (float)tkey : 1.0f;
  const float hi = ok ? (float)(rgx * rgx * (1.0 + delta)) * 1.0000002f : -1.0f;
  int cnt = 0;
  bool ovf = false;
  for (int pc_ = 0; pc_ < npc; ++pc_) {
    const int rr = pc_ / nix, ix = ix0 + (pc_ - rr * nix), ru = r0 + rr;
    const int iy = g.wrapy ? img_idx(ru, g.ncy) : 0, row = ru - iy * g.ncy, uL = ix * g.ncx;
    const int a = max(c0, uL) - uL, b = min(c1, uL + g.ncx - 1) - uL;
    const int ca = max(clo, a + uL) - uL, cb = min(chi, b + uL) - uL;
    const int ps = __shfl_sync(0xffffffffu, P.s, pc_), po = __shfl_sync(0xffffffffu, P.off, pc_), pl = __shfl_sync(0xffffffffu, P.len, pc_);
    int ws = po, we = po;
    if (ok && ru >= rlo && ru <= rhi && ca <= cb) {
      ws = (int)cellStart[row * g.ncx + ca] - ps + po;
      we = (int)cellStart[row * g.ncx + cb + 1] - ps + po;
    }
    const int wal = ws & ~1;
    const int trips = __reduce_max_sync(0xffffffffu, (we - wal + 7) >> 3);
    const int pend8 = po + ((pl + 7) & ~7);
    int c = min(wal, pend8 - 8 * trips);
    for (int k = 0; k < trips; ++k, c += 8) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = candF4[(c >> 1) + u];
      if (cnt > ANN_CAP - 8) ovf = true;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 ax = __fadd2_rn(make_float2(v[u].x, v[u].y), nqx2), ay = __fadd2_rn(make_float2(v[u].z, v[u].w), nqy2);
        const float2 d2 = __ffma2_rn(ay, ay, __fmul2_rn(ax, ax));
        if (!ovf && d2.x >= lo && d2.x < hi) { col[cnt * 32] = (__float_as_uint(d2.x) & ~511u) | (uint32_t)(c + 2 * u); ++cnt; }
        if (!ovf && d2.y >= lo && d2.y < hi) { col[cnt * 32] = (__float_as_uint(d2.y) & ~511u) | (uint32_t)(c + 2 * u + 1); ++cnt; }
      }
    }
  }
  __syncwarp();
  if (ovf) ok = false;
  // beyond SPHB_KX candidates the farthest are dropped: largest packed entries
  const int nb = ok ? cnt : 0;
  uint32_t T = 0xffffffffu;
  {
    int mrem = ok && nb > SPHB_KX ? nb - SPHB_KX : 0;
    uint32_t bound = 0xffffffffu;
    bool done = mrem == 0;
    while (__any_sync(0xffffffffu, !done)) {
      uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
      const int lim = done ? 0 : nb;
      for (int s = 0; s < lim; ++s) {
        uint32_t k = col[s * 32];
        k = k < bound ? k : 0u;
        uint32_t a;
        a = max(t0, k); k = min(t0, k); t0 = a;
        a = max(t1, k); k = min(t1, k); t1 = a;
        a = max(t2, k); k = min(t2, k); t2 = a;
        a = max(t3, k); k = min(t3, k); t3 = a;
        t4 = max(t4, k);
      }
      if (!done) {
        if (mrem <= 5) { T = mrem == 1 ? t0 : (mrem == 2 ? t1 : (mrem == 3 ? t2 : (mrem == 4 ? t3 : t4))); done = true; }
        else { mrem -= 5; bound = t4; }
      }
    }
  }
  uint32_t* nsw = ex.nx + (size_t)tile * (SPHB_KX * 32) + lane;
  int w = 0;
  double dx = 0.0;
  if (ok) {
    double d2x = rgx * rgx * (1.0 - 1e-5);
    if (T != 0xffffffffu) d2x = fmin(d2x, (double)__uint_as_float(T & ~511u) * (1.0 - 2.0 * delta));
    for (int s = 0; s < nb; ++s) {
      const uint32_t e = col[s * 32];
      if (e < T && w < SPHB_KX) { nsw[w * 32] = candE[e & 511u]; ++w; }
    }
    dx = sqrt(d2x);
    // slab mode: the exclusion radius is only valid if everything within it was local (ghost layer wide enough)
    if (g.sides && (((g.sides & 1) && xa - g.ox < dx) || ((g.sides & 2) && g.ox + (double)g.ncx * g.dx - xa < dx)))
      atomicOr(dflags, DFLAG_GHOST_THIN);
  }
  if (valid) {
    for (; w < SPHB_KX; ++w) nsw[w * 32] = 0xffffffffu;
    ex.dexcl[i] = dx;
  }
}

// -------------------------------------------------------------------------------------------------
// Local displacement bound.  The global D is set by the fastest of all particles; what the certificate of particle i
// needs is a bound on |u_i - u_j| for the particles j that can reach it.  The cells of the rebuild's grid are grouped
// into coarse cells of 3 rows x (>= 3 rows' worth of) columns; after every step of a cycle k_ucum_bbox takes the bounding
// box of the cumulative displacements u of the particles of each coarse cell (by their cell at the rebuild: the sorted
// order does not change).  For particle i, L_i = largest distance from u_i to a corner of the union of the 3 x 3 boxes
// around its coarse cell bounds |u_i - u_j| for every j of that block; a particle outside the block was at least one
// coarse cell edge E away at the rebuild and has come closer by at most the global D.  So the certificate may use
// L_i instead of D whenever  E - D > h'.  (u is accumulated in fp32: the rounding, < 64 x 2^-24 of |u| over a cycle, is
// added to L_i.)
// -------------------------------------------------------------------------------------------------
// the fine cells of an axis are dealt out evenly: coarse index = (fine index x ng) / nc, so that EVERY coarse cell is
// at least floor(nc / ng) fine cells wide (a wrapping block must not have a narrow last cell on one side)
struct CoarseGeom { int ngx, ngy; double edge; };
__device__ __forceinline__ CoarseGeom coarse_geom(const GridP& g) {
  CoarseGeom c;
  const int fy = 3, fx = max(1, (int)ceil(3.0 * g.dy * g.inv_dx));
  c.ngx = max(1, g.ncx / fx); c.ngy = max(1, g.ncy / fy);
  c.edge = fmin((double)(g.ncx / c.ngx) * g.dx, (double)(g.ncy / c.ngy) * g.dy);
  return c;
}
__device__ __forceinline__ int coarse_of(int fine, int nc, int ng) { return (int)(((long long)fine * ng) / nc); }
__device__ __forceinline__ int coarse_first(int coarse, int nc, int ng) { return (int)(((long long)coarse * nc + ng - 1) / ng); }

// one warp per coarse cell (grid-stride): box[cell] = {min ux, max ux, min uy, max uy}; an empty cell gets an empty box
__global__ void __launch_bounds__(256) k_ucum_bbox(const float2* __restrict__ ucum, const uint32_t* __restrict__ cellStart,
                                                  const GridP* __restrict__ gp, float4* __restrict__ box, int box_cap,
                                                  uint32_t* __restrict__ ok_flag) {
  const GridP g = *gp;
  const CoarseGeom c = coarse_geom(g);
  const int ncoarse = c.ngx * c.ngy;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  if (gwarp == 0 && lane == 0) *ok_flag = (ncoarse <= box_cap && c.ngx >= 3 && c.ngy >= 3) ? 1u : 0u;
  if (ncoarse > box_cap) return;
  for (int cc = gwarp; cc < ncoarse; cc += nwarps) {
    const int gy = cc / c.ngx, gx = cc - gy * c.ngx;
    const int x0 = coarse_first(gx, g.ncx, c.ngx), x1 = coarse_first(gx + 1, g.ncx, c.ngx) - 1;
    float mnx = 3e38f, mxx = -3e38f, mny = 3e38f, mxy = -3e38f;
    for (int r = coarse_first(gy, g.ncy, c.ngy); r < coarse_first(gy + 1, g.ncy, c.ngy); ++r) {
      const int s = (int)cellStart[r * g.ncx + x0], e = (int)cellStart[r * g.ncx + x1 + 1];
      for (int t = s + lane; t < e; t += 32) {
        const float2 u = ucum[t];
        mnx = fminf(mnx, u.x); mxx = fmaxf(mxx, u.x); mny = fminf(mny, u.y); mxy = fmaxf(mxy, u.y);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) box[cc] = make_float4(mnx, mxx, mny, mxy);
  }
}

// L_i (see above) for the particle with cell key `key` and cumulative displacement u; 1.8e308 if not applicable
__device__ __forceinline__ double reuse_local_bound(const GridP& g, uint32_t key, float2 u, const float4* __restrict__ box,
                                                    double& edge) {
  const CoarseGeom c = coarse_geom(g);
  edge = c.edge;
  const int cy = (int)(key / (uint32_t)g.ncx), cx = (int)(key - (uint32_t)cy * (uint32_t)g.ncx);
  const int gx = coarse_of(cx, g.ncx, c.ngx), gy = coarse_of(cy, g.ncy, c.ngy);
  float mnx = u.x, mxx = u.x, mny = u.y, mxy = u.y;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    int y = gy + dy;
    if (g.wrapy) y = y < 0 ? y + c.ngy : (y >= c.ngy ? y - c.ngy : y);
    else if (y < 0 || y >= c.ngy) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      int x = gx + dx;
      if (g.wrapx) x = x < 0 ? x + c.ngx : (x >= c.ngx ? x - c.ngx : x);
      else if (x < 0 || x >= c.ngx) continue;
      const float4 b = __ldg(&box[y * c.ngx + x]);
      mnx = fminf(mnx, b.x); mxx = fmaxf(mxx, b.y); mny = fminf(mny, b.z); mxy = fmaxf(mxy, b.w);
    }
  }
  const double ex = fmax((double)u.x - (double)mnx, (double)mxx - (double)u.x), ey = fmax((double)u.y - (double)mny, (double)mxy - (double)u.y);
  const double umax = fmax(fmax(fabs((double)mnx), fabs((double)mxx)), fmax(fabs((double)mny), fabs((double)mxy)));
  return sqrt(ex * ex + ey * ey) * (1.0 + 1e-6) + 1.6e-5 * umax;
}

// -------------------------------------------------------------------------------------------------
// exact kNN(32) from the stored candidates.  One thread per particle; slots 0..31 (nn) hold the neighbours of the
// previous evaluation, slots 32..47 (nx) the further candidates.  Pass 1 evaluates d^2 of every candidate exactly as the
// reference does ((p + offset) - b, nearest periodic image; linear-algebra.go:61-64) into a shared-memory column and
// keeps, on integer keys, the 4 largest of the first group and the 4 smallest of the second.  The 32 smallest of the 48
// are the first group with its m largest exchanged for the m smallest of the second, m = number of crossing pairs
// (m = 4: refused).  Keys are the high words of the fp64 d^2 (fp32 build: the fp32 d^2): key order implies exact order
// when the keys differ, so the partition is certified by  max key kept < min key not kept  (ties / near ties: refused).
// Then: the exchanged entries swap places in the lists (nn stays "the 32 neighbours", any order), h^2 = max, density
// and sound speed as in the tile kernel, and the certificate  h + D < dexcl.
// -------------------------------------------------------------------------------------------------
#define REUSE_THREADS 128
#define REUSE_NC (SPHB_K + SPHB_KX)

template <typename T> __device__ __forceinline__ uint32_t reuse_key(T d2);
template <> __device__ __forceinline__ uint32_t reuse_key<double>(double d2) { return (uint32_t)__double2hiint(d2); }
template <> __device__ __forceinline__ uint32_t reuse_key<float>(float d2) { return __float_as_uint(d2); }

template <int KERNEL, bool F32>
__global__ void __launch_bounds__(REUSE_THREADS) k_knn_reuse(const double2* __restrict__ spos, const double* __restrict__ epred,
                                                            int n, const GridP* __restrict__ gp, PhysP ph, KnnOut out,
                                                            uint32_t* __restrict__ nx, const double* __restrict__ dexcl,
                                                            const ReuseState* __restrict__ rs,
                                                            const uint8_t* __restrict__ gflag, uint32_t* __restrict__ dflags,
                                                            const uint32_t* __restrict__ keys, const float2* __restrict__ ucum,
                                                            const float4* __restrict__ ubox, const uint32_t* __restrict__ ubox_ok) {
  typedef typename std::conditional<F32, float, double>::type TD;
  extern __shared__ __align__(16) unsigned char rsm[];
  TD* dcol = reinterpret_cast<TD*>(rsm) + threadIdx.x;  // [slot * REUSE_THREADS]
  const GridP g = *gp;
  const int i = blockIdx.x * REUSE_THREADS + threadIdx.x;
  const bool valid = i < n && (gflag == nullptr || gflag[i] != GF_OUTER);
  const bool owned = i < n && (gflag == nullptr || gflag[i] == GF_OWNED);
  const double D = rs->D;
  const int ii = valid ? i : 0;
  const double2 pa = spos[ii];
  const double dex = dexcl[ii];
  uint32_t* cn = out.nn + (size_t)(ii >> 5) * 1024 + (ii & 31);
  uint32_t* cx = nx + (size_t)(ii >> 5) * (SPHB_KX * 32) + (ii & 31);
  const bool wrap = g.wrapx | g.wrapy;  // (slab frames do not wrap: ghosts stand in for the images)
  const double hLx = 0.5 * g.Lx, hLy = 0.5 * g.Ly;
  // Only warps with a query next to the periodic seam look for images.  (Were a candidate beyond the seam after all,
  // its unshifted distance is about a period: the half-period test on the largest key below refuses the particle.)
  const double reach = 2.0 * dex + D;
  const bool near_seam = valid && ((g.wrapx && (pa.x - reach < g.lox || pa.x + reach >= g.lox + g.Lx)) ||
                                   (g.wrapy && (pa.y - reach < g.loy || pa.y + reach >= g.loy + g.Ly)));
  const bool img = wrap && __any_sync(0xffffffffu, near_seam);
  const double qxm = __dadd_rn(pa.x, g.Lx), qxp = __dadd_rn(pa.x, -g.Lx);  // candidate image -1 / +1: query + (-img L)
  const double qym = __dadd_rn(pa.y, g.Ly), qyp = __dadd_rn(pa.y, -g.Ly);
  const TD INF = F32 ? (TD)3.0e38f : (TD)1.7976931348623157e308;

  uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;                                          // largest keys of slots 0..31
  uint32_t b0 = 0xffffffffu, b1 = 0xffffffffu, b2 = 0xffffffffu, b3 = 0xffffffffu;  // smallest keys of slots 32..47
  uint32_t kmax = 0;  // largest key of any candidate
  bool empty_in = false;
  // the list entries are requested one batch ahead of their use, so that only the position gathers are waited for
  const bool live = valid && dex > 0.0;
  uint32_t en[8], en_next[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) en[u] = live ? cn[u * 32] : 0xffffffffu;
#pragma unroll 1
  for (int s0 = 0; s0 < REUSE_NC; s0 += 8) {
    uint32_t* cs = s0 < SPHB_K ? cn + s0 * 32 : cx + (s0 - SPHB_K) * 32;
    double2 pb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) pb[u] = spos[en[u] == 0xffffffffu ? ii : (int)(en[u] & IDX_MASK)];
    if (s0 + 8 < REUSE_NC) {
      const uint32_t* cs1 = s0 + 8 < SPHB_K ? cn + (s0 + 8) * 32 : cx + (s0 + 8 - SPHB_K) * 32;
#pragma unroll
      for (int u = 0; u < 8; ++u) en_next[u] = live ? cs1[u * 32] : 0xffffffffu;
    }
    if (!img) {  // away from the seam every candidate is a centre image: entries that still say otherwise (the pair has
      uint32_t stale = 0;  // crossed the seam together since the build) are reset, the force kernel reads the code
#pragma unroll
      for (int u = 0; u < 8; ++u) stale |= (en[u] >> IMG_SHIFT) ^ 5u;
      if (stale) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (en[u] != 0xffffffffu && (en[u] >> IMG_SHIFT) != 5u) cs[u * 32] = (en[u] & IDX_MASK) | (5u << IMG_SHIFT);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool have = en[u] != 0xffffffffu;
      double qx = pa.x, qy = pa.y;
      if (img) {  // nearest image from the current positions (a candidate may have crossed the seam since the build)
        int sx = 0, sy = 0;
        if (g.wrapx) { const double d0 = pa.x - pb[u].x; sx = d0 > hLx ? 1 : (d0 < -hLx ? -1 : 0); }
        if (g.wrapy) { const double d0 = pa.y - pb[u].y; sy = d0 > hLy ? 1 : (d0 < -hLy ? -1 : 0); }
        qx = sx == 0 ? pa.x : (sx < 0 ? qxm : qxp);
        qy = sy == 0 ? pa.y : (sy < 0 ? qym : qyp);
        const uint32_t ne = (en[u] & IDX_MASK) | (img_code(sx, sy) << IMG_SHIFT);
        if (have && ne != en[u]) cs[u * 32] = ne;
      }
      TD d2;
      if (F32) {
        const float fx = (float)(qx - pb[u].x), fy = (float)(qy - pb[u].y);
        d2 = (TD)fmaf(fy, fy, fx * fx);
      } else {
        d2 = (TD)dist_sq(qx - pb[u].x, qy - pb[u].y);
      }
      d2 = have ? d2 : INF;
      dcol[(s0 + u) * REUSE_THREADS] = d2;
      uint32_t k = reuse_key<TD>(d2), a;
      if (have) kmax = max(kmax, k);
      if (s0 < SPHB_K) {
        empty_in |= !have;
        a = max(t0, k); k = min(t0, k); t0 = a;
        a = max(t1, k); k = min(t1, k); t1 = a;
        a = max(t2, k); k = min(t2, k); t2 = a;
        t3 = max(t3, k);
      } else {
        a = min(b0, k); k = max(b0, k); b0 = a;
        a = min(b1, k); k = max(b1, k); b1 = a;
        a = min(b2, k); k = max(b2, k); b2 = a;
        b3 = min(b3, k);
      }
    }
  #pragma unroll
    for (int u = 0; u < 8; ++u) en[u] = en_next[u];
  }
  // crossing pairs: the k-th largest of the first group against the k-th smallest of the second
  const bool c0 = t0 > b0, c1 = c0 && t1 > b1, c2 = c1 && t2 > b2, c3 = c2 && t3 > b3;
  const int m = (int)c0 + (int)c1 + (int)c2;
  const uint32_t tIN = m == 0 ? 0xffffffffu : (m == 1 ? t0 : (m == 2 ? t1 : t2));   // keys >= tIN leave the first group
  const uint32_t tOUT = m == 0 ? 0u : (m == 1 ? b0 : (m == 2 ? b1 : b2));           // keys <= tOUT enter it
  const uint32_t in_max = max(m == 0 ? t0 : (m == 1 ? t1 : (m == 2 ? t2 : t3)), m == 0 ? 0u : tOUT);
  const uint32_t out_min = min(m == 0 ? b0 : (m == 1 ? b1 : (m == 2 ? b2 : b3)), tIN);
  bool ok = valid && !empty_in && !c3 && dex > 0.0;
  // strict key order <=> exact order (fp32 build: fp32 order; ties may fall either way there but must be consistent)
  if (!(in_max < out_min)) ok = false;
  // the nearest image is the only one in reach while every candidate is nearer than half a period
  if (wrap) {
    const double hl = fmin(g.wrapx ? hLx : 1.7976931348623157e308, g.wrapy ? hLy : 1.7976931348623157e308);
    const TD lim = (TD)(hl * hl * (1.0 - 1e-6));
    if (!(kmax < reuse_key<TD>(lim))) ok = false;
  }
  // masks of the exchanged slots and h^2 = largest d^2 kept
  uint32_t lmask = 0, emask = 0;
  TD h2 = (TD)0;
#pragma unroll 8
  for (int s = 0; s < SPHB_K; ++s) {
    const TD d = dcol[s * REUSE_THREADS];
    const bool leave = m > 0 && reuse_key<TD>(d) >= tIN;
    lmask |= leave ? (1u << s) : 0u;
    h2 = leave ? h2 : (d > h2 ? d : h2);
  }
#pragma unroll 8
  for (int s = 0; s < SPHB_KX; ++s) {
    const TD d = dcol[(SPHB_K + s) * REUSE_THREADS];
    const bool enter = m > 0 && reuse_key<TD>(d) <= tOUT;
    emask |= enter ? (1u << s) : 0u;
    h2 = enter ? (d > h2 ? d : h2) : h2;
  }
  if (__popc(lmask) != m || __popc(emask) != m) ok = false;  // equal keys inside a group
  if (ok) {
    while (lmask) {  // exchange (at most 3 pairs)
      const int sl = __ffs(lmask) - 1, so = __ffs(emask) - 1;
      lmask &= lmask - 1; emask &= emask - 1;
      const uint32_t ea = cn[sl * 32], eb = cx[so * 32];
      cn[sl * 32] = eb; cx[so * 32] = ea;
      const TD da = dcol[sl * REUSE_THREADS], db = dcol[(SPHB_K + so) * REUSE_THREADS];
      dcol[sl * REUSE_THREADS] = db; dcol[(SPHB_K + so) * REUSE_THREADS] = da;
    }
  }
  // certificate: nothing outside the candidate set can be within h
  double h = 0.0;
  if (ok) {
    h = F32 ? (double)sqrtf((float)h2) * 1.0000002 : sqrt((double)h2);
    double Di = D;
    if (ucum && *ubox_ok) {  // local displacement bound, where a particle from beyond the 3 x 3 block cannot reach h
      double edge;
      const double Li = reuse_local_bound(g, keys[i], ucum[i], ubox, edge);
      if (edge - D > h * (1.0 + 1e-9)) Di = fmin(D, Li);
    }
    if (!((h + Di) * (1.0 + 1e-12) < dex)) ok = false;
  }
  if (valid && !ok) {
    const int slot = atomicAdd(out.failCount, 1);
    out.failList[slot] = i;
  }
  if (ok) {
    // (slab mode: no geometric ghost-layer test here - the grid is the rebuild's, the particles have moved on.  The
    // rebuild verified that everything within dexcl was local; the certificate above does the rest, and the force kernel
    // still refuses a neighbour whose rho was never evaluated.)
    const double ep = epred[i];
    if (F32) {
      const float h2f = (float)h2;
      float inv_h;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_h) : "f"(h2f));
      float acc = 0.0f;
#pragma unroll 8
      for (int s = 0; s < SPHB_K; ++s) {
        const float q2 = (float)dcol[s * REUSE_THREADS] * (inv_h * inv_h);
        float rq;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(fmaxf(q2, 1e-30f)));
        const float q = fminf(q2 * rq, 1.0f);
        if (KERNEL == 0) acc += 1.0f;
        else if (KERNEL == 1) {
          const float lo = fmaf(q * q, q - 1.0f, 1.0f / 6.0f);
          const float t = 1.0f - q;
          acc += q < 0.5f ? lo : t * t * t * (1.0f / 3.0f);
        } else {
          const float t = 1.0f - q, tt = t * t;
          acc += tt * tt * fmaf(4.0f, q, 1.0f);
        }
      }
      const float hf = h2f * inv_h;
      h = (double)hf;
      const float rho = (float)(ph.Fpref * ph.mass) * acc * (inv_h * inv_h);
      const float c = sqrtf((float)(ph.cfac * ep));
      out.pc[i] = make_double4((double)rho, (double)c, h, (double)(c * c / ((float)ph.gamma * rho)));
    } else {
      const double h2d = (double)h2;
      const double inv_h = fast_rsqrt(h2d);
      h = fast_sqrt(h2d, inv_h);
      double acc = 0.0;
#pragma unroll 4
      for (int s = 0; s < SPHB_K; ++s) {
        const double d2 = (double)dcol[s * REUSE_THREADS];
        const double d = d2 * fast_rsqrt(d2 + 1e-300);  // coincident particles: d = 0
        acc += kern_F<KERNEL>(d * inv_h);
      }
      const double rho = ph.Fpref * ph.mass * acc * (inv_h * inv_h);
      const double c2 = ph.cfac * ep;
      const double c = c2 > 0.0 ? fast_sqrt(c2, fast_rsqrt(c2)) : sqrt(c2);
      out.pc[i] = make_double4(rho, c, h, c * c * fast_rcp(ph.gamma * rho));
    }
  }
  knn_accumulate_h(out, ok, owned, ok ? h : 0.0);
}

// bookkeeping after the force kernel of every step that may be followed by a reuse evaluation (one warp)
__global__ void k_reuse_update(ReuseState* __restrict__ rs, const GridP* __restrict__ gp, int n, int rebuild, const int* __restrict__ failCount,
                               const unsigned long long* __restrict__ hacc, double hscale, double coord_scale,
                               ReuseStat* __restrict__ stat) {
  long long sx = 0, sy = 0;
  for (int k = threadIdx.x; k < RS_SLOTS; k += 32) { sx += rs->sum[k][0]; sy += rs->sum[k][1]; rs->sum[k][0] = 0; rs->sum[k][1] = 0; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
  if (threadIdx.x != 0) return;
  const double dscale = 1048576.0 * gp->inv_dy;  // the force epilogue's fixed point
  const double mx = (double)sx / (dscale * (double)(n > 0 ? n : 1)), my = (double)sy / (dscale * (double)(n > 0 ? n : 1));
  const double M = (double)__uint_as_float(rs->Mbits);
  // this step's contribution: any two particles moved apart by at most 2 max|delta - mref|; the absolute term covers the
  // rounding of the displacement differences themselves
  const double Dstep = 2.0 * M * (1.0 + 1e-6) + 64.0 * 2.220446049250313e-16 * coord_scale;
  if (rebuild) { rs->D = Dstep; rs->ubx = mx; rs->uby = my; rs->age = 0; }
  else { rs->D += Dstep; rs->ubx += mx; rs->uby += my; rs->age += 1; }
  rs->mrx = mx; rs->mry = my;  // reference displacement of the next step
  rs->Mbits = 0;
  rs->seq += 1;
  if (stat) {
    ReuseStat r;
    r.seq = rs->seq; r.age = rs->age; r.refused = (unsigned)failCount[0]; r.n = (unsigned)n;  // (rebuild: refused by the tile search)
    r.D = (float)rs->D;
    r.dy = (float)gp->dy;
    r.rebuild = (unsigned)rebuild; r.seq2 = r.seq;
    *stat = r;
  }
}
