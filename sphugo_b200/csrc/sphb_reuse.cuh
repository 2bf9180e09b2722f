// sphb_reuse.cuh — kernels of a REUSE evaluation (certified reuse of the neighbour lists, see ReuseState in
// sphb_kernels.cuh).  Pipeline of such a step, in place of keys / sort / reorder / tile search:
//   k_predict      drift-1 + predict, streaming, particle order unchanged            (sph.go:108-117)
//   k_knn_reuse    exact kNN(32) from the <= 48 stored candidates + density + sound speed, certificate per particle
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
                                                            const uint8_t* __restrict__ gflag, uint32_t* __restrict__ dflags) {
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
#pragma unroll 1
  for (int s0 = 0; s0 < REUSE_NC; s0 += 8) {
    uint32_t* cs = s0 < SPHB_K ? cn + s0 * 32 : cx + (s0 - SPHB_K) * 32;
    uint32_t en[8];
    double2 pb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) en[u] = valid ? cs[u * 32] : 0xffffffffu;
#pragma unroll
    for (int u = 0; u < 8; ++u) pb[u] = spos[en[u] == 0xffffffffu ? ii : (int)(en[u] & IDX_MASK)];
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
    if (!((h + D) * (1.0 + 1e-12) < dex)) ok = false;
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
    unsigned long long hs = 0, hc = 0;
    if (hacc && hscale > 0.0) for (int k = 0; k < HACC_N; ++k) { hs += hacc[2 * k]; hc += hacc[2 * k + 1]; }
    r.hmean = hc ? (float)((double)hs / hscale / (double)hc) : 0.0f;
    r.rebuild = (unsigned)rebuild; r.seq2 = r.seq;
    *stat = r;
  }
}
