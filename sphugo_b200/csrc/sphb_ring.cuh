// sphb_ring.cuh — device helpers of the in-library slab ring (sphb_ring.inc): the per-step halo of a REUSE evaluation.
//
// A rebuild evaluation packs, exchanges and sorts its ghosts as before (k_pack_halo / k_add_ghosts / k_reorder) and
// remembers where everything went: halo_src[k] = sorted index of the k-th particle it sent, ghost_dst[k] = sorted index
// of the k-th ghost it received (both through the inverse permutation the reorder kernel leaves).  A reuse evaluation
// then sends the SAME particles in the SAME order - {position after drift-1, VPred, EPred}, 5 doubles - and the receiver
// overwrites its ghosts in place: fixed message sizes, no counts to exchange, no host round trip.
#pragma once

#define RING_REC 5  // doubles per record of a reuse-evaluation halo: x, y, vpred x, vpred y, epred

// src[k] = inv[idx[k]]  (pack order -> sorted order)
__global__ void __launch_bounds__(256) k_ring_translate(const int* __restrict__ idx, int count, const uint32_t* __restrict__ inv,
                                                       uint32_t* __restrict__ src) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) src[k] = inv[idx[k]];
}

__global__ void __launch_bounds__(256) k_ring_gather(const uint32_t* __restrict__ src, int count, const double2* __restrict__ pos,
                                                    const double2* __restrict__ vpred, const double* __restrict__ epred,
                                                    double* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t j = src[k];
  const double2 p = pos[j], v = vpred[j];
  double* r = buf + (size_t)k * RING_REC;
  r[0] = p.x; r[1] = p.y; r[2] = v.x; r[3] = v.y; r[4] = epred[j];
}

// the receiver's copy of a ghost is the owner's predicted state bit for bit; only the search position is taken in the
// receiver's own image frame
__global__ void __launch_bounds__(256) k_ring_scatter(const double* __restrict__ buf, int count, const uint32_t* __restrict__ dst,
                                                     const GridP* __restrict__ gp, double2* __restrict__ pos,
                                                     double2* __restrict__ spos, double2* __restrict__ vpred,
                                                     double* __restrict__ epred) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const GridP g = *gp;
  const uint32_t j = dst[k];
  const double* r = buf + (size_t)k * RING_REC;
  const double2 p = make_double2(r[0], r[1]);
  pos[j] = p;
  vpred[j] = make_double2(r[2], r[3]);
  epred[j] = r[4];
  double2 sp;
  sp.x = (g.wrapx | g.framex) ? wrap_coord(p.x, g.lox, g.Lx) : p.x;
  sp.y = g.wrapy ? wrap_coord(p.y, g.loy, g.Ly) : p.y;
  spos[j] = sp;
}

// indices of the owned particles while ghosts are interleaved with them (downloads in the middle of a cycle)
__global__ void __launch_bounds__(256) k_ring_owned_list(const uint8_t* __restrict__ gflag, int ntot, uint32_t* __restrict__ list,
                                                        int* __restrict__ counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = warp_append_slot(i < ntot && gflag[i] == GF_OWNED, counter);
  if (slot >= 0) list[slot] = (uint32_t)i;
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather_list(const T* __restrict__ in, const uint32_t* __restrict__ list, int count,
                                                    T* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) out[k] = in[list[k]];
}

// the device-side half of a rebuild's single host round trip: {max h, max |v|} for the max all-reduce, {refused, n,
// row height x n} for the sum all-reduce, the two pack counts for the neighbours, the status word - gathered into one
// block so that ONE copy and ONE wait bring everything to the host after the NCCL group
struct RingMsg {
  double vmax[2];      // in / out of the max all-reduce
  double vsum[3];      // in / out of the sum all-reduce
  int cnt_out[2];      // records packed for the low / high neighbour
  int cnt_in[2];       // records the left / right neighbour packed for this rank (ncclRecv)
  unsigned dflags;
  int pad;
};
__global__ void k_ring_msg(RingMsg* __restrict__ m, const uint32_t* __restrict__ qmax, int qmax_valid, const int* __restrict__ failCount,
                           int pending, int n, const GridP* __restrict__ gp, const int* __restrict__ packCount,
                           const uint32_t* __restrict__ dflags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  m->vmax[0] = qmax_valid ? (double)__uint_as_float(qmax[0]) : 0.0;
  m->vmax[1] = qmax_valid ? sqrt((double)__uint_as_float(qmax[1])) * (1.0 + 1e-7) : 0.0;
  m->vsum[0] = pending ? (double)failCount[0] : 0.0;
  m->vsum[1] = (double)n;
  m->vsum[2] = pending ? gp->dy * (double)n : 0.0;
  m->cnt_out[0] = packCount[0]; m->cnt_out[1] = packCount[1];
  m->cnt_in[0] = 0; m->cnt_in[1] = 0;
  m->dflags = *dflags;
}
