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
