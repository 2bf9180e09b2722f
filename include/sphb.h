/* sphb.h — C ABI of libsphb.so, the B200 (sm_100a) SPH step for bbeni/sphugo.
 *
 * This is the drop-in boundary for ONE hot path of the reference: everything
 * (*Simulation).Step() does (reference sim/sph.go:64-198), i.e. spatial index,
 * periodic kNN (k = 32), density, pressure + artificial-viscosity force and the
 * leapfrog drift/kick with periodic wrap and reflections.  The reference has no
 * FFI of its own (it is pure Go); these entry points are what a cgo shim in
 * package `sim` binds (see INTEGRATION.md and go/sim/backend_cuda.go).
 *
 * Conventions
 *   - every function returns 0 (SPHB_OK) or a negative sphb_status; the text
 *     of the last failure is sphb_last_error(sim) (library-owned string).
 *   - plain pointers and sizes only; host pointers are borrowed for the
 *     duration of the call (cgo rule: no Go pointer is retained).
 *   - there is NO CPU fallback in this library: without a CUDA device
 *     sphb_create fails with SPHB_E_CUDA.
 *   - one call at a time per handle (the Go shim holds sim.IsBusy, sph.go:20,66).
 *   - every entry point calls cudaSetDevice (goroutines hop OS threads).
 *   - particle order on the device is cell order and changes every step, like
 *     the reference's in-place tree partition (core.go:126-164); join on `id`
 *     (the reference's Particle.Z, core.go:41).
 */
#ifndef SPHB_H
#define SPHB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB_NN 32 /* NN_SIZE, reference sim/core.go:14 */

/* open axis marker: == Go math.MaxFloat64 (config-parser.go:139-145) */
#define SPHB_OPEN_HI 1.7976931348623157e308
#define SPHB_OPEN_LO (-1.7976931348623157e308)

typedef enum {
  SPHB_OK = 0,
  SPHB_E_INVALID = -1,       /* bad argument / bad handle / half-open axis (nearest-neighbour.go:44,53 panics) */
  SPHB_E_CUDA = -2,          /* CUDA runtime error or no device */
  SPHB_E_NOMEM = -3,         /* capacity exceeded / allocation failed */
  SPHB_E_KNN_UNDERFULL = -4, /* fewer than 32 (particle,image) candidates exist: the reference would
                                keep sentinel slots (nearest-neighbour.go:155-165); we refuse instead */
  SPHB_E_KERNEL = -5,        /* TopHat2D has no derivative (sph.go:251-253 panics) */
  SPHB_E_STATE = -6          /* call order: e.g. density before knn */
} sphb_status;

typedef enum {
  SPHB_KERNEL_TOPHAT = 0,   /* sim.TopHat2D   sph.go:244-257 (density only) */
  SPHB_KERNEL_MONAGHAN = 1, /* sim.Monahan2D  sph.go:259-276 */
  SPHB_KERNEL_WENDLAND = 2  /* sim.Wendtland2D sph.go:284-304 */
} sphb_kernel;

/* Numeric part of sim.SphConfig (config-parser.go:111-128). */
typedef struct {
  double dt_half;        /* DeltaTHalf */
  double gamma;          /* Gamma */
  double particle_mass;  /* ParticleMass */
  double accel[2];       /* Acceleration */
  double hor[2];         /* HorPeriodicity;  {SPHB_OPEN_LO, SPHB_OPEN_HI} = open */
  double ver[2];         /* VertPeriodicity */
  double refl_L, refl_R; /* Reflections.L/.R on x; disabled = SPHB_OPEN_LO / SPHB_OPEN_HI */
  double refl_U, refl_D; /* Reflections.U (low y) / .D (high y) */
  int32_t kernel;        /* sphb_kernel */
  int32_t precision;     /* 64: everything fp64 (reference precision). 32: pair arithmetic in fp32 on
                            cell-relative coordinates, state kept in fp64 */
  int32_t device;        /* CUDA device ordinal of this handle (one process / handle per GPU) */
  int32_t flags;         /* SPHB_FLAG_* */
} sphb_params;

#define SPHB_FLAG_KEEP_NN_LIST 1u /* always materialise the neighbour list (needed by download of NN_*) */
#define SPHB_FLAG_REUSE_LISTS 2u  /* certified reuse of the neighbour lists between rebuilds (exact kNN from stored
                                     candidates under a displacement certificate; DESIGN.md "list reuse").  Results are
                                     the same exact kNN either way; off by default - see DESIGN.md for what it buys */

typedef struct sphb_sim sphb_sim; /* opaque, owned by the library */

/* Fields for sphb_download / sphb_upload: bit i of the mask <-> host_ptrs[i]. */
enum {
  SPHB_F_POS = 0,    /* double[2n]  Particle.Pos            */
  SPHB_F_VEL = 1,    /* double[2n]  Particle.Vel            */
  SPHB_F_RHO = 2,    /* double[n]   Particle.Rho            */
  SPHB_F_C = 3,      /* double[n]   Particle.C              */
  SPHB_F_E = 4,      /* double[n]   Particle.E              */
  SPHB_F_EDOT = 5,   /* double[n]   Particle.EDot           */
  SPHB_F_VDOT = 6,   /* double[2n]  Particle.VDot           */
  SPHB_F_EPRED = 7,  /* double[n]   Particle.EPred          */
  SPHB_F_VPRED = 8,  /* double[2n]  Particle.VPred          */
  SPHB_F_H = 9,      /* double[n]   Particle.NNDists[0]     */
  SPHB_F_ID = 10,    /* int64[n]    Particle.Z              */
  SPHB_F_NN_IDX = 11,/* int32[32n]  index (current device order) of Particle.NearestNeighbours[k], -1 = none; slots in
                                     the reference's order: descending distance, slot 0 = farthest = h
                                     (nearest-neighbour.go:139-153) */
  SPHB_F_NN_DIST = 12,/* double[32n] Particle.NNDists[k] for the same slots */
  SPHB_F_NN_POS = 13, /* double[64n] Particle.NNPos[k] (neighbour image position in the query frame).  Like the reference's
                         field it describes the EVALUATION (nearest-neighbour.go:80): after sphb_knn / sphb_calc_forces it
                         is exact; after sphb_step the particle has since been kicked and drifted, and the value returned
                         is its current Pos plus the evaluation-time offset to the neighbour (a particle that wrapped or
                         was reflected in that step carries the jump) */
  SPHB_F_COUNT = 14
};
#define SPHB_MASK(f) (1u << (f))

/* reductions, sph.go:441-463 */
enum { SPHB_SUM_E = 0, SPHB_SUM_RHO = 1, SPHB_LAST_VEL_NORM = 2 /* TotalMomentum's `=` bug, sph.go:460 */ };

/* phases reported by sphb_phase_times (milliseconds, CUDA events, last step) */
enum {
  SPHB_PH_KEYS = 0,   /* drift-1 + cell keys                       (replaces Partition, core.go:126) */
  SPHB_PH_SORT = 1,   /* counting sort by cell                      (replaces Treebuild, core.go:172) */
  SPHB_PH_REORDER = 2,/* SoA gather + predict + cell table                                         */
  SPHB_PH_KNN = 3,    /* kNN + density + sound speed (+ fallback)   (nearest-neighbour.go:28, sph.go:306,423) */
  SPHB_PH_FORCE = 4,  /* force + kick + drift-2 + boundaries        (sph.go:327, 122-193) */
  SPHB_PH_TOTAL = 5,
  SPHB_PH_COUNT = 6
};

/* counters reported by sphb_counters (cumulative since create) */
enum {
  SPHB_CNT_STEPS = 0,
  SPHB_CNT_KERNEL_LAUNCHES = 1,
  SPHB_CNT_KNN_FALLBACK = 2, /* particles that needed the ring-expansion search */
  SPHB_CNT_REGRIDS = 3,
  SPHB_CNT_REUSE_STEPS = 4,  /* evaluations that took the exact kNN from the stored candidate lists (no sort, no search) */
  SPHB_CNT_COUNT = 5
};

/* == sim.MakeSimulationFromConf + MakeCells (sph.go:40-54, core.go:93-105).
 * pos_xy, vel_xy: interleaved x,y. vel/e/rho/id may be NULL (zeros; id = index).
 * capacity >= n reserves room for sphb_append (Sources, sph.go:72-86); an append beyond it reallocates at
 * twice the size, like Go's append. */
int sphb_create(const sphb_params* p, int64_t n, int64_t capacity, const double* pos_xy,
                const double* vel_xy, const double* e, const double* rho, const int64_t* id,
                sphb_sim** out);
void sphb_destroy(sphb_sim* s);
const char* sphb_last_error(const sphb_sim* s); /* s may be NULL: error of the last failed create */

/* sim.Config is a public mutable field (sph.go:15); precision/device cannot change. */
int sphb_set_params(sphb_sim* s, const sphb_params* p);
int sphb_get_params(const sphb_sim* s, sphb_params* p);
int64_t sphb_count(const sphb_sim* s);        /* len(sim.Root.Particles) */
int64_t sphb_current_step(const sphb_sim* s); /* sim.CurrentStep */
/* for a handle created from a simulation that has already stepped (sim.CurrentStep > 0: the Go shim re-creates its device
 * copy when simviewer replaces the Simulation value): sets the counter, so that the next sphb_step does not run the
 * step-0 initialisation (VPred = Vel, EPred = E, forces once; sph.go:89-103) again.  Upload VDOT / EDOT as well. */
int sphb_set_current_step(sphb_sim* s, int64_t step);

/* == append(sim.Root.Particles, spawned...) + MakeCells (sph.go:75-86).  Beyond the capacity the device arrays are
 * reallocated at twice the size (SPHB_E_NOMEM if that fails or 2^28-1 particles per device would be exceeded; the
 * handle is unchanged then).  Not while ghosts are attached (slab mode, between step_begin and step_end). */
int sphb_append(sphb_sim* s, int64_t n, const double* pos_xy, const double* vel_xy, const double* e,
                const double* rho, const int64_t* id);

/* == (*Simulation).Step() x nsteps, including the step-0 special case (sph.go:64-198). Asynchronous:
 * returns after enqueueing; sphb_sync / download / reduce wait for it. */
int sphb_step(sphb_sim* s, int32_t nsteps);
/* == (*Simulation).CalculateForces() (sph.go:403-435) */
int sphb_calc_forces(sphb_sim* s);
/* batch == for all i: Particles[i].FindNearestNeighboursPeriodic(root, hor, ver) (nearest-neighbour.go:28-67);
 * both {SPHB_OPEN_LO,SPHB_OPEN_HI} == FindNearestNeighbours (:15-23). Materialises the neighbour list. */
int sphb_knn(sphb_sim* s, const double hor[2], const double ver[2]);
/* batch == for all i: Particles[i].Rho = Density2D(p, sim, kernel) (sph.go:306-323); needs a prior knn */
int sphb_density(sphb_sim* s, int32_t kernel);

/* the CUDA stream (cudaStream_t) every kernel of this handle is enqueued on: lets a caller record its own
 * CUDA events around asynchronous sphb_step calls or order its own work (NCCL exchange) after them */
void* sphb_stream(sphb_sim* s);
int sphb_sync(sphb_sim* s); /* wait for the device; surfaces asynchronous errors (kNN underfull, ...) */

/* copy the selected fields into caller-owned host buffers (each sized for `capacity` particles);
 * *n_out = number of particles. Order = current device order; join on SPHB_F_ID. */
int sphb_download(sphb_sim* s, uint32_t field_mask, void* const* host_ptrs, int64_t capacity,
                  int64_t* n_out);
/* overwrite device state (same order as the last download) for POS, VEL, E, RHO, VDOT, EDOT:
 * Root.Particles is public and examples poke it directly (density.go:12-15). */
int sphb_upload(sphb_sim* s, uint32_t field_mask, const void* const* host_ptrs, int64_t n);
/* the same for a caller that keeps its particles in a fixed order of its own: host array element k belongs to the
 * particle with id k.  Needs dense ids (a permutation of 0..N-1, e.g. the ids sphb_create assigns when id == NULL;
 * verified on the device, SPHB_E_STATE otherwise).  POS, VEL, E only.  The device order changes with every sort, so
 * this - not sphb_upload - is the call for "Go mutated Root.Particles between two steps". */
int sphb_upload_by_id(sphb_sim* s, uint32_t field_mask, const void* const* host_ptrs, int64_t n);

/* sphb_upload_by_id in two halves, for a caller that has the next step's particle fields ready while the current step is
 * still running (a replay, a coupled solver on the host): _begin starts the host-to-device copies on a copy stream of the
 * handle, into a staging area of their own, and returns; steps enqueued before or after it keep running.  _end waits for
 * the copies (the host buffers are borrowed from _begin to _end), then scatters the staged fields into the device order
 * behind the steps enqueued so far.  One upload in flight per handle. */
int sphb_upload_by_id_begin(sphb_sim* s, uint32_t field_mask, const void* const* host_ptrs, int64_t n);
int sphb_upload_by_id_end(sphb_sim* s);

int sphb_reduce(sphb_sim* s, int32_t which, double* out); /* TotalEnergy / TotalDensity / TotalMomentum */

/* Per-particle frame data as (*Animator).CurrentFrame derives it (sim/animator.go:75-101), computed on the device so
 * that a frame costs 9 (+8 with ids) bytes per particle over the bus instead of the 40 of {Pos, Rho, NNDists[0], Z}:
 *   xy_out[2 i], xy_out[2 i + 1] = float32(Pos.X) * float32(width), float32(Pos.Y) * float32(height)
 *   colour_out[i]                = uint8(min(Rho / (ParticleMass * N * 10) * 256, 255))        (the ramp index)
 *   id_out[i] (may be NULL)      = Z: the renderer draws in order of descending Z
 * Host buffers sized for `capacity` particles; *n_out = N.  Order = current device order when id_out is given.
 * id_out == NULL: element k describes the particle with id k (dense ids required, as for sphb_upload_by_id) - a
 * caller that numbers its particles in drawing order receives the frame ready to rasterise, 9 bytes per particle. */
int sphb_frame(sphb_sim* s, int32_t width, int32_t height, float* xy_out, uint8_t* colour_out, int64_t* id_out,
               int64_t capacity, int64_t* n_out);
int sphb_phase_times(sphb_sim* s, double* ms, int32_t n); /* n <= SPHB_PH_COUNT */
int sphb_counters(const sphb_sim* s, int64_t* out, int32_t n); /* n <= SPHB_CNT_COUNT */

/* device-resident variants for callers that already hold device memory (bench.py `value`, the
 * multi-GPU slab driver): pointers are CUDA device pointers on the handle's device. */
int sphb_create_device(const sphb_params* p, int64_t n, int64_t capacity, const double* d_pos_xy,
                       const double* d_vel_xy, const double* d_e, const int64_t* d_id, sphb_sim** out);

/* ---- slab decomposition: one handle per GPU owns a set of particles, nominally those with x in
 * [x_lo, x_hi).  The exchange itself (NCCL send/recv over NVLink) is done by the caller on device buffers; see
 * sphugo_b200/slab.py.  One force evaluation:
 *     sphb_slab_set            ghost widths for this evaluation (from the all-reduced max h)
 *     sphb_slab_step_begin     selects the evaluation mode (drift-1 + predict run inside step_end, sph.go:108-117)
 *     sphb_slab_pack_halo      owned particles within ghost_w of an edge -> records   -> exchange
 *     sphb_slab_add_ghosts x2  append the neighbour's records as ghosts
 *     sphb_slab_step_end       sort, kNN, density on owned + inner ghosts, forces + kick + drift-2 + boundaries on
 *                              owned, ghosts dropped                                   (sph.go:119-193)
 * Ghost layer: ghosts within inner_w of the slab are full queries (their rho, c, h are recomputed redundantly, so
 * no second exchange is needed), the rest up to ghost_w are candidates only; inner_w >= max h and
 * ghost_w >= inner_w + max h are verified on the device (SPHB_E_GHOST_THIN otherwise).
 * Ownership is a set, not a strict interval: a particle that drifts out of [x_lo, x_hi) stays owned (the ghost
 * widths must cover the excursion) until the caller migrates it: pack_migrants x2 -> exchange -> add_migrants x2 ->
 * finish_migration.  Positions are always the reference's global coordinates; with a periodic x axis each handle
 * works in the image frame centred on its slab. */
typedef struct {
  double x_lo, x_hi;        /* nominal owned interval */
  double ghost_w, inner_w;  /* ghost layer width / width of the layer whose ghosts are evaluated */
  int32_t has_left;         /* a neighbour slab exists on the low-x side (periodic ring or interior edge) */
  int32_t has_right;
} sphb_slab;

#define SPHB_E_GHOST_THIN (-7) /* a smoothing length reached past the ghost layer: widen and redo the evaluation */

int sphb_slab_set(sphb_sim* s, const sphb_slab* slab);
/* max smoothing length of the owned particles after the last evaluation (device reduction; the caller all-reduces) */
int sphb_max_h(sphb_sim* s, double* out);
/* largest |Vel| of the owned particles (device reduction): bounds how far a particle can leave its slab per step,
 * so the caller can migrate only when the accumulated excursion approaches the ghost-layer slack */
int sphb_max_speed(sphb_sim* s, double* out);
/* mode 0: CalculateForces on the state as is; 1: step-0 initialisation VPred = Vel, EPred = E (sph.go:97-100);
 * 2: drift-1 + predict (sph.go:108-117).  The mode applies to the pack / add_ghosts / step_end calls that follow:
 * ghosts travel with the predictor's inputs and are drifted and predicted by the receiver together with its owned
 * particles, with identical arithmetic. */
int sphb_slab_step_begin(sphb_sim* s, int32_t mode);
/* one pass: pack the owned particles that will lie within ghost_w of the low / high edge at evaluation time as ghost
 * records {x, y, vx, vy, vdotx, vdoty, e, edot, id (int64 bits), h_prev} = 10 x 8 bytes (mode 0: VPred / EPred in the
 * velocity / energy slots, zero derivatives) into d_buf_lo / d_buf_hi (device); count_out[0..1] = records written */
int sphb_slab_pack_halo(sphb_sim* s, void* d_buf_lo, void* d_buf_hi, int64_t cap_records, int64_t* count_out);
int sphb_slab_add_ghosts(sphb_sim* s, const void* d_buf, int64_t count);
/* integrate != 0: kick + drift-2 + wrap + reflections on the owned particles and CurrentStep += 1 */
int sphb_slab_step_end(sphb_sim* s, int32_t integrate);
/* pack owned particles whose x (slab frame) left [x_lo, x_hi) towards side 0/1 as migration records
 * {x, y, vx, vy, e, vdotx, vdoty, edot, h, id, rho, c} = 12 x 8 bytes; they are removed by finish_migration */
int sphb_slab_pack_migrants(sphb_sim* s, int32_t side, void* d_buf, int64_t cap_records, int64_t* count_out);
int sphb_slab_add_migrants(sphb_sim* s, const void* d_buf, int64_t count);
int sphb_slab_finish_migration(sphb_sim* s);

#define SPHB_HALO_RECORD_DOUBLES 10
#define SPHB_MIGRANT_RECORD_DOUBLES 12

/* ---- the slab ring inside the library: multi-GPU behind this ABI (SURVEY §8b `n_devices`, §8e).
 * One process and one handle per GPU; every rank creates its handle with the particles of its x-slab, describes the ring
 * (sphb_ring_set), joins the communicator (sphb_comm_init: NCCL, loaded with dlopen - libsphb.so has no link-time NCCL
 * dependency) and from then on calls sphb_ring_step where a single-GPU caller calls sphb_step: halo exchange, ghost
 * handling, migration, the all-reduces and the schedule of the neighbour-list reuse all happen inside, on the handle's
 * stream (ncclSend / ncclRecv / ncclAllReduce).  Only the rebuild evaluation of a cycle waits for the host (counts, max
 * h / max speed); the reuse evaluations send fixed-size messages and are asynchronous like sphb_step.
 * Collective contract: sphb_ring_step and every state-changing call (append, upload, set_params, knn, calc_forces) are
 * made by all ranks between the same two steps; download / reduce / counters may be called by any rank at any time. */
typedef struct {
  int32_t rank, nranks;   /* position in the ring of x-slabs, ascending x */
  int32_t periodic;       /* the x axis wraps: rank nranks-1 and rank 0 are neighbours */
  int32_t migrate_every;  /* 0: migrate when the excursion bound nears the ghost-layer slack; 1: at every rebuild */
  double x_lo, x_hi;      /* this rank's nominal slab */
  double h_hint;          /* first guess of the largest smoothing length (until one has been measured) */
  double safety;          /* ghost layer = safety x max h (+ slack); 0 = default 1.15 */
  int64_t halo_cap;       /* records per side of the exchange buffers */
} sphb_ring;

#define SPHB_E_NCCL (-8) /* NCCL missing or a NCCL call failed */
#define SPHB_NCCL_ID_BYTES 128
int sphb_comm_unique_id(void* id_out /* SPHB_NCCL_ID_BYTES */); /* ncclGetUniqueId: rank 0 makes it, the caller distributes it */
int sphb_comm_init(sphb_sim* s, const void* unique_id, int32_t rank, int32_t nranks); /* ncclCommInitRank on the handle's device */
int sphb_ring_set(sphb_sim* s, const sphb_ring* ring);
int sphb_ring_step(sphb_sim* s, int32_t nsteps); /* == Step() x nsteps of the whole ring; called by every rank */
/* all slabs of the ring in one process (handles[k] = rank k, any devices): the exchange is a device copy.  Used by the
 * single-GPU tests of the whole protocol. */
int sphb_ring_step_local(sphb_sim* const* handles, int32_t n, int32_t nsteps);
enum { SPHB_RING_H_MAX = 0, SPHB_RING_V_MAX = 1, SPHB_RING_MIGRATIONS = 2, SPHB_RING_GHOSTS = 3, SPHB_RING_PERIOD = 4,
       SPHB_RING_EXCURSION = 5, SPHB_RING_INFO_COUNT = 6 };
int sphb_ring_info(const sphb_sim* s, double* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SPHB_H */
