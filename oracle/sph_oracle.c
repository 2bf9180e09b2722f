/* sph_oracle.c — CPU restatement of bbeni/sphugo's SPH step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (sphugo_b200/, libsphb.so) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or as the
 * timed CPU baseline.
 *
 * Parity status: the reference is pure Go and Go is not installed in the build container,
 * so the reference itself cannot be run here.  This restatement is pinned against every
 * fixture the reference's own tests hold for the path:
 *   - sim/partition_test.go:11-160  (13 Partition half-length KATs)      -> orc_partition
 *   - sim/bounding-sphere_test.go:30-64 (inside-any-circle property)     -> orc_tree_all_inside_any
 *   - README.md:88-162 (heap BuildHeap/Insert/ExtractMin/Replace KAT)    -> orc_heap_*
 * and against the reference's recorded OUTPUT for kNN and density: the PNGs its Go binary drew
 * for examples/density (doc/density_compare.png, density_test.png, density_test_periodic.png)
 * are reproduced pixel for pixel from this oracle's results (tests/test_reference_images.py),
 * as are doc/tree.png (Partition / Treebuild) and doc/nearest_neighbours*.png (leaf bounding
 * circles, open and periodic neighbour sets of one particle).
 * The reference has NO test and no recorded output that pins force, leapfrog or boundaries, so
 * for those functions parity is UNPINNED by the reference: they are restated line by line below
 * (same operation order, no FMA contraction: build with -ffp-contract=off) and cross-checked
 * against an independent exact brute-force kNN (mode 1) and scipy's cKDTree in tests/, and
 * against a second restatement of the same Go functions written separately in numpy
 * (tests/np_restatement.py), with which every field agrees bit for bit over several steps
 * (tests/test_oracle_crosscheck.py).  The small scenes use the reference's own inputs: Go's
 * math/rand stream, reconstructed and pinned by the reference's recorded output (README.md:89,
 * the rand.Seed(101) draws of examples/heap) and Go's published known answers
 * (sphugo_b200/gorand.py, tests/test_gorand.py).
 *
 * Layout and algorithm follow the reference so that the timing of orc_step is an honest
 * "reference CPU path" figure: 1136-byte AoS particle (core.go:17-42), in-place Partition of
 * whole structs, recursive midpoint bisection tree with leaf <= 8, bounding circles, recursive
 * pruned tree walk per particle and per periodic image, sorted-array top-32 queue.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define NN_SIZE 32                 /* core.go:14 */
#define MAX_PARTICLES_PER_CELL 8   /* core.go:11 */
#define SPLIT_FRACTION 0.5         /* core.go:12 */

typedef struct { double X, Y; } Vec2; /* linear-algebra.go:36-38 */

typedef struct Particle {          /* core.go:17-42, same field order and size (1136 B) */
  Vec2 Pos, Vel;
  double Rho, C, E;
  double EDot;
  Vec2 VDot;
  double EPred;
  Vec2 VPred;
  struct Particle* NearestNeighbours[NN_SIZE];
  double NNDists[NN_SIZE];
  Vec2 NNPos[NN_SIZE];
  int64_t Z;
} Particle;

typedef struct Cell {              /* core.go:46-60 */
  Particle* Particles;
  int64_t Len;
  Vec2 LowerLeft, UpperRight;
  Vec2 BCenter;
  double BRadius;
  struct Cell *Lower, *Upper;
} Cell;

enum { Vertical = 0, Horizontal = 1 }; /* core.go:62-67 */

typedef struct {                   /* == sphb_params (include/sphb.h), mirrored to keep this file standalone */
  double dt_half, gamma, particle_mass, accel[2];
  double hor[2], ver[2];
  double refl_L, refl_R, refl_U, refl_D;
  int32_t kernel, precision, device, flags;
} orc_params;

enum { ORC_OK = 0, ORC_PANIC_HALF_OPEN = -1, ORC_PANIC_Q_RANGE = -2, ORC_PANIC_TOPHAT_DF = -3,
       ORC_PANIC_NOT_INIT = -4, ORC_E_NOMEM = -5 };

/* ---------- node arena (the reference heap-allocates children on every rebuild) ---------- */
#define ARENA_CHUNK 65536
typedef struct ArenaChunk { struct ArenaChunk* next; int used; Cell cells[ARENA_CHUNK]; } ArenaChunk;

typedef struct orc_sim {
  orc_params cfg;
  Particle* ps;
  int64_t n, cap;
  Cell root;               /* persistent root, sph.go:17 */
  ArenaChunk* arena;
  int64_t current_step;
  int status;              /* first "panic" seen */
  int stale_root_child;    /* quirk: core.go:189,207 would keep a stale child here (not reproduced) */
  int64_t underfull;       /* particles whose queue kept >= 1 sentinel slot in the last kNN */
} orc_sim;

static void arena_reset(orc_sim* s) {
  ArenaChunk* c = s->arena;
  while (c) { ArenaChunk* nx = c->next; free(c); c = nx; }
  s->arena = NULL;
}
static Cell* arena_new(orc_sim* s) {
  if (!s->arena || s->arena->used == ARENA_CHUNK) {
    ArenaChunk* c = (ArenaChunk*)malloc(sizeof(ArenaChunk));
    if (!c) abort();
    c->next = s->arena; c->used = 0; s->arena = c;
  }
  Cell* cell = &s->arena->cells[s->arena->used++];
  memset(cell, 0, sizeof(Cell));
  return cell;
}

/* ---------- linear-algebra.go:61-71 ---------- */
static inline double DistSq(Vec2 a, Vec2 b) { double dx = a.X - b.X, dy = a.Y - b.Y; return dx * dx + dy * dy; }
static inline double Dist(Vec2 a, Vec2 b) { return hypot(a.X - b.X, a.Y - b.Y); } /* math.Hypot: pruning only */

/* ---------- Partition, core.go:126-164 (verbatim control flow, incl. the Vertical/Horizontal asymmetry) ---------- */
static int64_t partition(Particle* ps, int64_t len, int orientation, double middle) {
  int64_t i = 0, j = len - 1;
  Particle tmp;
  if (orientation == Vertical) {
    while (i < j) {
      while (i < j && ps[i].Pos.Y <= middle) i++;
      while (i < j && ps[j].Pos.Y > middle) j--;
      if (ps[i].Pos.Y > ps[j].Pos.Y) { tmp = ps[i]; ps[i] = ps[j]; ps[j] = tmp; }
      if (i == j && middle > ps[i].Pos.Y) i++;
    }
  } else {
    while (i < j) {
      while (i < j && ps[i].Pos.X <= middle) i++;
      while (i < j && ps[j].Pos.X > middle) j--;
      if (ps[i].Pos.X > ps[j].Pos.X) { tmp = ps[i]; ps[i] = ps[j]; ps[j] = tmp; }
    }
    if (len > 0 && i == j && middle > ps[i].Pos.X) i++; /* Go would index ps[0] of an empty slice only if i==j==0<len */
  }
  return i; /* a = ps[:i], b = ps[i:] */
}

/* ---------- Treebuild, core.go:172-224 ---------- */
static void treebuild(orc_sim* s, Cell* root, int orientation) {
  if (root->LowerLeft.Y == root->UpperRight.Y && root->LowerLeft.X == root->UpperRight.X) return;
  double mid;
  if (orientation == Vertical) mid = SPLIT_FRACTION * root->LowerLeft.Y + (1 - SPLIT_FRACTION) * root->UpperRight.Y;
  else mid = SPLIT_FRACTION * root->LowerLeft.X + (1 - SPLIT_FRACTION) * root->UpperRight.X;

  int64_t na = partition(root->Particles, root->Len, orientation, mid);
  int64_t nb = root->Len - na;
  int other = orientation == Vertical ? Horizontal : Vertical;

  if (na > 0) {
    Cell* c = arena_new(s);
    c->Particles = root->Particles; c->Len = na;
    c->LowerLeft = root->LowerLeft; c->UpperRight = root->UpperRight;
    if (orientation == Vertical) c->UpperRight.Y = mid; else c->UpperRight.X = mid;
    root->Lower = c;
    if (na > MAX_PARTICLES_PER_CELL) treebuild(s, c, other);
  }
  if (nb > 0) {
    Cell* c = arena_new(s);
    c->Particles = root->Particles + na; c->Len = nb;
    c->LowerLeft = root->LowerLeft; c->UpperRight = root->UpperRight;
    if (orientation == Vertical) c->LowerLeft.Y = mid; else c->LowerLeft.X = mid;
    root->Upper = c;
    if (nb > MAX_PARTICLES_PER_CELL) treebuild(s, c, other);
  }
}

/* ---------- BoundingSpheres, core.go:229-312 ---------- */
static void bounding_spheres(Cell* root) {
  if (!root->Upper && !root->Lower) {
    if (root->Len == 1) { root->BCenter = root->Particles[0].Pos; root->BRadius = 0; }
    else {
      double dSquaredMax = 0.0; Vec2 pA = {0, 0}, pB = {0, 0};
      for (int64_t a = 0; a < root->Len; a++)
        for (int64_t b = 0; b < root->Len; b++) {
          Vec2 p1 = root->Particles[a].Pos, p2 = root->Particles[b].Pos;
          double x = p2.X - p1.X, y = p2.Y - p1.Y;
          double dSq = x * x + y * y;
          if (dSq > dSquaredMax) { dSquaredMax = dSq; pA = p1; pB = p2; }
        }
      Vec2 rMax = {(pB.X - pA.X) * 0.5, (pB.Y - pA.Y) * 0.5};
      root->BCenter.X = rMax.X + pA.X; root->BCenter.Y = rMax.Y + pA.Y;
      double BRadiusSq = rMax.X * rMax.X + rMax.Y * rMax.Y;
      for (int64_t a = 0; a < root->Len; a++) {
        double x = root->BCenter.X - root->Particles[a].Pos.X, y = root->BCenter.Y - root->Particles[a].Pos.Y;
        double rNewSq = x * x + y * y;
        if (rNewSq > BRadiusSq) BRadiusSq = rNewSq;
      }
      root->BRadius = sqrt(BRadiusSq);
    }
    return;
  }
  if (root->Upper) bounding_spheres(root->Upper);
  if (root->Lower) bounding_spheres(root->Lower);
  if (!root->Upper) { root->BRadius = root->Lower->BRadius; root->BCenter = root->Lower->BCenter; return; }
  if (!root->Lower) { root->BRadius = root->Upper->BRadius; root->BCenter = root->Upper->BCenter; return; }
  /* circle enclosing the two child circles, core.go:300-311 (non-enclosing on containment: kept as is) */
  Vec2 AB = {root->Upper->BCenter.X - root->Lower->BCenter.X, root->Upper->BCenter.Y - root->Lower->BCenter.Y};
  double ABNorm = sqrt(AB.X * AB.X + AB.Y * AB.Y);
  double rA = root->Lower->BRadius, rB = root->Upper->BRadius;
  double rC = (rA + rB + ABNorm) * 0.5;
  double f = (rB - rC) / ABNorm;
  root->BCenter.X = AB.X * f + root->Upper->BCenter.X;
  root->BCenter.Y = AB.Y * f + root->Upper->BCenter.Y;
  root->BRadius = rC;
}

/* ---------- top-32 queue, nearest-neighbour.go:131-165 ---------- */
static int64_t g_inserts = 0; /* bookkeeping only: lets orc_knn count queues that kept a sentinel */
static inline void NNQueueInsert(Particle* p, double dist, Particle* nb, Vec2 realPos) {
  int i = 1;
  g_inserts++;
  for (; i < NN_SIZE && p->NNDists[i] > dist; i++) {
    p->NNDists[i - 1] = p->NNDists[i];
    p->NearestNeighbours[i - 1] = p->NearestNeighbours[i];
    p->NNPos[i - 1] = p->NNPos[i];
  }
  p->NNDists[i - 1] = dist; p->NearestNeighbours[i - 1] = nb; p->NNPos[i - 1] = realPos;
}
static inline void NNQueueInitSentinel(Particle* p) { for (int i = 0; i < NN_SIZE; i++) p->NNDists[i] += 0.4; }

/* ---------- findNNRec, nearest-neighbour.go:70-119 ---------- */
static void findNNRec(Particle* particle, Cell* root, Vec2 offset) {
  Vec2 pos = {particle->Pos.X + offset.X, particle->Pos.Y + offset.Y};
  if (!root->Upper && !root->Lower) {
    for (int64_t i = 0; i < root->Len; i++) {
      Particle* b = &root->Particles[i];
      double d2 = DistSq(pos, b->Pos);
      if (d2 < particle->NNDists[0] && particle != b) {
        Vec2 rp = {b->Pos.X - offset.X, b->Pos.Y - offset.Y};
        NNQueueInsert(particle, d2, b, rp);
      }
    }
    return;
  }
  if (root->Upper && root->Lower) {
    double distUpper = Dist(root->Upper->BCenter, pos);
    double distLower = Dist(root->Lower->BCenter, pos);
    double maxDist = sqrt(particle->NNDists[0]); /* not refreshed between the two children */
    if (distLower < distUpper) {
      if (distLower - root->Lower->BRadius < maxDist) findNNRec(particle, root->Lower, offset);
      if (distUpper - root->Upper->BRadius < maxDist) findNNRec(particle, root->Upper, offset);
    } else {
      if (distUpper - root->Upper->BRadius < maxDist) findNNRec(particle, root->Upper, offset);
      if (distLower - root->Lower->BRadius < maxDist) findNNRec(particle, root->Lower, offset);
    }
    return;
  }
  if (root->Upper) findNNRec(particle, root->Upper, offset);
  if (root->Lower) findNNRec(particle, root->Lower, offset);
}

/* image loop shared by the faithful and the exact mode, nearest-neighbour.go:28-55 */
static int image_ranges(const double hor[2], const double ver[2], int* iS, int* iE, int* jS, int* jE,
                        double* dX, double* dY) {
  *iS = -1; *jS = -1; *iE = 1; *jE = 1;
  *dX = hor[1] - hor[0]; *dY = ver[1] - ver[0];
  if (hor[0] == -DBL_MAX) { *iS = 0; *iE = 0; *dX = 0; if (hor[1] != DBL_MAX) return ORC_PANIC_HALF_OPEN; }
  if (ver[0] == -DBL_MAX) { *jS = 0; *jE = 0; *dY = 0; if (ver[1] != DBL_MAX) return ORC_PANIC_HALF_OPEN; }
  return ORC_OK;
}

/* FindNearestNeighboursPeriodic, nearest-neighbour.go:28-67 (FindNearestNeighbours :15-23 is the
 * special case of two open axes: one image, offset 0) */
static int find_nn_periodic(Particle* p, Cell* root, const double hor[2], const double ver[2]) {
  NNQueueInitSentinel(p);
  int iS, iE, jS, jE; double dX, dY;
  int rc = image_ranges(hor, ver, &iS, &iE, &jS, &jE, &dX, &dY);
  if (rc) return rc;
  for (int i = iS; i <= iE; i++)
    for (int j = jS; j <= jE; j++) {
      Vec2 off = {(double)i * dX, (double)j * dY};
      findNNRec(p, root, off);
    }
  for (int i = 0; i < NN_SIZE; i++) p->NNDists[i] = sqrt(p->NNDists[i]);
  return ORC_OK;
}

/* exact mode: brute force over every (particle, image) pair, no sentinel cap; same queue, same
 * image order, same self-exclusion by identity.  Independent of the tree. */
static int find_nn_exact(Particle* p, Particle* ps, int64_t n, const double hor[2], const double ver[2]) {
  for (int i = 0; i < NN_SIZE; i++) { p->NNDists[i] = DBL_MAX; p->NearestNeighbours[i] = NULL; p->NNPos[i].X = p->NNPos[i].Y = 0; }
  int iS, iE, jS, jE; double dX, dY;
  int rc = image_ranges(hor, ver, &iS, &iE, &jS, &jE, &dX, &dY);
  if (rc) return rc;
  for (int i = iS; i <= iE; i++)
    for (int j = jS; j <= jE; j++) {
      Vec2 off = {(double)i * dX, (double)j * dY};
      Vec2 pos = {p->Pos.X + off.X, p->Pos.Y + off.Y};
      for (int64_t b = 0; b < n; b++) {
        double d2 = DistSq(pos, ps[b].Pos);
        if (d2 < p->NNDists[0] && p != &ps[b]) {
          Vec2 rp = {ps[b].Pos.X - off.X, ps[b].Pos.Y - off.Y};
          NNQueueInsert(p, d2, &ps[b], rp);
        }
      }
    }
  for (int i = 0; i < NN_SIZE; i++) p->NNDists[i] = sqrt(p->NNDists[i]);
  return ORC_OK;
}

/* ---------- kernels, sph.go:237-304.  Prefactors are Go untyped-constant expressions, evaluated
 * exactly and rounded once; the hex literals are those roundings (checked in tests/test_oracle.py). */
static const double MONAGHAN_PREF = 0x1.5d3b3e3583243p+3; /* 6*40/(pi*7)  = 10.913... */
static const double WENDLAND_FPREF = 0x1.1d34a60108f72p+1; /* 4*7/(pi*4) = 2.2281... */
static const double WENDLAND_DFPREF = 0x1.1d34a60108f72p+2;/* 8*7/(pi*4) = 4.4563... */
static const double TOPHAT_FPREF = 0x1.45f306dc9c883p-2;   /* 1/pi */

static inline double kernel_F(int k, double q) {
  switch (k) {
    case 0: return 1;
    case 1: if (q < 0.5) return q * q * q - q * q + 1.0 / 6; return (1 - q) * (1 - q) * (1 - q) / 3;
    default: return (1 - q) * (1 - q) * (1 - q) * (1 - q) * (1 + 4 * q);
  }
}
static inline double kernel_DF(int k, double q) {
  if (k == 1) { if (q < 0.5) return (3 * q * q - 2 * q); return -(1 - q) * (1 - q); }
  return -10 * q * (1 - q) * (1 - q) * (1 - q);
}
static inline double kernel_FPref(int k) { return k == 0 ? TOPHAT_FPREF : k == 1 ? MONAGHAN_PREF : WENDLAND_FPREF; }
static inline double kernel_DFPref(int k) { return k == 0 ? 1.0 : k == 1 ? MONAGHAN_PREF : WENDLAND_DFPREF; }

/* Density2D, sph.go:306-323 */
static int density2d(const Particle* p, const orc_params* cfg, int kernel, double* out) {
  double maxR = p->NNDists[0], acc = 0.0;
  for (int i = 0; i < NN_SIZE; i++) {
    double x = p->NNDists[i] / maxR;
    if (x > 1 || x < 0) return ORC_PANIC_Q_RANGE;
    acc += kernel_F(kernel, x);
  }
  *out = kernel_FPref(kernel) * cfg->particle_mass * acc / (maxR * maxR);
  return ORC_OK;
}

/* AccelerationAndEDot2D, sph.go:327-401 */
static int acceleration_and_edot(Particle* p, const orc_params* cfg) {
  const int kernel = cfg->kernel;
  double gamma = cfg->gamma, maxR = p->NNDists[0];
  double contributionA = p->C * p->C / (gamma * p->Rho);
  double acc_ax = 0.0, acc_ay = 0.0, acc_edot = 0.0;
  for (int i = 0; i < NN_SIZE; i++) {
    Particle* nn = p->NearestNeighbours[i];
    if (!nn) break; /* sph.go:347-349 */
    double q = p->NNDists[i] / maxR;
    if (q > 1 || q < 0) return ORC_PANIC_Q_RANGE;
    if (kernel == 0) return ORC_PANIC_TOPHAT_DF;
    double dRKernel = kernel_DF(kernel, q);
    double contributionB = nn->C * nn->C / (gamma * nn->Rho);
    Vec2 vA = p->VPred, vB = nn->VPred, rA = p->Pos, rB = p->NNPos[i];
    Vec2 vAB = {vB.X - vA.X, vB.Y - vA.Y}, rAB = {rB.X - rA.X, rB.Y - rA.Y};
    double dot = vAB.X * rAB.X + vAB.Y * rAB.Y;
    double piAB = 0.0;
    if (dot < 0) {
      const double alpha = 0.75, beta = 1.5, etaSq = 0.01;
      double cAB = 0.5 * (p->C + nn->C);
      double rhoAB = 0.5 * (p->Rho + nn->Rho);
      double hAB = 0.5 * (p->NNDists[0] + nn->NNDists[0]);
      double muAB = dot * hAB / ((rAB.X * rAB.X + rAB.Y * rAB.Y) + etaSq);
      piAB = (-alpha * cAB * muAB + beta * muAB * muAB) / rhoAB;
    }
    acc_ax += rAB.X * (piAB + contributionA + contributionB) * dRKernel / p->NNDists[i];
    acc_ay += rAB.Y * (piAB + contributionA + contributionB) * dRKernel / p->NNDists[i];
    acc_edot += dot * dRKernel;
  }
  double f = cfg->particle_mass * kernel_DFPref(kernel) / (maxR * maxR * maxR);
  p->VDot.X = acc_ax * f + cfg->accel[0];
  p->VDot.Y = acc_ay * f + cfg->accel[1];
  p->EDot = contributionA * acc_edot * cfg->particle_mass;
  return ORC_OK;
}

/* MakeCells, core.go:93-105: fresh root box [0,1]^2 */
static void make_cells(orc_sim* s) {
  arena_reset(s);
  memset(&s->root, 0, sizeof(Cell));
  s->root.UpperRight.X = 1; s->root.UpperRight.Y = 1;
  s->root.Particles = s->ps; s->root.Len = s->n;
  treebuild(s, &s->root, Vertical);
  bounding_spheres(&s->root);
}

static void rebuild_tree(orc_sim* s) { /* sph.go:406-408 on the persistent root */
  Cell *oldL = s->root.Lower, *oldU = s->root.Upper;
  arena_reset(s);
  s->root.Lower = NULL; s->root.Upper = NULL;
  s->root.Particles = s->ps; s->root.Len = s->n;
  treebuild(s, &s->root, Vertical);
  /* quirk 8: the reference would keep the old child where the new half is empty */
  if ((oldL && !s->root.Lower) || (oldU && !s->root.Upper)) s->stale_root_child = 1;
  bounding_spheres(&s->root);
}

static void note(orc_sim* s, int rc) { if (rc && !s->status) s->status = rc; }


/* ---------- public API (ctypes) ---------- */
orc_sim* orc_create(const orc_params* cfg, int64_t n, int64_t cap, const double* pos, const double* vel,
                    const double* e, const double* rho, const int64_t* id) {
  orc_sim* s = (orc_sim*)calloc(1, sizeof(orc_sim));
  if (!s) return NULL;
  if (cap < n) cap = n;
  s->cfg = *cfg; s->n = n; s->cap = cap;
  s->ps = (Particle*)calloc((size_t)(cap > 0 ? cap : 1), sizeof(Particle));
  if (!s->ps) { free(s); return NULL; }
  for (int64_t i = 0; i < n; i++) {
    Particle* p = &s->ps[i];
    p->Pos.X = pos[2 * i]; p->Pos.Y = pos[2 * i + 1];
    if (vel) { p->Vel.X = vel[2 * i]; p->Vel.Y = vel[2 * i + 1]; }
    if (e) p->E = e[i];
    if (rho) p->Rho = rho[i];
    p->Z = id ? id[i] : i;
  }
  make_cells(s);
  return s;
}

void orc_destroy(orc_sim* s) { if (!s) return; arena_reset(s); free(s->ps); free(s); }
void orc_set_params(orc_sim* s, const orc_params* cfg) { s->cfg = *cfg; }
int64_t orc_count(const orc_sim* s) { return s->n; }
int64_t orc_current_step(const orc_sim* s) { return s->current_step; }
int orc_status(const orc_sim* s) { return s->status; }
int orc_stale_root_child(const orc_sim* s) { return s->stale_root_child; }
int64_t orc_underfull(const orc_sim* s) { return s->underfull; }
int64_t orc_sizeof_particle(void) { return (int64_t)sizeof(Particle); }

/* append + MakeCells, sph.go:75-86 */
int orc_append(orc_sim* s, int64_t n, const double* pos, const double* vel, const double* e,
               const double* rho, const int64_t* id) {
  if (s->n + n > s->cap) return ORC_E_NOMEM; /* Go would reallocate; pointers are rebuilt anyway */
  for (int64_t i = 0; i < n; i++) {
    Particle* p = &s->ps[s->n + i];
    memset(p, 0, sizeof(Particle));
    p->Pos.X = pos[2 * i]; p->Pos.Y = pos[2 * i + 1];
    if (vel) { p->Vel.X = vel[2 * i]; p->Vel.Y = vel[2 * i + 1]; }
    if (e) p->E = e[i];
    if (rho) p->Rho = rho[i];
    p->Z = id ? id[i] : s->n + i;
  }
  s->n += n;
  make_cells(s);
  return ORC_OK;
}

/* batch kNN. mode 0 = faithful (tree walk + "+0.4" sentinel), mode 1 = exact brute force.
 * rebuild != 0 first does Treebuild + BoundingSpheres on the persistent root like CalculateForces. */
int orc_knn(orc_sim* s, const double hor[2], const double ver[2], int mode, int rebuild) {
  if (rebuild) rebuild_tree(s);
  s->underfull = 0;
  for (int64_t i = 0; i < s->n; i++) {
    int64_t before = g_inserts;
    int rc = mode == 0 ? find_nn_periodic(&s->ps[i], &s->root, hor, ver)
                       : find_nn_exact(&s->ps[i], s->ps, s->n, hor, ver);
    if (rc) { note(s, rc); return rc; }
    if (g_inserts - before < NN_SIZE) s->underfull++; /* >= 1 slot still holds a sentinel */
  }
  return ORC_OK;
}

int orc_density(orc_sim* s, int kernel) {
  for (int64_t i = 0; i < s->n; i++) {
    int rc = density2d(&s->ps[i], &s->cfg, kernel, &s->ps[i].Rho);
    if (rc) { note(s, rc); return rc; }
  }
  return ORC_OK;
}

/* CalculateForces, sph.go:403-435. knn_mode 0 = faithful to the reference. */
int orc_calc_forces_mode(orc_sim* s, int knn_mode) {
  int rc = orc_knn(s, s->cfg.hor, s->cfg.ver, knn_mode, 1);
  if (rc) return rc;
  rc = orc_density(s, s->cfg.kernel);
  if (rc) return rc;
  double factor = s->cfg.gamma * (s->cfg.gamma - 1);
  for (int64_t i = 0; i < s->n; i++) s->ps[i].C = sqrt(factor * s->ps[i].EPred);
  for (int64_t i = 0; i < s->n; i++) {
    rc = acceleration_and_edot(&s->ps[i], &s->cfg);
    if (rc) { note(s, rc); return rc; }
  }
  return ORC_OK;
}
int orc_calc_forces(orc_sim* s) { return orc_calc_forces_mode(s, 0); }

/* Step, sph.go:64-198 (sources are fed from outside through orc_append before the call) */
int orc_step_mode(orc_sim* s, int knn_mode) {
  const orc_params* c = &s->cfg;
  double dtHalf = c->dt_half;
  int rc;
  if (s->current_step == 0) {
    if (s->n == 0) { note(s, ORC_PANIC_NOT_INIT); return ORC_PANIC_NOT_INIT; }
    for (int64_t i = 0; i < s->n; i++) { s->ps[i].VPred = s->ps[i].Vel; s->ps[i].EPred = s->ps[i].E; }
    rc = orc_calc_forces_mode(s, knn_mode);
    if (rc) return rc;
  }
  for (int64_t i = 0; i < s->n; i++) { /* drift 1 + predict, sph.go:108-117 */
    Particle* p = &s->ps[i];
    p->Pos.X = p->Pos.X + p->Vel.X * dtHalf; p->Pos.Y = p->Pos.Y + p->Vel.Y * dtHalf;
    p->VPred.X = p->Vel.X + p->VDot.X * dtHalf; p->VPred.Y = p->Vel.Y + p->VDot.Y * dtHalf;
    p->EPred = p->E + p->EDot * dtHalf;
  }
  rc = orc_calc_forces_mode(s, knn_mode);
  if (rc) return rc;
  for (int64_t i = 0; i < s->n; i++) { /* kick, sph.go:122-127 */
    Particle* p = &s->ps[i];
    double f = 2 * dtHalf;
    p->Vel.X = p->Vel.X + p->VDot.X * f; p->Vel.Y = p->Vel.Y + p->VDot.Y * f;
    p->E = p->E + p->EDot * 2 * dtHalf;
  }
  for (int64_t i = 0; i < s->n; i++) { /* drift 2, sph.go:130-135 */
    Particle* p = &s->ps[i];
    p->Pos.X = p->Pos.X + p->Vel.X * dtHalf; p->Pos.Y = p->Pos.Y + p->Vel.Y * dtHalf;
  }
  for (int64_t i = 0; i < s->n; i++) { /* periodic wrap with the `continue` quirk, sph.go:147-167 */
    Particle* p = &s->ps[i];
    if (p->Pos.X < c->hor[0]) { p->Pos.X += (c->hor[1] - c->hor[0]); continue; }
    if (p->Pos.X > c->hor[1]) { p->Pos.X -= (c->hor[1] - c->hor[0]); continue; }
    if (p->Pos.Y < c->ver[0]) { p->Pos.Y += (c->ver[1] - c->ver[0]); continue; }
    if (p->Pos.Y > c->ver[1]) { p->Pos.Y -= (c->ver[1] - c->ver[0]); }
  }
  for (int64_t i = 0; i < s->n; i++) { /* reflections, sph.go:170-193 */
    Particle* p = &s->ps[i];
    if (p->Pos.X < c->refl_L) { p->Pos.X -= p->Pos.X - c->refl_L; p->Vel.X = -p->Vel.X; }
    if (p->Pos.X > c->refl_R) { p->Pos.X -= p->Pos.X - c->refl_R; p->Vel.X = -p->Vel.X; }
    if (p->Pos.Y < c->refl_U) { p->Pos.Y -= p->Pos.Y - c->refl_U; p->Vel.Y = -p->Vel.Y; }
    if (p->Pos.Y > c->refl_D) { p->Pos.Y -= p->Pos.Y - c->refl_D; p->Vel.Y = -p->Vel.Y; }
  }
  s->current_step += 1;
  return ORC_OK;
}
int orc_step(orc_sim* s) { return orc_step_mode(s, 0); }
int orc_run(orc_sim* s, int nsteps, int knn_mode) {
  for (int i = 0; i < nsteps; i++) { int rc = orc_step_mode(s, knn_mode); if (rc) return rc; }
  return ORC_OK;
}

/* reductions, sph.go:441-463 */
double orc_total_energy(const orc_sim* s) { double t = 0; for (int64_t i = 0; i < s->n; i++) t += s->ps[i].E; return t; }
double orc_total_density(const orc_sim* s) { double t = 0; for (int64_t i = 0; i < s->n; i++) t += s->ps[i].Rho; return t; }
double orc_total_momentum(const orc_sim* s) { /* `=` not `+=`, sph.go:460 */
  double t = 0;
  for (int64_t i = 0; i < s->n; i++) t = sqrt(s->ps[i].Vel.X * s->ps[i].Vel.X + s->ps[i].Vel.Y * s->ps[i].Vel.Y);
  return t;
}

/* state extraction in the current (tree-permuted) order; any pointer may be NULL.
 * nn_id[k] = Z of NearestNeighbours[k] or -1 for nil. */
void orc_get(const orc_sim* s, double* pos, double* vel, double* rho, double* c, double* e, double* edot,
             double* vdot, double* epred, double* vpred, double* h, int64_t* id, int64_t* nn_id,
             double* nn_dist, double* nn_pos) {
  for (int64_t i = 0; i < s->n; i++) {
    const Particle* p = &s->ps[i];
    if (pos) { pos[2 * i] = p->Pos.X; pos[2 * i + 1] = p->Pos.Y; }
    if (vel) { vel[2 * i] = p->Vel.X; vel[2 * i + 1] = p->Vel.Y; }
    if (rho) rho[i] = p->Rho;
    if (c) c[i] = p->C;
    if (e) e[i] = p->E;
    if (edot) edot[i] = p->EDot;
    if (vdot) { vdot[2 * i] = p->VDot.X; vdot[2 * i + 1] = p->VDot.Y; }
    if (epred) epred[i] = p->EPred;
    if (vpred) { vpred[2 * i] = p->VPred.X; vpred[2 * i + 1] = p->VPred.Y; }
    if (h) h[i] = p->NNDists[0];
    if (id) id[i] = p->Z;
    for (int k = 0; k < NN_SIZE; k++) {
      if (nn_id) nn_id[i * NN_SIZE + k] = p->NearestNeighbours[k] ? p->NearestNeighbours[k]->Z : -1;
      if (nn_dist) nn_dist[i * NN_SIZE + k] = p->NNDists[k];
      if (nn_pos) { nn_pos[(i * NN_SIZE + k) * 2] = p->NNPos[k].X; nn_pos[(i * NN_SIZE + k) * 2 + 1] = p->NNPos[k].Y; }
    }
  }
}

/* overwrite Rho / E / Vel etc. in current order (examples poke Root.Particles directly) */
void orc_set(orc_sim* s, const double* pos, const double* vel, const double* rho, const double* e) {
  for (int64_t i = 0; i < s->n; i++) {
    Particle* p = &s->ps[i];
    if (pos) { p->Pos.X = pos[2 * i]; p->Pos.Y = pos[2 * i + 1]; }
    if (vel) { p->Vel.X = vel[2 * i]; p->Vel.Y = vel[2 * i + 1]; }
    if (rho) p->Rho = rho[i];
    if (e) p->E = e[i];
  }
}

/* ----- KAT helpers ----- */
/* Partition on bare positions; returns len(a). partition_test.go:18-160 */
int64_t orc_partition(double* pos_xy, int64_t n, int orientation, double middle) {
  Particle* ps = (Particle*)calloc((size_t)(n > 0 ? n : 1), sizeof(Particle));
  for (int64_t i = 0; i < n; i++) { ps[i].Pos.X = pos_xy[2 * i]; ps[i].Pos.Y = pos_xy[2 * i + 1]; }
  int64_t a = partition(ps, n, orientation, middle);
  for (int64_t i = 0; i < n; i++) { pos_xy[2 * i] = ps[i].Pos.X; pos_xy[2 * i + 1] = ps[i].Pos.Y; }
  free(ps);
  return a;
}

/* isInsideAny, bounding-sphere_test.go:7-28 */
static int inside_any(Vec2 pos, const Cell* cell) {
  double x = pos.X - cell->BCenter.X, y = pos.Y - cell->BCenter.Y;
  if (sqrt(x * x + y * y) <= cell->BRadius) return 1;
  if (cell->Upper && inside_any(pos, cell->Upper)) return 1;
  if (cell->Lower && inside_any(pos, cell->Lower)) return 1;
  return 0;
}
/* returns the number of particles NOT inside any node circle (the reference tests expect 0) */
int64_t orc_tree_count_outside_all(const orc_sim* s) {
  int64_t bad = 0;
  for (int64_t i = 0; i < s->n; i++) if (!inside_any(s->ps[i].Pos, &s->root)) bad++;
  return bad;
}
static void tree_stats(const Cell* c, int depth, int64_t* nodes, int64_t* leaves, int64_t* maxleaf, int* maxdepth) {
  (*nodes)++;
  if (depth > *maxdepth) *maxdepth = depth;
  if (!c->Upper && !c->Lower) { (*leaves)++; if (c->Len > *maxleaf) *maxleaf = c->Len; return; }
  if (c->Upper) tree_stats(c->Upper, depth + 1, nodes, leaves, maxleaf, maxdepth);
  if (c->Lower) tree_stats(c->Lower, depth + 1, nodes, leaves, maxleaf, maxdepth);
}
void orc_tree_stats(const orc_sim* s, int64_t* out4) {
  int64_t nodes = 0, leaves = 0, maxleaf = 0; int maxdepth = 0;
  tree_stats(&s->root, 1, &nodes, &leaves, &maxleaf, &maxdepth);
  out4[0] = nodes; out4[1] = leaves; out4[2] = maxleaf; out4[3] = maxdepth;
}

/* the tree in pre-order for the tests that redraw the reference's pictures (visualization.go:33-75): per node
 * geo = {LowerLeft.X, .Y, UpperRight.X, .Y, BCenter.X, .Y, BRadius}, link = {index of Lower, index of Upper (-1 = nil),
 * first particle (offset in the current order), particle count}.  Returns the node count (nothing is written past cap). */
static int64_t tree_dump(const Cell* c, const Particle* base, int64_t cap, double* geo, int64_t* link, int64_t* next) {
  const int64_t me = (*next)++;
  int64_t lo = -1, up = -1;
  if (c->Lower) lo = tree_dump(c->Lower, base, cap, geo, link, next);
  if (c->Upper) up = tree_dump(c->Upper, base, cap, geo, link, next);
  if (me < cap) {
    double* g = geo + 7 * me;
    g[0] = c->LowerLeft.X; g[1] = c->LowerLeft.Y; g[2] = c->UpperRight.X; g[3] = c->UpperRight.Y;
    g[4] = c->BCenter.X; g[5] = c->BCenter.Y; g[6] = c->BRadius;
    int64_t* l = link + 4 * me;
    l[0] = lo; l[1] = up; l[2] = (int64_t)(c->Particles - base); l[3] = c->Len;
  }
  return me;
}
int64_t orc_tree_dump(const orc_sim* s, int64_t cap, double* geo, int64_t* link) {
  int64_t next = 0;
  tree_dump(&s->root, s->ps, cap, geo, link, &next);
  return next;
}

/* generic min-heap on int64, heap.go:51-146 (not on the executed path; README KAT only) */
static void heapify(int64_t* a, int64_t len, int64_t i) {
  for (;;) {
    int64_t l = i * 2 + 1, r = i * 2 + 2, m = i;
    if (l < len && a[l] < a[m]) m = l;
    if (r < len && a[r] < a[m]) m = r;
    if (m == i) break;
    int64_t t = a[m]; a[m] = a[i]; a[i] = t;
    i = m;
  }
}
void orc_heap_build(int64_t* a, int64_t len) { if (len < 2) return; for (int64_t i = len / 2 - 1; i >= 0; i--) heapify(a, len, i); }
int64_t orc_heap_insert(int64_t* a, int64_t len, int64_t element) { /* a has room for len+1 */
  a[len] = element; len++;
  int64_t index = len - 1, parent = len / 2 - 1;
  while (parent >= 0 && a[index] < a[parent]) {
    int64_t t = a[parent]; a[parent] = a[index]; a[index] = t;
    index = parent; parent = (parent + 1) / 2 - 1;
  }
  return len;
}
int64_t orc_heap_extract_min(int64_t* a, int64_t len, int64_t* min_out) {
  if (len == 0) return -1;
  *min_out = a[0]; a[0] = a[len - 1]; len--;
  heapify(a, len, 0);
  return len;
}
int64_t orc_heap_replace(int64_t* a, int64_t len, int64_t element, int64_t* min_out) {
  if (len == 0) return -1;
  *min_out = a[0]; a[0] = element;
  heapify(a, len, 0);
  return len;
}
