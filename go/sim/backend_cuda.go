//go:build cuda

// backend_cuda.go — cgo shim that puts libsphb.so (include/sphb.h) behind the exported API of package sim.
//
// Drop this file into github.com/bbeni/sphugo/sim and build with `-tags cuda` (INTEGRATION.md lists the
// three-line patch that moves the CPU bodies of Step / CalculateForces / Total* behind `//go:build !cuda`).
// simviewer and the examples keep compiling unchanged: Simulation, SphConfig, Root.Particles and every
// method signature stay as they are (sim/sph.go:14-64, 403, 441-463).
//
// NOTE: written against Go 1.22 (go.mod:3) but NOT compiled in the build container of this repository (no Go
// toolchain there); the same C ABI is exercised by the Python ctypes binding (sphugo_b200/_lib.py) in the tests.
//
// Data flow: particle state lives on the GPU between steps.  Root.Particles (AoS, 1136 B per particle,
// sim/core.go:17-42) is refreshed lazily by Sync(), which the animator calls before reading particles
// (sim/animator.go:60) — downloading 1.1 KB per particle after every step would dwarf the step itself.
package sim

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../sphugo_b200 -lsphb -Wl,-rpath,${SRCDIR}/../../sphugo_b200
#include <stdlib.h>
#include "sphb.h"
*/
import "C"

import (
	"fmt"
	"math"
	"os"
	"runtime"
	"sync"
	"unsafe"
)

// gpuBackend is stored per Simulation (keyed by pointer: Simulation is copied by value in simviewer.go:147,224,
// so the handle must not live inside the struct that gets copied while a step is running).
type gpuBackend struct {
	h         *C.sphb_sim
	n         int
	hostStale bool  // Root.Particles is older than the device state
	uploaded  int   // number of particles the device knows about (Sources append, sph.go:75-86)
	root      *Cell // the Root this device copy was made from: a different Root = the Simulation value was replaced
}

var (
	backends   = map[*Simulation]*gpuBackend{}
	backendsMu sync.Mutex // simviewer steps on one goroutine and replaces / closes simulations on another (simviewer.go:218-237)
)

func lookup(sim *Simulation) *gpuBackend {
	backendsMu.Lock()
	defer backendsMu.Unlock()
	return backends[sim]
}

// kernelID identifies a Kernel by value: closures cannot cross the C ABI (sph.go:237-242).
func kernelID(k Kernel) C.int32_t {
	switch k.FPrefactor {
	case TopHat2D.FPrefactor:
		return C.SPHB_KERNEL_TOPHAT
	case Monahan2D.FPrefactor:
		return C.SPHB_KERNEL_MONAGHAN
	case Wendtland2D.FPrefactor:
		return C.SPHB_KERNEL_WENDLAND
	}
	panic("unknown kernel")
}

func paramsOf(c *SphConfig) C.sphb_params {
	var p C.sphb_params
	p.dt_half = C.double(c.DeltaTHalf)
	p.gamma = C.double(c.Gamma)
	p.particle_mass = C.double(c.ParticleMass)
	p.accel[0], p.accel[1] = C.double(c.Acceleration.X), C.double(c.Acceleration.Y)
	p.hor[0], p.hor[1] = C.double(c.HorPeriodicity[0]), C.double(c.HorPeriodicity[1])
	p.ver[0], p.ver[1] = C.double(c.VertPeriodicity[0]), C.double(c.VertPeriodicity[1])
	p.refl_L, p.refl_R = C.double(c.Reflections.L), C.double(c.Reflections.R)
	p.refl_U, p.refl_D = C.double(c.Reflections.U), C.double(c.Reflections.D)
	p.kernel = kernelID(c.Kernel)
	p.precision = 64 // SPHB_PRECISION=32 selects the fp32 build (results within 1e-5 instead of 1e-12)
	if os.Getenv("SPHB_PRECISION") == "32" {
		p.precision = 32
	}
	p.device = 0
	return p
}

func check(b *gpuBackend, rc C.int) {
	if rc == C.SPHB_OK {
		return
	}
	var h *C.sphb_sim
	if b != nil {
		h = b.h
	}
	// the reference panics on these conditions too (sph.go:93,251,317,354; nearest-neighbour.go:44,53)
	panic(fmt.Sprintf("libsphb %d: %s", int(rc), C.GoString(C.sphb_last_error(h))))
}

// soa flattens ps (= Root.Particles[base:]) for sphb_create / sphb_append.  The device id of a particle is its index
// in Root.Particles: with this backend the slice is never permuted (no Treebuild on the host), and Particle.Z is not a
// safe key because every Spawn re-seeds math/rand, so two rectangles can repeat a Z (config-parser.go:60-64,75).
func soa(ps []Particle, base int) (pos, vel, e, rho []float64, id []int64) {
	n := len(ps)
	pos, vel = make([]float64, 2*n), make([]float64, 2*n)
	e, rho, id = make([]float64, n), make([]float64, n), make([]int64, n)
	for i := range ps {
		p := &ps[i]
		pos[2*i], pos[2*i+1] = p.Pos.X, p.Pos.Y
		vel[2*i], vel[2*i+1] = p.Vel.X, p.Vel.Y
		e[i], rho[i], id[i] = p.E, p.Rho, int64(base+i)
	}
	return
}

func dptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}
func iptr(s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}

// backend creates the device copy on first use (== MakeCells, core.go:93-105) and appends particles that were
// added to Root.Particles since (Sources).
func (sim *Simulation) backend() *gpuBackend {
	b := lookup(sim)
	ps := sim.Root.Particles
	if b != nil && b.root != sim.Root {
		// simviewer assigns a new Simulation value to the same variable (simviewer.go:147,224): same key, new particles
		backendsMu.Lock()
		delete(backends, sim)
		backendsMu.Unlock()
		C.sphb_destroy(b.h)
		b = nil
	}
	if b == nil {
		b = &gpuBackend{root: sim.Root}
		prm := paramsOf(&sim.Config)
		pos, vel, e, rho, id := soa(ps, 0)
		capacity := len(ps) + 100000 // sph.go:45 reserves 100000 as well
		check(nil, C.sphb_create(&prm, C.int64_t(len(ps)), C.int64_t(capacity), dptr(pos), dptr(vel), dptr(e), dptr(rho), iptr(id), &b.h))
		b.uploaded = len(ps)
		if sim.CurrentStep > 0 && len(ps) > 0 {
			// a device copy made in the middle of a run (simviewer replaced the Simulation value, or the first GPU call
			// comes after CPU steps): carry the derivatives the predictor reads (sph.go:108-117) and the step counter
			// over, so that the next Step() does not repeat the step-0 initialisation (sph.go:89-103)
			vdot := make([]float64, 2*len(ps))
			edot := make([]float64, len(ps))
			for i := range ps {
				vdot[2*i], vdot[2*i+1] = ps[i].VDot.X, ps[i].VDot.Y
				edot[i] = ps[i].EDot
			}
			var ptrs [C.SPHB_F_COUNT]unsafe.Pointer
			ptrs[C.SPHB_F_VDOT] = unsafe.Pointer(&vdot[0])
			ptrs[C.SPHB_F_EDOT] = unsafe.Pointer(&edot[0])
			mask := C.uint32_t(1<<C.SPHB_F_VDOT | 1<<C.SPHB_F_EDOT)
			check(b, C.sphb_upload(b.h, mask, (*unsafe.Pointer)(unsafe.Pointer(&ptrs[0])), C.int64_t(len(ps))))
			check(b, C.sphb_set_current_step(b.h, C.int64_t(sim.CurrentStep)))
		}
		backendsMu.Lock()
		backends[sim] = b
		backendsMu.Unlock()
	} else if len(ps) > b.uploaded {
		pos, vel, e, rho, id := soa(ps[b.uploaded:], b.uploaded)
		check(b, C.sphb_append(b.h, C.int64_t(len(id)), dptr(pos), dptr(vel), dptr(e), dptr(rho), iptr(id)))
		b.uploaded = len(ps)
	}
	prm := paramsOf(&sim.Config) // Config is a public mutable field (sph.go:15)
	check(b, C.sphb_set_params(b.h, &prm))
	return b
}

// Step == sim/sph.go:64-198 with the particle loops on the GPU.
func (sim *Simulation) Step() {
	sim.IsBusy.Lock()
	defer sim.IsBusy.Unlock()

	// sources spawn on the host exactly as before (math/rand stays in Go), sph.go:72-86
	t := float64(sim.CurrentStep) * sim.Config.DeltaTHalf * 2
	for i := range sim.Config.Sources {
		sim.Root.Particles = append(sim.Root.Particles, sim.Config.Sources[i].Spawn(t)...)
	}
	if sim.Root == nil || len(sim.Root.Particles) == 0 {
		panic("int Run(): Simulation not initialized!") // sph.go:92-94
	}
	b := sim.backend()
	check(b, C.sphb_step(b.h, 1)) // asynchronous; includes the step-0 special case (sph.go:89-103)
	sim.CurrentStep++
	b.hostStale = true
}

// CalculateForces == sim/sph.go:403-435.
func (sim *Simulation) CalculateForces() {
	b := sim.backend()
	check(b, C.sphb_calc_forces(b.h))
	b.hostStale = true
}

func (sim *Simulation) reduce(which C.int32_t) float64 {
	var out C.double
	b := sim.backend()
	check(b, C.sphb_reduce(b.h, which, &out))
	return float64(out)
}

func (sim *Simulation) TotalEnergy() float64   { return sim.reduce(C.SPHB_SUM_E) }
func (sim *Simulation) TotalDensity() float64  { return sim.reduce(C.SPHB_SUM_RHO) }
func (sim *Simulation) TotalMomentum() float64 { return sim.reduce(C.SPHB_LAST_VEL_NORM) } // keeps sph.go:460

// Sync refreshes Root.Particles[i].{Pos,Vel,Rho,C,E,EDot,VDot,EPred,VPred,NNDists[0]} from the device.  The device
// order is cell order; particles are matched by id (= index in Root.Particles), so the host slice keeps its own order.
// withNeighbours additionally fills NearestNeighbours / NNDists / NNPos (descending distance like
// nearest-neighbour.go:139-153) for the examples that draw them.
func (sim *Simulation) Sync(withNeighbours bool) {
	b := lookup(sim)
	if b == nil || b.root != sim.Root || !b.hostStale {
		return
	}
	n := int(C.sphb_count(b.h))
	if n == 0 {
		b.hostStale = false
		return
	}
	f2 := func() []float64 { return make([]float64, 2*n) }
	f1 := func() []float64 { return make([]float64, n) }
	pos, vel, vdot, vpred := f2(), f2(), f2(), f2()
	rho, c, e, edot, epred, h := f1(), f1(), f1(), f1(), f1(), f1()
	id := make([]int64, n)
	mask := C.uint32_t(0)
	ptrs := (*[C.SPHB_F_COUNT]unsafe.Pointer)(C.calloc(C.SPHB_F_COUNT, C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(ptrs))
	// the pointer table lives in C memory; a Go pointer may be stored there only while it is pinned (cgo rules, Go 1.21+)
	var pinner runtime.Pinner
	defer pinner.Unpin()
	set := func(f int, p unsafe.Pointer) { pinner.Pin(p); ptrs[f] = p; mask |= 1 << uint(f) }
	set(C.SPHB_F_POS, unsafe.Pointer(&pos[0]))
	set(C.SPHB_F_VEL, unsafe.Pointer(&vel[0]))
	set(C.SPHB_F_RHO, unsafe.Pointer(&rho[0]))
	set(C.SPHB_F_C, unsafe.Pointer(&c[0]))
	set(C.SPHB_F_E, unsafe.Pointer(&e[0]))
	set(C.SPHB_F_EDOT, unsafe.Pointer(&edot[0]))
	set(C.SPHB_F_VDOT, unsafe.Pointer(&vdot[0]))
	set(C.SPHB_F_EPRED, unsafe.Pointer(&epred[0]))
	set(C.SPHB_F_VPRED, unsafe.Pointer(&vpred[0]))
	set(C.SPHB_F_H, unsafe.Pointer(&h[0]))
	set(C.SPHB_F_ID, unsafe.Pointer(&id[0]))
	var nnIdx []int32
	var nnDist, nnPos []float64
	if withNeighbours {
		nnIdx, nnDist, nnPos = make([]int32, NN_SIZE*n), make([]float64, NN_SIZE*n), make([]float64, 2*NN_SIZE*n)
		set(C.SPHB_F_NN_IDX, unsafe.Pointer(&nnIdx[0]))
		set(C.SPHB_F_NN_DIST, unsafe.Pointer(&nnDist[0]))
		set(C.SPHB_F_NN_POS, unsafe.Pointer(&nnPos[0]))
	}
	var nOut C.int64_t
	check(b, C.sphb_download(b.h, mask, (*unsafe.Pointer)(unsafe.Pointer(ptrs)), C.int64_t(n), &nOut))

	ps := sim.Root.Particles
	devToHost := make([]int, n)
	for j := 0; j < n; j++ {
		devToHost[j] = int(id[j])
	}
	for j := 0; j < n; j++ {
		p := &ps[devToHost[j]]
		p.Pos, p.Vel = Vec2{pos[2*j], pos[2*j+1]}, Vec2{vel[2*j], vel[2*j+1]}
		p.VDot, p.VPred = Vec2{vdot[2*j], vdot[2*j+1]}, Vec2{vpred[2*j], vpred[2*j+1]}
		p.Rho, p.C, p.E, p.EDot, p.EPred = rho[j], c[j], e[j], edot[j], epred[j]
		p.NNDists[0] = h[j]
		if withNeighbours {
			for k := 0; k < NN_SIZE; k++ {
				o := j*NN_SIZE + k
				if nnIdx[o] < 0 {
					p.NearestNeighbours[k] = nil
					continue
				}
				p.NearestNeighbours[k] = &ps[devToHost[nnIdx[o]]]
				p.NNDists[k] = nnDist[o]
				p.NNPos[k] = Vec2{nnPos[2*o], nnPos[2*o+1]}
			}
		}
	}
	b.hostStale = false
}

// FrameData returns what (*Animator).CurrentFrame needs per particle (animator.go:75-101) without mirroring the whole
// state: pixel coordinates (float32(Pos)*float32(size)), the colour-ramp index and Z, in device order.
func (sim *Simulation) FrameData(width, height int) (xy []float32, colour []uint8, z []int64) {
	b := sim.backend()
	n := int(C.sphb_count(b.h))
	if n == 0 {
		return
	}
	xy, colour, z = make([]float32, 2*n), make([]uint8, n), make([]int64, n)
	var nOut C.int64_t
	check(b, C.sphb_frame(b.h, C.int32_t(width), C.int32_t(height), (*C.float)(unsafe.Pointer(&xy[0])),
		(*C.uint8_t)(unsafe.Pointer(&colour[0])), (*C.int64_t)(unsafe.Pointer(&z[0])), C.int64_t(n), &nOut))
	for j := range z { // device id -> Particle.Z, the key the renderer orders by (animator.go:71)
		z[j] = int64(sim.Root.Particles[z[j]].Z)
	}
	return
}

// FindAllNearestNeighboursPeriodic is the batch form of the per-particle loop the examples run
// (examples/density/density.go:65-69): for i { Particles[i].FindNearestNeighboursPeriodic(root, hor, ver) }.
func (sim *Simulation) FindAllNearestNeighboursPeriodic(hor, ver [2]float64) {
	if hor[0] == -math.MaxFloat64 && hor[1] != math.MaxFloat64 {
		panic("cannot have open and periodic boundary in horizontal at same time!")
	}
	b := sim.backend()
	h := [2]C.double{C.double(hor[0]), C.double(hor[1])}
	v := [2]C.double{C.double(ver[0]), C.double(ver[1])}
	check(b, C.sphb_knn(b.h, &h[0], &v[0]))
	b.hostStale = true
	sim.Sync(true)
}

// DensityAll is the batch form of `p.Rho = Density2D(p, sim, kernel)` for all particles (density.go:71-72).
func (sim *Simulation) DensityAll(kernel Kernel) {
	b := sim.backend()
	check(b, C.sphb_density(b.h, kernelID(kernel)))
	b.hostStale = true
	sim.Sync(false)
}

// Close releases the device memory of a simulation that is being replaced (simviewer.go:218-237).
func (sim *Simulation) Close() {
	sim.IsBusy.Lock() // not while a step is being enqueued on this handle
	defer sim.IsBusy.Unlock()
	backendsMu.Lock()
	b := backends[sim]
	delete(backends, sim)
	backendsMu.Unlock()
	if b != nil {
		C.sphb_destroy(b.h)
	}
}
