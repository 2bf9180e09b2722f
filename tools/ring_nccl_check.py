#!/usr/bin/env python
"""The in-library ring over NCCL against the single handle (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ring_nccl_check.py

Every rank owns an x-slab of a periodic jittered lattice with a bulk flow (particles migrate and cross the seam), steps
it with sphb_ring_step (reuse cycles included) and sends its particles to rank 0, which compares them with the same
steps on one handle holding all particles.  Exit code 0 = agreement to 1e-9 along the trajectory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sphugo_b200 import _lib as L, gen, slab  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = int(os.environ.get("RING_CHECK_NX", "192"))
    steps = int(os.environ.get("RING_CHECK_STEPS", "11"))
    pos = gen.jittered_lattice(nx, nx)
    n = len(pos)
    vx = float(os.environ.get("RING_CHECK_VX", "1.5"))  # 1.5: a migration nearly every step; 0.1: one in ten (fused keys)
    vel = np.tile([[vx, -0.7 * vx / 1.5]], (n, 1))
    e = np.full(n, 0.01)
    kw = dict(hor=(0.0, 1.0), ver=(0.0, 1.0), accel=(0.0, 0.2), dt_half=0.4 / nx)
    topo = slab.Topology(world, [k / world for k in range(world + 1)], True)
    own = topo.owner_of(pos[:, 0]) == rank
    ids = np.arange(n, dtype=np.int64)
    sim = slab.RingSim(L.make_params(device=local, **kw), topo, rank, pos[own], vel[own], e[own], ids[own],
                       h_max_hint=slab.default_h_hint(n, 1.0), capacity=n, halo_cap=n)
    sim.step(steps)
    st = sim.handle.download(["pos", "vel", "e", "rho", "h", "id"])
    parts = [None] * world
    dist.gather_object(st, parts if rank == 0 else None, 0)
    info, cnt = sim.handle.ring_info(), sim.handle.counters()
    rc = 0
    if rank == 0:
        d = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        o = np.argsort(d["id"], kind="stable")
        d = {k: v[o] for k, v in d.items()}
        g = L.Handle(L.make_params(device=local, **kw), pos, vel, e)
        g.step(steps)
        ref = g.state(["pos", "vel", "e", "rho", "h", "id"])
        g.close()
        assert len(d["id"]) == n and (d["id"] == ref["id"]).all(), "particles lost or duplicated"
        worst = {}
        for f in ("pos", "vel", "e", "rho", "h"):
            den = np.maximum(np.abs(ref[f]), np.abs(ref[f]).max() * 1e-3)
            worst[f] = float(np.max(np.abs(d[f] - ref[f]) / den))
        ok = all(v <= 1e-9 for v in worst.values())
        print(f"ring over NCCL, {world} ranks, {n} particles, {steps} steps: worst relative differences {worst}; "
              f"rank 0: reuse evaluations {cnt['reuse_steps']}, migrations {int(info['migrations'])}, period {int(info['period'])} -> "
              + ("OK" if ok else "MISMATCH"), flush=True)
        rc = 0 if ok else 1
    sim.handle.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
