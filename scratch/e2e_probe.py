import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from sphugo_b200 import _lib as L, gen
nx, ny = 8192, 4096
pos = gen.jittered_lattice(nx, ny, (0, 0), (0.5, 0.25))
n = len(pos)
prm = L.make_params(hor=(0.0, 0.5), ver=(0.0, 0.25), accel=(0.0, 0.2), dt_half=6e-5)
g = L.Handle(prm, pos, None, np.full(n, 0.01))
g.step(3); g.sync()
host = {f: torch.empty((n,) + L.FIELD_SHAPE[f][0], dtype=torch.float64).pin_memory().numpy() for f in ("pos", "vel", "e")}
g.download(["pos", "vel", "e"], out=host)
fr = {"xy": torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(),
      "colour": torch.empty((n,), dtype=torch.uint8).pin_memory().numpy(),
      "id": torch.empty((n,), dtype=torch.int64).pin_memory().numpy()}
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
for it in range(4):
    t0 = time.perf_counter(); g.upload(**host); tick("upload", t0)
    t0 = time.perf_counter(); g.step(1); g.sync(); tick("step", t0)
    t0 = time.perf_counter(); g.frame(1280, 720, out=fr); tick("frame", t0)
    t0 = time.perf_counter(); g.reduce(L.SUM_E); tick("reduce", t0)
    t0 = time.perf_counter(); g.download(["pos", "rho", "h", "id"]); tick("download_unpinned", t0)
print({k: round(v / 4 * 1e3, 2) for k, v in T.items()})
