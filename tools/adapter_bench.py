"""Drop-in measurement (test infrastructure, not product): the UNMODIFIED reference builds cfg2, steps it with its own
TimeStepController on the host, then integration/GpuTimeStepController.h is installed with Simulation::setTimeStep and the same
reference call site (`ts->step(model)`) is timed again.  Needs oracle/_ref/libpbdref_gpu_f32.so (oracle/Makefile target refgpu)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from oracle import pyoracle

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cpu-steps", type=int, default=2)
ap.add_argument("--gpu-steps", type=int, default=20)
ap.add_argument("--threads", type=int, default=16)
ap.add_argument("--precision", default="f32")
ap.add_argument("--device-authoritative", action="store_true", help="GpuTimeStepController::setHostStateAuthoritative(false): no per-step upload")
a = ap.parse_args()

m = pyoracle.CpuPbd("refgpu", a.precision)
m.set_threads(a.threads)
t0 = time.perf_counter(); scenes.cfg2(m, a.n, a.iters); m.init_groups(); t_build = time.perf_counter() - t0
nc = m.num_constraints()
proj = nc * a.iters
m.step(1)
cpu_s = m.step(a.cpu_steps) / a.cpu_steps
x_cpu = m.get("x").copy()
m.use_gpu_timestep(0, 4)   # PBD_MODE_AUTO
if a.device_authoritative:
    m.set_host_state_authoritative(False)
t_bind = m.step(1)                          # first step: bind = flatten the reference model + upload + CUDA graph capture
m.step(2)                                   # warm-up
gpu_s = m.step(a.gpu_steps) / a.gpu_steps   # wall clock around the reference's own ts->step(model) loop
assert m.gpu_error() == "" and np.isfinite(m.get("x")).all()
print(json.dumps({"scene": "cfg2 %dx%d, %d constraints, %d iterations" % (a.n, a.n, nc, a.iters), "Real": a.precision,
                  "reference_build_s": round(t_build, 2), "reference_cpu_ms_per_step": round(cpu_s * 1e3, 2), "cpu_threads": a.threads,
                  "host_state": "device authoritative (download only)" if a.device_authoritative else "host authoritative (upload + download every step)",
                  "adapter_first_step_s": round(t_bind, 3), "adapter_gpu_ms_per_step": round(gpu_s * 1e3, 3), "speedup": round(cpu_s / gpu_s, 1),
                  "adapter_projections_per_s": proj / gpu_s, "reference_projections_per_s": proj / cpu_s}))
