"""Worker of tests/test_replicas_gloo.py: one rank of a world_size-2 gloo job.  Each rank owns one scene replica (here
stepped by the CPU restatement, since the box has no GPU), the ranks exchange only timings and checksums."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from positionbaseddynamics_b200 import replicas  # noqa: E402
import scenes  # noqa: E402
from oracle import pyoracle  # noqa: E402

rank, world, local = replicas.env_world()
dist = replicas.init("gloo")
cpu = pyoracle.CpuPbd("oracle", "f32")
cpu.set_threads(1)
scenes.cloth(cpu, 20, 20, 4, 3, dist_k=1e5, bend_k=100.0, max_iter=4)
if dist is not None:
    dist.barrier()
secs = cpu.step(3) + 0.01 * rank  # make the ranks' timings differ deterministically
proj = cpu.num_constraints() * 1 * 4 * 3
times = replicas.gather(dist, secs)
sums = replicas.gather(dist, replicas.checksum(cpu.get("x")))
if rank == 0:
    print(json.dumps({"world": world, "times": times, "checksums": sums, "value": replicas.whole_job_throughput(proj, times),
                      "proj_per_rank": proj}))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
