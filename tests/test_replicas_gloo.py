"""N>1 path on CPU: world_size-2 gloo job, one scene replica per rank, only timings/checksums cross ranks."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_two_replicas_over_gloo(cpu_libs):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_replica_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and len(r["times"]) == 2
    # replicas of the same scene are bit-identical; the whole-job value uses the slowest rank
    assert r["checksums"][0] == r["checksums"][1]
    assert abs(r["value"] - 2 * r["proj_per_rank"] / max(r["times"])) < 1e-6 * r["value"]
    assert r["times"][1] > r["times"][0] - 1.0


def test_single_process_degenerates_cleanly():
    from positionbaseddynamics_b200 import replicas
    assert replicas.gather(None, 1.5) == [1.5]
    assert replicas.whole_job_throughput(100, [2.0]) == 50.0
