"""bench.py's CPU arm (`--impl reference`): runs without a GPU, prints ONE JSON line with the keys the driver reads,
uses every host core even when the launcher exports OMP_NUM_THREADS=1 (torchrun does), and non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, ROCQ_REF_MAX_QUBITS="18", OMP_NUM_THREADS="1", **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_json_line():
    lines = _run({})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gates_per_sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("C2: 30-qubit")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "18 qubits" in cb["sample"]
    ncores = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncores                                   # not the launcher's OMP_NUM_THREADS=1
    comp = cb["compiled_reference"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhipStateVec_ref_c64.so")):
        assert comp["kind"] == "reference" and comp["value"] > 0 and comp["cores"] == ncores
    else:
        assert comp is None


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
