"""bench.py contract checks that need no GPU: the reference arm's JSON line and the synthetic generators."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "cfg3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "molecules/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
    # both arms print the same config keys for the same arguments
    class A:   # noqa: N801
        workload, batch, augment, out_dtype, molecules = "cfg3", 0, False, "float32", 0
    assert d["config"] == bench.workload_config(A, 1)
    assert d["config"]["molecules"] == 262_144 and d["scaling"] == "strong"


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_batches_are_seeded_and_shaped():
    a, b = bench.make_batch("cfg4", 32, seed=5), bench.make_batch("cfg4", 32, seed=5)
    assert np.array_equal(a["coords"], b["coords"]) and np.array_equal(a["types"], b["types"])
    counts = np.diff(a["offs"])
    assert counts.min() >= 40 and counts.max() <= 60 and a["types"].max() < 9
    assert np.array_equal(a["coords"], a["coords"].astype(np.float32).astype(np.float64))   # fp32-representable
    # ligands are recentred random walks with 1.5 A steps
    m0 = a["coords"][a["offs"][0]:a["offs"][1]]
    assert np.abs(m0.mean(0)).max() < 1e-5
    assert np.allclose(np.linalg.norm(np.diff(m0, axis=0), axis=1), 1.5, atol=1e-4)
    c2 = bench.make_batch("cfg2", 2, seed=1)
    assert c2["feats"].shape == (4000, 16) and set(np.unique(c2["feats"])) <= {0.0, 1.0}
    c5 = bench.make_batch("cfg5", 1, seed=1)
    assert c5["feats"].shape == (10000, 32) and c5["radii"].min() >= 1.0 and c5["radii"].max() <= 2.0
