"""bench.py pieces that run without a GPU: roofline bookkeeping and the reference arm's JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_match_survey_table():
    """c2 forward 20.4 MB; the step here skips the separate derivative gather"""
    b = _bench()
    alg = b.algorithmic_bytes(32768, 529617, 64, 4)
    fwd = alg["pair_forward"] + alg["spread"] + alg["kfilter"] + alg["gather"]
    assert abs(fwd / 1e6 - 20.4) < 0.1
    n, s = 32768, 4
    survey_step = 45.6e6
    skipped = 64 ** 3 * s + (3 + 1) * s * n + 3 * s * n      # R + A + 3sN: dV/dr comes out of the forward gather
    assert abs(sum(alg.values()) + skipped - survey_step) < 0.1e6


def test_dominant_kernel_selection():
    b = _bench()
    with open(os.path.join(ROOT, "profiles", "r01b_bench_c4.json")) as f:
        stages = json.load(f)["stages"]
    stage, name, launches, ms, alg, share = b.dominant_kernel(stages, 5)
    assert (stage, launches) == ("spread", 2) and "spread" in name
    assert abs(ms - stages["spread"]["ms"]) < 1e-12 and alg == stages["spread"]["alg_bytes"]
    assert 0.1 < share < 0.5
    # a filter stage that dwarfs everything else wins as one representative FFT kernel
    fat = dict(stages)
    fat["kfilter"] = dict(stages["kfilter"], ms=10.0)
    stage, name, launches, ms, alg, _ = b.dominant_kernel(fat, 5)
    assert stage == "kfilter" and launches == 2 and abs(ms - 2.0) < 1e-12
    assert alg == stages["kfilter"]["alg_bytes"] / 5


def test_reference_arm_json_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "2",
           "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "atom-steps/s" and line["value"] > 0
    assert (line["steps"], line["warmup"]) == (2, 1)          # the arm honours the requested steps / warm-up
    # the unmodified reference when oracle/_ref exists (build()), the numpy port otherwise
    have_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "torchpme"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("c2")
    assert line["e2e"] == {"value": line["value"], "unit": "atom-steps/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    # other ranks of a torchrun launch print nothing and exit 0
    env["RANK"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_default_workloads():
    """N = 1 headlines c3 (the largest single-GPU config), N > 1 the slab-decomposed c4"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'args.workload = "c4" if (world > 1 or args.gpus > 1) and args.decomposition == "slab" else "c3"' in src
    assert 'args.decomposition = "slab" if world > 1 else "replica"' in src


def test_shuffle_inputs_is_a_relabelling():
    import torch

    b = _bench()
    from torchpme_b200.synthetic import rocksalt

    pos, q, cell, idx, d = rocksalt(4, cutoff=5.0)
    p2, q2, i2, d2 = b.shuffle_inputs(pos, q, idx, d)
    assert (i2[1:, 0] >= i2[:-1, 0]).all()                     # sorted by the first index
    # same multiset of (distance, charge product) pairs, same geometry
    key = lambda qq, ii, dd: torch.sort(dd * 1000 + qq[ii[:, 0], 0] * qq[ii[:, 1], 0]).values  # noqa: E731
    assert torch.allclose(key(q, idx, d), key(q2, i2, d2))
    delta = p2[i2[:, 1]] - p2[i2[:, 0]]
    delta = delta - torch.round(delta / cell[0, 0]) * cell[0, 0]
    assert torch.allclose(delta.norm(dim=1), d2, atol=1e-12)
