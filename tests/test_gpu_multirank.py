"""N ranks == 1 rank: the point-sharded NCCL path must reproduce the single-GPU trajectory (sharding only changes the
summation order of the camera blocks, the reduced system and the scalars).  Needs >= 2 visible GPUs (skipped otherwise; run with
`gpurun --gpus 2`).  Also the two rank-divergent termination inputs: one rank's clock / one rank's callback flag must
stop ALL ranks at the same iteration instead of stranding the others in a collective."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(nproc, out, *extra, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_worker.py"), "--out", out, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return np.load(out)


needs2 = pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("shape,iters", [("ladybug", 8), ("300,60000,300000", 6)])
def test_nrank_equals_one_rank(pkg, tmp_path, shape, iters):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multirank_worker as mw
    nproc = min(_ngpus(), int(os.environ.get("NLLS_TEST_RANKS", "2")))
    d = _launch(nproc, str(tmp_path / "n.npz"), "--shape", shape, "--iters", str(iters))
    p = mw.build_problem(pkg, shape)
    trace1, res1, cams1, pts1, _ = mw.run(pkg, p, 0, 1, 0, iters, 1e5)
    # every rank reports the same trajectory (bitwise: all decisions are taken on all-reduced / broadcast values)
    for r in range(1, nproc):
        assert np.array_equal(d["traces"][r], d["traces"][0])
        assert np.array_equal(d["cams"][r], d["cams"][0])
    tn = d["trace"]
    assert tn.shape == trace1.shape
    assert np.array_equal(tn[:, 1], trace1[:, 1])                       # same inner-try counts
    assert np.array_equal(tn[:, 3], trace1[:, 3])                       # same termination words
    assert np.max(np.abs(tn[:, 0] - trace1[:, 0]) / trace1[:, 0]) <= 1e-10   # per-iteration cost
    assert np.max(np.abs(tn[:, 2] - trace1[:, 2]) / trace1[:, 2]) <= 1e-6    # lambda
    assert abs(d["bestcost"][0] - res1.bestcost) <= 1e-10 * res1.bestcost
    scale = max(np.max(np.abs(cams1)), np.max(np.abs(pts1)))
    assert np.max(np.abs(d["cams"][0] - cams1)) <= 1e-7 * scale          # variables (the gauge directions are the loosest)
    assert np.max(np.abs(d["points"] - pts1)) <= 1e-7 * scale


@needs2
@pytest.mark.parametrize("mode,word", [("maxtime", 1 << 9), ("callback", 5 << 16)])
def test_rank_divergent_termination_is_collective(tmp_path, mode, word):
    d = _launch(2, str(tmp_path / "t.npz"), "--shape", "ladybug", "--iters", "6", "--mode", mode, timeout=300)
    assert d["termination"][0] == d["termination"][1]                   # no rank left alone (the run did not hang)
    assert d["niterations"][0] == d["niterations"][1] == (1 if mode == "maxtime" else 2)
    assert int(d["termination"][0]) & word == word
