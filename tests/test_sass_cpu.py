"""The shipped library is sm_100a code that uses the Blackwell-native instructions an FP64 solver has: FP64 tensor-core MMAs (SASS DMMA),
TMA bulk copies (UBLKCP) and mbarriers (SYNCS).  (tcgen05 / TMEM have no f64 kind — DESIGN.md.)  Needs only cuobjdump, no GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "nllssolver.jl_b200", "libnlls_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run([exe, "-sass", SO], capture_output=True, text=True, timeout=300).stdout
    per = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = {"DMMA": 0, "UBLKCP": 0, "SYNCS": 0, "arch": None}
            continue
        if name is None:
            continue
        for k in ("DMMA", "UBLKCP", "SYNCS"):
            if re.search(r"\b" + k + r"\b|" + k + r"\.", line):
                per[name][k] += 1
    return out, per


def test_only_sm_100a(sass):
    out, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def test_hot_kernels_use_fp64_tensor_cores_and_tma(sass):
    _, per = sass
    def find(sub):
        ks = [k for k in per if sub in k]
        assert ks, sub
        return ks
    for k in find("schur5_kernel"):                       # Schur elimination: DMMA accumulation fed by a TMA / mbarrier pipeline
        assert per[k]["DMMA"] >= 200 and per[k]["UBLKCP"] >= 3 and per[k]["SYNCS"] >= 8, (k, per[k])
    for k in find("lin_point_kernel"):                    # fused residual + Jacobian + J'WJ: one bulk store of the tile's H span
        assert per[k]["UBLKCP"] >= 1, (k, per[k])
    for sub in ("ldl_diag_kernel", "ldl_off_kernel", "ldl_upd_kernel"):   # tile LDL' of the reduced system on DMMA
        for k in find(sub):
            assert per[k]["DMMA"] >= 8, (k, per[k])
    for k in find("backsub_kernel"):
        assert per[k]["UBLKCP"] >= 1, (k, per[k])
