"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/fs2_b200.h
declares, and fails loudly (no CPU fallback) when asked to create a handle without a CUDA device."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fs2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from smart_nar_fast_tts_b200.capi import SIGNATURES
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib.lib, s), f"{s} declared in include/fs2_b200.h but not exported"
        assert s in SIGNATURES, f"{s} has no ctypes signature"
    assert set(SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.fs2_version()


def test_library_is_sm100a_with_tcgen05(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", lib.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert re.search(r"\bUTC\w*MMA", out), "no tcgen05.mma (UTC*MMA) in SASS"
    assert "UTMALDG" in out, "no TMA loads in SASS"
    assert "LDTM" in out, "no tcgen05.ld in SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from smart_nar_fast_tts_b200.capi import Dims, Fs2Error
    hp = C.c_void_p()
    d = Dims(361, 256, 4, 4, 2, 1024, 9, 1, 256, 3, 256, 80, 512, 5, 5, 1000, 0, 0)
    rc = lib.fs2_create(C.byref(hp), C.byref(d), 0)
    assert rc == -6 and hp.value is None
    assert b"no CPU fallback" in lib.fs2_last_error(None)
    with pytest.raises(Fs2Error):
        lib.check(rc, None)
    # the module refuses CPU tensors instead of computing in PyTorch
    import fs2_oracle as O
    from helpers import build_model
    m = build_model(O.make_state_dict(0), O.STATS_NAN_BINS, device="cpu")
    sp, tx, sl, L = O.make_inputs(2, 4, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(sp, tx, sl, L)


def test_no_global_access_before_pdl_wait():
    """Every kernel is launched with programmatic dependent launch; ptxas may hoist non-coherent loads above
    griddepcontrol.wait.  scripts/check_pdl_sass.py scans the SASS of the built objects for any global access scheduled
    before the wait (the bug class behind the stale layout-table reads of profiles/r1g)."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "smart-nar_fast_tts_b200", "build")):
        pytest.skip("objects not built in this checkout")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "check_pdl_sass.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
