"""The C-ABI shared library loads, exports every symbol include/msm_b200.h declares, agrees with the header on
struct layout, and fails loudly (no CPU fallback) when no GPU is present.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import msm_b200
from msm_b200 import _lib
from conftest import HAS_GPU, ROOT

HEADER = os.path.join(ROOT, "include", "msm_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msm_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 40
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in msm_b200.h but not exported by libmsm_b200.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    for n in _lib.PROTOTYPES:
        assert n in names, f"{n} bound in Python but not declared in the header"


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "msm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(msm_config),sizeof(msm_sim_params),sizeof(msm_stream_state),sizeof(msm_derived),"
                   "sizeof(msm_profile_record));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.MsmConfig), C.sizeof(_lib.MsmSimParams), C.sizeof(_lib.MsmStreamState),
            C.sizeof(_lib.MsmDerived), C.sizeof(_lib.MsmProfileRecord)]
    assert got == want


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89.c"
    src.write_text('#include "msm_b200.h"\nint main(void){return MSM_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "c89.o")], check=True)


def test_version_and_strerror():
    assert b"sm_100a" in _lib.lib.msm_version()
    assert _lib.lib.msm_strerror(_lib.MSM_E_ALIASING) == b"Fourier aliasing above threshold"


def test_bad_arguments_are_rejected_before_touching_the_gpu():
    cfg = _lib.MsmConfig()
    h = C.c_void_p()
    assert _lib.lib.msm_create(C.byref(cfg), C.byref(h)) == _lib.MSM_E_ARG       # struct_size = 0
    cfg.struct_size = C.sizeof(_lib.MsmConfig)
    cfg.dims, cfg.size, cfg.n_streams, cfg.dx = 3, 48, 1, 1.0                     # not a power of two
    assert _lib.lib.msm_create(C.byref(cfg), C.byref(h)) == _lib.MSM_E_ARG
    assert b"power of two" in _lib.lib.msm_last_error(None)
    cfg.size, cfg.dims = 16, 4
    assert _lib.lib.msm_create(C.byref(cfg), C.byref(h)) == _lib.MSM_E_ARG


@pytest.mark.skipif(HAS_GPU, reason="checks the behaviour on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(msm_b200.MsmError) as ei:
        msm_b200.Context(3, 16, 1, 1.0, 1.0, -1.0)
    assert ei.value.code == _lib.MSM_E_CUDA and "no CPU fallback" in str(ei.value)
    import numpy as np
    with pytest.raises(msm_b200.MsmError):
        msm_b200.forward(np.ones((4, 4), dtype=np.complex128))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "msm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "msm_oracle" not in text, f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py may touch the oracle only inside the two CPU-baseline functions
    assert bench.count("from oracle import") == 3


def test_native_host_binary_is_built_and_links_only_the_c_abi():
    exe = os.path.join(ROOT, "msm_b200", "msm-simulator-b200")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libmsm_b200.so" in needed and "torch" not in needed and "python" not in needed


def test_no_store_backedge_hazard_in_the_built_kernels():
    """Static guard for the ptxas problem of DESIGN.md section 6: in the SASS of every built pass kernel, no instruction
    reached through a loop back-edge overwrites a register that a store before the branch still has to read without a
    wait for that store's read barrier (or a MEMBAR) in between.  scripts/check_war_hazard.py flags exactly the pattern
    found in the build without the tile-boundary fence."""
    import glob
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    objs = sorted(glob.glob(os.path.join(ROOT, "msm_b200", "csrc", "build", "fft_*.o")))   # incl. fft_tma.o
    core = os.path.join(ROOT, "msm_b200", "csrc", "build", "core.o")
    if os.path.exists(core):
        objs.append(core)
    if not objs:
        pytest.skip("no kernel objects (library built elsewhere)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_war_hazard.py")] + objs, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
