"""The C-ABI library: builds for sm_100a, loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

import sys
sys.path.insert(0, str(Path(__file__).resolve().parent))
from _util import has_cuda  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from upright_b200 import bindings
    return bindings.load_library()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "upright_b200.h").read_text()
    declared = set(re.findall(r"\b(ub_[a-z_]+)\s*\(", header))
    from upright_b200 import bindings
    assert declared == set(bindings.EXPORTED_SYMBOLS), declared ^ set(bindings.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_struct_sizes_match_header(lib):
    """ctypes mirror and C struct agree (compiled probe with gcc)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "upright_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu", sizeof(ub_problem_desc_t), sizeof(ub_joint_t), sizeof(ub_contact_t), sizeof(ub_sphere_t), sizeof(ub_closed_loop_params_t), sizeof(ub_obstacle_mode_t));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        p = Path(d) / "probe.c"
        p.write_text(src)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(d) / "probe")])
        sizes = list(map(int, subprocess.check_output([str(Path(d) / "probe")]).split()))
    from upright_b200 import bindings as B
    assert sizes == [C.sizeof(B.ProblemDesc), C.sizeof(B.Joint), C.sizeof(B.Contact), C.sizeof(B.Sphere), C.sizeof(B.ClosedLoopParams), C.sizeof(B.ObstacleMode)]


def test_sass_is_sm100a(lib):
    import subprocess
    from upright_b200 import bindings
    out = subprocess.run(["cuobjdump", "-lelf", str(bindings.library_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(has_cuda(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly(lib):
    from upright_b200 import bindings, problem_io
    desc, _ = problem_io.load_fixture("cfg2_thing_demo")
    handle = C.c_void_p()
    rc = lib.ub_problem_create(C.byref(desc), C.byref(handle))
    assert rc == -3  # UB_E_NO_DEVICE
    assert b"no CPU fallback" in lib.ub_last_error()
    from upright_b200.engine import BatchedMPC
    with pytest.raises(RuntimeError):
        BatchedMPC(desc)


def test_product_does_not_import_oracle():
    """The product package must not reference the test oracle."""
    for py in (ROOT / "upright_b200").rglob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py
    for src in (ROOT / "upright_b200" / "csrc").glob("*"):
        assert not re.search(r'#include\s*[<"][^>"]*oracle', src.read_text()), src
