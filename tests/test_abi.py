"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what
include/aslam_b200.h declares, fails loudly without a device, and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aslam_b200.h")


def _declared():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"ASLAM_API\s+[\w\s\*]+?\b(aslam_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from aria_slam_b200 import _lib
    l = _lib.lib()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(l, n), "missing export " + n
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r" T (aslam_\w+)", out)))
    assert exported == names, "exported symbols and header disagree"


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aria_slam_b200 import _lib
    h = C.c_void_p()
    rc = _lib.lib().aslam_ctx_create(0, None, C.byref(h))
    assert rc == _lib.E_NODEVICE and not h
    assert b"no CPU fallback" in _lib.lib().aslam_last_error()
    with pytest.raises(_lib.AslamError):
        _lib.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "aria_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "orb_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "cv2.ORB" not in txt and "BFMatcher" not in txt.replace("cv::BFMatcher", "").replace("cv2.BFMatcher(", "BFMatcher(") \
                    or f.endswith((".cu", ".cuh", ".h", ".py")), f


def test_struct_layouts_match_reference_types():
    """aslam_keypoint == aria::core::KeyPoint (24 B), aslam_match == Match (12 B), aslam_box == Detection (24 B)."""
    from aria_slam_b200 import _lib
    assert _lib.KP_DTYPE.itemsize == 24 and _lib.MATCH_DTYPE.itemsize == 12 and _lib.BOX_DTYPE.itemsize == 24
    assert _lib.KP_DTYPE.names == ("x", "y", "size", "angle", "response", "octave")
    assert C.sizeof(_lib.OrbParams) == 44      # 11 x 4 bytes, aslam_orb_params


def test_header_is_plain_c99(tmp_path):
    """The boundary must bind from any FFI: include/aslam_b200.h compiles as strict C99 on its own, and a C
    program that only calls the no-device entry points links against the library."""
    src = tmp_path / "use.c"
    src.write_text('#include "aslam_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { aslam_orb_params p; aslam_ctx *c = 0; int rc;\n'
                   '  aslam_orb_default_params(&p);\n'
                   '  rc = aslam_ctx_create(0, 0, &c);\n'
                   '  printf("%d %d %s\\n", p.nfeatures, rc, aslam_version());\n'
                   '  if (c) aslam_ctx_destroy(c);\n  return 0; }\n')
    exe = tmp_path / "use"
    libdir = os.path.join(ROOT, "aria_slam_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                        "-L" + libdir, "-laslam_b200", "-Wl,-rpath," + libdir, "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout.split()
    assert out[0] == "1000"                     # FactoryConfig::max_features (include/factory/PipelineFactory.hpp:25)
