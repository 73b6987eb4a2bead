"""CPU: the drop-in adapters compile (a) against the stand-in headers of adapter/shim and link with libb200ba.so, and (b),
when the reference tree is present, against the reference's OWN headers (src/sfm_ba.h, src/sfm_utils.h,
src/sfm_triangulation.h + its vendored OpenCV 2.4.9 headers) - i.e. the signatures are exactly the ones the tracker calls."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
ADAPTERS = ["sfm_ba_b200.cpp", "sfm_post_b200.cpp", "sfm_tri_b200.cpp"]


@pytest.mark.parametrize("src", ADAPTERS)
def test_adapter_builds_against_the_shim_and_links(tmp_path, src):
    so = str(tmp_path / (src + ".so"))
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", f"-I{ROOT}/adapter/shim", f"-I{ROOT}/include",
                           f"{ROOT}/adapter/{src}", f"-L{ROOT}/sequencesfm_b200", "-lb200ba", "-Wl,--no-undefined", "-o", so])


@pytest.mark.skipif(not os.path.isdir(f"{REF}/src"), reason="reference tree not present")
@pytest.mark.parametrize("src", ADAPTERS)
def test_adapter_matches_the_reference_declarations(src):
    """-fsyntax-only with the reference's headers first on the include path: a signature that differed from the reference's
    declaration would be a redeclaration error / an unresolved overload here."""
    extra = []
    if src == "sfm_tri_b200.cpp":
        extra = ["-include", f"{REF}/src/sfm_triangulation.h"]
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-fpermissive", "-w", f"-I{REF}/src", f"-I{REF}/cmake/opencv-2.4.9/include",
                           f"-I{REF}/Thirdparty/eigen3", f"-I{ROOT}/include"] + extra + [f"{ROOT}/adapter/{src}"])
