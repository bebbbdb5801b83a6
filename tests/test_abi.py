"""The C-ABI library loads, exports every symbol include/nrays_b200.h declares, and the ctypes mirror
matches the C struct layout.  No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from nrays_b200 import _abi as A
from nrays_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrays_b200.h")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(nrb_[a-z0-9_]+)\s*\(", text))
    assert declared == set(A.EXPORTS), declared ^ set(A.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nrb_version().startswith(b"nrays_b200")


def test_struct_layout_matches_c():
    structs = ["NrbNodeDesc", "NrbLightDesc", "NrbMaterialDesc", "NrbTextureDesc", "NrbSceneDesc", "NrbCamera",
               "NrbTileSet", "NrbStats", "NrbBuildInfo", "NrbBuildOptions", "NrbIpcHandle"]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void){"]
    for s in structs:
        prog.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f, _t in getattr(A, s)._fields_:
            prog.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    prog.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write("\n".join(prog))
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c11", "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    for s in structs:
        cls = getattr(A, s)
        assert int(got[s]) == C.sizeof(cls), s
        for f, _t in cls._fields_:
            assert int(got["%s.%s" % (s, f)]) == getattr(cls, f).offset, "%s.%s" % (s, f)


def test_tile_helpers_need_no_device():
    lib = _lib.load()
    assert lib.nrb_tile_count(1920, 1080) == 120 * 68
    assert lib.nrb_tile_count(16, 16) == 1 and lib.nrb_tile_count(17, 1) == 2
    ts = A.NrbTileSet(3, 8)
    assert lib.nrb_tile_count_local(1920, 1080, C.byref(ts)) == len(range(3, 120 * 68, 8))


def test_product_fails_loudly_without_device():
    """On a box without a GPU the product path must raise, never fall back to a CPU renderer."""
    lib = _lib.load()
    if lib.nrb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from nrays_b200 import Ball, Isometry3, Light, NormalMaterial, Scene, SceneNode

    with pytest.raises(_lib.NraysError) as e:
        Scene([SceneNode(NormalMaterial(), 0, 0, 1.0, 1.0, Isometry3.identity(), Ball(1.0))], [Light((0, 5, 0), 0, 1, (1, 1, 1))])
    assert e.value.status == A.NRB_ERR_NO_DEVICE


def test_no_oracle_on_product_path():
    """Nothing under nrays_b200/ may import, link or execute the oracle."""
    banned = ["oracle_lib", "libnrays_oracle", "nro_", "nrays_oracle", "import oracle", "from oracle", "oracle/"]
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "nrays_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for b in banned:
                    assert b not in txt, "%s references %r" % (os.path.join(dirpath, f), b)
