"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
headers declare, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import HAS_GPU, ROOT


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fvdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared("fvdb.h") + _declared("fvdb_synth.h") + _declared("fvdb_chunk.h")
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"


def test_binding_covers_header(built_lib):
    from fabstir_vectordb_b200 import _lib
    declared = set(_declared("fvdb.h") + _declared("fvdb_synth.h") + _declared("fvdb_chunk.h"))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    assert lib.fvdb_abi_version() == 1


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "fvdb.h"\n#include "fvdb_synth.h"\n#include "fvdb_chunk.h"\nint main(void){return FVDB_OK;}\n')
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "include"), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


@pytest.mark.skipif(HAS_GPU, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(built_lib):
    from fabstir_vectordb_b200 import Engine, NoDevice
    with pytest.raises(NoDevice):
        Engine(384)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fabstir_vectordb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "fvdb_oracle" not in text, f


def _build_c_driver(tmp_path, built_lib):
    import subprocess
    exe = tmp_path / "abi_driver"
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_driver.c"),
                           "-L", libdir, "-lfvdb_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)])
    return exe


@pytest.mark.skipif(HAS_GPU, reason="only meaningful on a box without a GPU")
def test_c_driver_links_and_sees_no_device(tmp_path, built_lib):
    """A plain C99 program (tests/c/abi_driver.c) builds against include/fvdb.h alone, links the library and —
    without a GPU — is told FVDB_ERR_NO_DEVICE: there is no CPU path behind the ABI."""
    import subprocess
    exe = _build_c_driver(tmp_path, built_lib)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("no-device"), (out.stdout, out.stderr)


@pytest.mark.gpu
def test_c_driver_end_to_end(tmp_path, built_lib):
    """The same program on a GPU: inserts, tombstone, hybrid search, 3x post-filter, a cosine handle — every
    result bit-identical to the scalar loops in the C file (which restate src/core/vector_ops.rs:39-57)."""
    import subprocess
    exe = _build_c_driver(tmp_path, built_lib)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "abi-driver ok" in out.stdout, (out.stdout, out.stderr)
