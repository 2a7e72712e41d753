"""The C-ABI shared library loads without a GPU and exports every symbol the header declares."""
import pytest
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scvae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from scvae_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libscvae_b200.so does not export " + name
    # the ctypes table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.scvae_abi_version() == 1
    assert [lib.scvae_num_heads(k) for k in range(4)] == [1, 2, 2, 3]
    assert lib.scvae_num_heads(99) == -1


def test_argument_errors_are_reported_without_a_gpu():
    from scvae_b200 import _lib
    lib = _lib.load()
    # NULL pointers are rejected before any CUDA call
    assert lib.scvae_likelihood_fwd(1, None, 0, 1, None, 0, 0, 1, 4, None, None, None) != 0
    assert b"likelihood" in lib.scvae_last_error()
    assert lib.scvae_gemm_f32(7, 1, 1, 1, None, 1, None, 1, None, 1, 0, None) != 0


def test_sass_contains_blackwell_tensor_core_and_tma_instructions():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from scvae_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scvae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", ""), f + " references the oracle"


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The descriptor structs cross the C ABI by pointer: every field of the ctypes mirrors in
    `_lib.py` must sit at the offset the C compiler gives it in `include/scvae_b200.h`."""
    import ctypes
    import shutil
    import subprocess
    from scvae_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    pairs = [("scvae_mid_layer", _lib.MidLayer), ("scvae_mid_desc", _lib.MidDesc),
             ("scvae_shadow", _lib.Shadow)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "scvae_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('printf("%s %zu\\n", "{0}", sizeof({0}));'.format(cname))
        for field, _ in cls._fields_:
            lines.append('printf("%s.%s %zu\\n", "{0}", "{1}", offsetof({0}, {1}));'.format(cname, field))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(out["{}.{}".format(cname, field)]) == getattr(cls, field).offset, (cname, field)
