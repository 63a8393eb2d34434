"""The C-ABI library loads on a CPU-only box and exports every symbol include/dmp.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dmp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|int64_t)\s+(dmp_\w+)\s*\(", src, flags=re.M)))


def test_header_symbols_exported():
    names = declared_functions()
    assert len(names) >= 19, names
    lib = ctypes.CDLL(os.path.join(ROOT, "snac_b200", "libdmp.so"))
    for n in names:
        assert hasattr(lib, n), "libdmp.so does not export %s" % n


def test_python_binding_covers_header():
    from snac_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_functions()
    assert _lib.lib.dmp_abi_version() == _lib.ABI_VERSION


def test_layout_and_argument_checks_without_gpu():
    from snac_b200 import _lib as L
    lay = L.DmpLayout()
    assert L.lib.dmp_layout(1, 10, ctypes.byref(lay)) == L.OK
    assert (lay.cells_bytes, lay.aux_bytes, lay.obs_dim, lay.n_actions, lay.grid_cols) == (640, 80, 7, 3, 34)
    assert L.lib.dmp_layout(2, 10, ctypes.byref(lay)) == L.OK
    assert (lay.cells_bytes, lay.aux_bytes, lay.obs_dim, lay.n_actions) == (640, 0, 51, 5)
    assert L.lib.dmp_layout(3, 10, ctypes.byref(lay)) == L.OK
    assert (lay.cells_bytes, lay.aux_bytes, lay.obs_dim, lay.n_actions) == (10080, 160, 51, 8)      # u16 maps + the nibble maps (208 B per env)
    assert (lay.total_step_static, lay.total_step_dynamic) == (1300, 1000)
    assert L.lib.dmp_layout(4, 10, ctypes.byref(lay)) == L.EINVAL
    # argument validation happens before any CUDA call
    assert L.lib.dmp_plan_static(2, 2, 1, 1, None) == L.EINVAL          # reference: ValueError at reset()
    assert L.lib.dmp_plan_static(1, 3, 1, 1, None) == L.EINVAL
    st = L.DmpState()
    assert L.lib.dmp_rollout(ctypes.byref(st), None, 1, None) == L.EINVAL
    assert L.lib.dmp_stats_scratch_bytes(1 << 20) == 1024 * 32


def test_struct_sizes_match_header():
    """ctypes mirrors of the C structs (field order and padding)."""
    from snac_b200 import _lib as L
    assert ctypes.sizeof(L.DmpIO) == 6 * 8 + 8
    assert ctypes.sizeof(L.DmpState) == 6 * 4 + 5 * 8 + 9 * 8
    assert ctypes.sizeof(L.DmpLayout) == 3 * 8 + 8 * 4


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "snac_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_header_is_plain_c_and_offsets_match_ctypes(tmp_path):
    """include/dmp.h compiles as C (a C or cgo caller can include it) and every struct field sits where the ctypes
    mirrors of snac_b200/_lib.py put it."""
    import shutil
    import subprocess
    from snac_b200 import _lib as L
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the build image"
    structs = {"DmpState": L.DmpState, "DmpIO": L.DmpIO, "DmpLayout": L.DmpLayout}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dmp.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('  printf("%s.sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.splitlines())
    for name, cls in structs.items():
        assert int(got[name + ".sizeof"]) == ctypes.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (name, fname)]) == getattr(cls, fname).offset, (name, fname)
