"""The C-ABI library loads on a CPU-only box, exports every symbol ``include/spgg.h``
declares, agrees with the ctypes mirror on struct layout, and fails loudly (no CPU
fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spgg.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(spgg_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    import spgg_b200
    lib = spgg_b200.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/spgg.h but not exported"
    from spgg_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared


def test_abi_version_and_error_string():
    import spgg_b200
    lib = spgg_b200.load()
    assert lib.spgg_abi_version() == 2
    assert isinstance(lib.spgg_last_error(), bytes)


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof as gcc sees include/spgg.h == the ctypes mirror in _lib.py."""
    from spgg_b200 import _lib
    fields_p = [f[0] for f in _lib.Params._fields_]
    fields_s = [f[0] for f in _lib.Status._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu %zu %d\\n", sizeof(spgg_params_t), sizeof(spgg_status_t), SPGG_NSTAT);']
    for f in fields_p:
        prog.append(f'printf("%zu\\n", offsetof(spgg_params_t, {f}));')
    for f in fields_s:
        prog.append(f'printf("%zu\\n", offsetof(spgg_status_t, {f}));')
    prog.append('return 0;}')
    src = tmp_path / "abi.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "abi"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == C.sizeof(_lib.Params)
    assert int(out[1]) == C.sizeof(_lib.Status)
    assert int(out[2]) == _lib.NSTAT
    offs = [int(x) for x in out[3:]]
    want = [getattr(_lib.Params, f).offset for f in fields_p] + \
           [getattr(_lib.Status, f).offset for f in fields_s]
    assert offs == want


def test_stat_enum_matches_python_mirror():
    from spgg_b200 import _lib
    src = open(HEADER).read()
    enum = dict((m.group(1), int(m.group(2)))
                for m in re.finditer(r"SPGG_(ST_[A-Z0-9_]+)\s*=\s*(\d+)", src))
    assert len(enum) >= 20
    for name, val in enum.items():
        assert getattr(_lib, name) == val, name


def test_argument_validation_does_not_need_a_gpu():
    """Bad arguments are rejected with ValueError (the reference raises ValueError for a bad
    algorithm / state_representation, spgg.py:118,309) before any CUDA call."""
    import spgg_b200
    with pytest.raises(ValueError):
        spgg_b200.Engine(dict(L=32, state_representation="nonsense"))
    with pytest.raises(ValueError):
        spgg_b200.Engine(dict(L=32, algorithm="bogus"))
    with pytest.raises(ValueError):
        spgg_b200.Engine(dict(L=2))
    with pytest.raises(ValueError):
        spgg_b200.Engine(dict(L=32), precision="fp16")


@pytest.mark.skipif(_has_gpu(), reason="box has a GPU")
def test_no_cpu_fallback_compute_fails_loudly_without_cuda():
    import spgg_b200
    with pytest.raises(RuntimeError):
        spgg_b200.Engine(dict(L=32))
    m = spgg_b200.SPGG(L=16, iterations=3, seed=1)
    with pytest.raises(RuntimeError):
        m.run(os.devnull)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neighbor-aware-reinforcement-learning-fosters-cooperation-in-"
                             "spatial-public-goods-games-_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text, f
                assert "spgg_oracle" not in text, f


def test_missing_library_is_an_error(monkeypatch):
    from spgg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libspgg_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
