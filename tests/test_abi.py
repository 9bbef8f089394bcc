"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/l2s.h declares, and the ctypes table mirrors the header (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "l2s.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(l2s_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from lang2seg_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libl2s.so does not export %s" % n
    lib.l2s_version.restype = ctypes.c_int
    assert lib.l2s_version() >= 100
    lib.l2s_last_error_string.restype = ctypes.c_char_p
    assert isinstance(lib.l2s_last_error_string(), bytes)


def test_ctypes_table_matches_header():
    from lang2seg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    # argument counts: count top-level commas of each prototype
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), "%s: header has %d parameters, ctypes table %d" % (name, n, len(args))


def test_error_paths_without_gpu():
    """Shape / argument validation happens on the host before any CUDA call."""
    from lang2seg_b200 import _lib
    lib = _lib.load()
    rc = lib.l2s_roi_crop_fwd(None, None, None, None, 1, 8, 16, 16, 1, 7, 0, 0.0, 0.0, None, 0, None)
    assert rc == -5 and b"null" in lib.l2s_last_error_string()
    rc = lib.l2s_att_step_fwd(*([ctypes.c_void_p(16)] * 7), 1, 4, 6, 8, None)
    assert rc == -1 and b"multiples of 4" in lib.l2s_last_error_string()
    with pytest.raises(_lib.L2SError):
        _lib.call("l2s_mask_head_fwd", *([ctypes.c_void_p(16)] * 8), 1, 30, 24, 5, None, 0, None)


def test_cpu_tensors_are_rejected():
    import torch
    import lang2seg_b200.functional as F
    with pytest.raises(AssertionError):
        F.roi_max_pool(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5))


def test_new_entry_points_validate_on_the_host():
    """l2s_embedding_bwd / the L2S_CROP_WS_PREPARED flag: argument checks run before any CUDA call."""
    from lang2seg_b200 import _lib
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    assert lib.l2s_embedding_bwd(p16, p16, p16, 4, 10, 6, None) != 0 and b"bad shape" in lib.l2s_last_error_string()      # D % 4
    assert lib.l2s_embedding_bwd(p16, p16, None, 4, 10, 8, None) != 0 and b"null" in lib.l2s_last_error_string()
    assert lib.l2s_embedding_bwd(p16, ctypes.c_void_p(20), p16, 4, 10, 8, None) != 0 and b"aligned" in lib.l2s_last_error_string()
    # unknown crop flag bits are rejected, the workspace-reuse bit (8) is accepted as a flag (the null pointers then fail)
    rc = lib.l2s_roi_crop_bwd(None, None, None, None, 1, 8, 16, 16, 1, 7, 16, 0.0, 0.0, None, 0, None)
    assert rc != 0
    rc = lib.l2s_roi_crop_bwd(p16, p16, None, p16, 1, 8, 16, 16, 1, 7, 16, 0.0, 0.0, p16, 1 << 20, None)
    assert rc != 0 and b"unknown flags" in lib.l2s_last_error_string()


def test_embedding_module_keeps_the_reference_interface():
    """L2F.Embedding is nn.Embedding for checkpoints and for CPU tensors (only CUDA training lookups use the library)."""
    import torch
    import lang2seg_b200.functional as F
    ref = torch.nn.Embedding(11, 8)
    mod = F.Embedding(11, 8)
    assert list(mod.state_dict()) == list(ref.state_dict()) == ["weight"]
    mod.load_state_dict(ref.state_dict(), strict=True)
    idx = torch.tensor([[1, 2, 2], [0, 10, 3]])
    out = mod(idx)
    assert torch.equal(out, ref(idx))
    out.sum().backward()
    ref(idx).sum().backward()
    assert torch.equal(mod.weight.grad, ref.weight.grad)


def test_header_is_plain_c99(tmp_path):
    """include/l2s.h is the C ABI: it must compile as C (no C++ or torch types in any signature)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "l2s.h"\nint main(void) { return l2s_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
