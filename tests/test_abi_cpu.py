"""CPU-only checks of the drop-in boundary: libvkgpu.so loads, exports every symbol include/vkgpu.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(built):
    from valkey_search_b200 import _lib as L
    header = open(os.path.join(ROOT, "include", "vkgpu.h")).read()
    declared = set(re.findall(r"^(?:int|void|uint32_t|uint64_t|const char \*|vkgpu_index \*)\s*(vkgpu_[a-z0-9_]+)\(", header, re.M))
    assert declared, "no declarations parsed"
    lib = C.CDLL(L.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in vkgpu.h but not exported: {missing}"
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    assert L.lib().vkgpu_abi_version() == 1


def test_struct_layout_matches_header(built):
    from valkey_search_b200 import _lib as L
    assert C.sizeof(L.Config) == 56
    assert C.sizeof(L.Filter) == 40
    assert C.sizeof(L.Stats) == 104


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import valkey_search_b200 as V
    with pytest.raises(V.VkgpuError) as ei:
        V.VectorFlat(8)
    assert ei.value.code == 4 and "no CPU fallback" in ei.value.message


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "valkey_search_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "vk_oracle" not in text and "libvkoracle" not in text and "libvkref" not in text, f


def test_packed_result_layout_size(built):
    """vkgpu_packed_result_bytes: labels u64 [B][k] | dist f32 [B][k] | n u32 [B], padded to 256 bytes (pure host)."""
    from valkey_search_b200 import _lib as L
    lib = L.lib()
    for B, k in ((1, 1), (33, 25), (1024, 100), (512, 10)):
        n = int(lib.vkgpu_packed_result_bytes(B, k))
        assert n % 256 == 0 and 0 <= n - (B * k * 12 + B * 4) < 256
