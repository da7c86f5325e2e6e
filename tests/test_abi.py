"""CPU-side checks of the drop-in boundary: libgtars_gpu.so loads and exports every symbol the header declares,
the ctypes binding covers exactly that set, and the library refuses to run without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gtars_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from gtars_b200 import ffi
    syms = _declared_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(ffi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gtars_gpu.h but not exported"


def test_binding_matches_header():
    from gtars_b200 import ffi
    assert sorted(ffi.SIGNATURES) == _declared_symbols()
    ffi.lib()


def test_no_cpu_fallback():
    """Without a GPU, init must fail loudly (GTGPU_ERR_CUDA) rather than fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gtars_b200 import ffi
    with pytest.raises(ffi.GtarsGpuError) as ei:
        ffi.Context(0)
    assert "no CPU fallback" in str(ei.value) or ei.value.code == 2


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gtars_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gtars_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"


def test_gtok_files_match_reference_bytes(golden, tmp_path):
    """K13: the .gtok files under the reference's tests/data/out, byte for byte (gtars-io/src/gtok.rs:126-300).
    Host-only code: runs without a GPU."""
    from gtars_b200 import api
    k = golden[1]["K13_gtok"]
    for name, want in k["tokens"].items():
        raw = bytes.fromhex(k["files"][name]["hex"])
        p = str(tmp_path / name)
        api.write_tokens_to_gtok(p, want)
        assert open(p, "rb").read() == raw, name
        open(p, "wb").write(raw)
        assert api.read_tokens_from_gtok(p) == want
    big = str(tmp_path / "big.gtok")
    api.write_tokens_to_gtok(big, [1, 70000, 3])                     # one token above u16 -> u32 flag
    assert open(big, "rb").read()[:5] == b"GTOK\x02" and api.read_tokens_from_gtok(big) == [1, 70000, 3]
    api.init_gtok_file(big)
    assert open(big, "rb").read() == b"GTOK\x02"
    api.append_tokens_to_gtok_file(big, [5, 0x12345678])
    assert api.read_tokens_from_gtok(big) == [5, 0x12345678]
    small = str(tmp_path / "small.gtok")
    api.write_tokens_to_gtok(small, [7])
    api.append_tokens_to_gtok_file(small, [65537])                  # u16 file: appended tokens are truncated (gtok.rs:278-284)
    assert api.read_tokens_from_gtok(small) == [7, 1]
    bad = str(tmp_path / "bad.gtok")
    open(bad, "wb").write(b"NOPE\x01\x00\x00")
    with pytest.raises(api.GtarsError):
        api.read_tokens_from_gtok(bad)
