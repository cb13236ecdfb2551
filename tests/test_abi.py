"""CPU: the C-ABI library loads and exports exactly the symbols include/stad.h declares; the ctypes structs agree with
the header; nothing computes without a GPU and nothing silently falls back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stad.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from simple_tad_b200 import _lib
    return _lib


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"STAD_API\s+[\w\s\*]+?\b(stad_\w+)\s*\(", text)))


def test_header_symbols_are_all_exported(lib):
    names = _declared()
    assert len(names) >= 13
    handle = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/stad.h but not exported by libstad.so"
    assert sorted(lib.EXPORTS) == names, "simple-tad_b200/_lib.py binds a different set of entry points than the header"


def test_abi_version_and_error_string(lib):
    l = lib.load()
    text = open(HEADER).read()
    assert l.stad_abi_version() == int(re.search(r"#define STAD_ABI_VERSION (\d+)", text).group(1))
    assert isinstance(lib.last_error(), str)


def test_struct_layouts_match_header(lib):
    # stad_input: pointer + 5 x int32 (+ 4 bytes of tail padding); stad_dims: 11 x int32; stad_block: 10 pointers;
    # stad_outputs: 4 pointers
    assert ctypes.sizeof(lib.StadInput) == 32
    assert ctypes.sizeof(lib.StadDims) == 44
    assert ctypes.sizeof(lib.StadBlock) == 80
    assert ctypes.sizeof(lib.StadOutputs) == 32
    text = open(HEADER).read()
    for struct, cls in (("stad_dims", lib.StadDims), ("stad_block", lib.StadBlock), ("stad_input", lib.StadInput),
                        ("stad_outputs", lib.StadOutputs), ("stad_model", lib.StadModel)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
        assert fields == [f[0] for f in cls._fields_], f"{struct}: header fields {fields} != ctypes fields"


def test_workspace_bytes_is_pure_host_arithmetic(lib):
    dims = lib.make_dims(dim=768, depth=12, heads=12, hidden=3072, num_classes=2)
    full = lib.load().stad_workspace_bytes(ctypes.byref(dims), 4, 1568)
    M = 4 * 1568
    assert full >= M * 768 * 2 + M * 8 + M * 3072 * 2 + 4 * 16 * 768 * 4
    assert full < 1.05 * (M * 768 * 2 + M * 8 + M * 3072 * 2 + 4 * 16 * 768 * 4) + 4096
    masked = lib.load().stad_workspace_bytes(ctypes.byref(dims), 4, 160)
    assert masked >= 4 * 160 * 1536 * 2  # visible-token path adds the im2col scratch
    assert lib.load().stad_workspace_bytes(ctypes.byref(dims), 0, 1568) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    with pytest.raises(RuntimeError):
        lib.init()
    rc = lib.load().stad_init(0)
    assert rc < 0 and lib.last_error()
    from simple_tad_b200 import modeling_finetune as mf
    m = mf.vit_small_patch16_224(num_classes=2).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 16, 224, 224))


def test_missing_library_raises(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist", "libstad.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        lib.load()


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke may touch oracle/."""
    pkg = os.path.join(ROOT, "simple-tad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle/"
