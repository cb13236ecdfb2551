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
    # stad_input: pointer + 5 x int32 (+ 4 bytes of padding) + pointer (ABI v6: window_starts); stad_dims: 11 x int32;
    # stad_block: 10 pointers; stad_outputs: 4 pointers
    assert ctypes.sizeof(lib.StadInput) == 40
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


def test_argument_errors_are_reported_before_any_device_work(lib):
    """Error convention of the boundary: bad shapes / alignment return STAD_E_SHAPE / STAD_E_ALIGN with a
    message and are mapped to ValueError — decided on the host, so this holds without a GPU."""
    l = lib.load()
    buf = (ctypes.c_uint8 * 256)()
    addr = (ctypes.addressof(buf) + 15) & ~15          # 16-byte aligned host address standing in for a device pointer
    p, odd = ctypes.c_void_p(addr), ctypes.c_void_p(addr + 2)
    # rows_norm_head: empty problem, misaligned rows, a head without logits
    assert l.stad_rows_norm_head(p, p, p, p, p, p, None, None, 0, 1, 0, 768, 2, 1e-6, None) == lib.STAD_E_SHAPE
    assert l.stad_rows_norm_head(odd, p, p, p, p, p, None, None, 4, 1, 0, 768, 2, 1e-6, None) == lib.STAD_E_ALIGN
    assert l.stad_rows_norm_head(p, p, p, p, p, None, None, None, 4, 1, 0, 768, 2, 1e-6, None) == lib.STAD_E_SHAPE
    assert "rows_norm_head" in lib.last_error()
    assert l.stad_rows_norm_head(p, p, p, p, p, p, None, None, 4, 0, 0, 768, 2, 1e-6, None) == lib.STAD_E_SHAPE
    # prepend_cls: empty batch, width that is not a multiple of 8
    assert l.stad_prepend_cls(p, p, p, p, 0, 1568, 768, 1e-6, None) == lib.STAD_E_SHAPE
    assert l.stad_prepend_cls(p, p, p, p, 2, 1568, 770, 1e-6, None) == lib.STAD_E_SHAPE
    # frame-buffer input: the last clip must fit the resident frames, also with an in-window frame step
    dims = lib.make_dims(dim=384, depth=1, heads=6, hidden=1536, num_classes=2)
    ok = lib.StadInput(addr, lib.STAD_IN_FRAMES, 46 + 5, 0, 5, 3)    # 2 clips: (16 - 1) * 3 + 1 = 46 frames each, 5 apart
    short = lib.StadInput(addr, lib.STAD_IN_FRAMES, 46 + 4, 0, 5, 3)
    rc = l.stad_patch_embed(ctypes.byref(short), p, p, None, p, None, ctypes.byref(dims), 2, 1568, None)
    assert rc == lib.STAD_E_SHAPE and "needs frame 50" in lib.last_error()
    neg = lib.StadInput(addr, lib.STAD_IN_FRAMES, 100, 0, 1, -1)
    assert l.stad_patch_embed(ctypes.byref(neg), p, p, None, p, None, ctypes.byref(dims), 2, 1568, None) == lib.STAD_E_SHAPE
    with pytest.raises(ValueError, match="frame_step"):
        lib.check(l.stad_patch_embed(ctypes.byref(neg), p, p, None, p, None, ctypes.byref(dims), 2, 1568, None), "patch_embed")
    del ok
    # vit_forward: unknown reduction, class token together with a visible-token list
    m = lib.StadModel()
    m.dims = dims
    blocks = (lib.StadBlock * 1)()
    m.blocks = ctypes.cast(blocks, ctypes.POINTER(lib.StadBlock))
    m.norm_g = m.norm_b = m.w_head = m.b_head = addr
    m.reduction = 7
    outs = lib.StadOutputs(addr, None, None, None)
    inp = lib.StadInput(addr, lib.STAD_IN_CLIPS, 0, 0, 1, 1)
    ws = ctypes.c_void_p((addr + 255) & ~255)
    rc = l.stad_vit_forward(ctypes.byref(m), ctypes.byref(inp), None, 1, 1568, ctypes.byref(outs), ws, 1 << 40, None)
    assert rc == lib.STAD_E_SHAPE and "reduction" in lib.last_error()
    m.reduction = lib.STAD_REDUCE_MEAN
    m.cls_token = addr
    rc = l.stad_vit_forward(ctypes.byref(m), ctypes.byref(inp), p, 1, 160, ctypes.byref(outs), ws, 1 << 40, None)
    assert rc == lib.STAD_E_SHAPE and "class token" in lib.last_error()
