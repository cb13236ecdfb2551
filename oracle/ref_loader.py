"""TEST / BASELINE INFRASTRUCTURE ONLY.  Makes the UNMODIFIED reference modules importable where the reference checkout
itself is absent (the GPU box): `build()` — called by __graft_entry__.build() in the build container, where
/root/reference exists — copies the files of the path (modeling_finetune.py, the flash_attention_class.py it imports, and
modeling_pretrain.py for the masked encoder / MAE forward) byte for byte into the git-ignored oracle/_ref/, which travels to the GPU box like the built libstad.so.
`load()` imports modeling_finetune from there behind the same 4-symbol `timm` shim oracle/make_golden.py uses (timm is
not installed; the shim does not touch the forward math).  Nothing in simple-tad_b200/ imports this; only
bench.py's `--impl reference` / cpu_baseline legs and tests/test_model_gpu.py (the live comparison with the reference
executed on the GPU box) do."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_FILES = ("modeling_finetune.py", "flash_attention_class.py", "modeling_pretrain.py")


def build(reference="/root/reference"):
    """Copy the reference files of the path into oracle/_ref/ (no-op when the checkout is absent).  Returns True when
    oracle/_ref/ holds them afterwards."""
    if os.path.isdir(reference):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in REF_FILES:
            src, dst = os.path.join(reference, f), os.path.join(REF_DIR, f)
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in REF_FILES)


def load(module="modeling_finetune"):
    """The reference's modeling_finetune (or modeling_pretrain) module imported from oracle/_ref/, or None (files absent /
    an import of theirs unavailable on this box)."""
    if not all(os.path.exists(os.path.join(REF_DIR, f)) for f in REF_FILES):
        return None
    from . import make_golden
    saved = list(sys.path)
    try:
        make_golden.install_shims(ref_path=REF_DIR)
        import importlib
        return importlib.import_module(module)
    except Exception as e:  # noqa: BLE001  (e.g. flash_attn not importable on a CPU-only box)
        sys.stderr.write(f"oracle/_ref: reference modules not importable here ({e!r}); using the port\n")
        sys.path[:] = saved
        return None
