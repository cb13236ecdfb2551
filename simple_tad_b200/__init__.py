"""Importable alias of the `simple-tad_b200/` package directory (a hyphen is not a valid module name).

`import simple_tad_b200.modeling_finetune` resolves to `simple-tad_b200/modeling_finetune.py`.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "simple-tad_b200")
if not _os.path.isdir(_impl):  # pragma: no cover
    raise ImportError("simple-tad_b200/ package directory not found next to simple_tad_b200/")
__path__.append(_impl)

__version__ = "0.1.0"
