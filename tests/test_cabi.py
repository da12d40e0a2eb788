"""The C-ABI library builds, loads and exports every symbol include/tfem_b200.h declares.
No compute calls are made (there is no GPU in the CPU test run)."""

import ctypes
import os

from pytorch_fem_solver_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _lib.header_symbols()
    assert len(declared) >= 33
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # the Python binding table covers exactly the header
    assert sorted(declared) == sorted(_lib.exported_symbols())


def test_status_strings_and_version():
    lib = _lib.load()
    assert lib.tfem_abi_version() == 1
    assert lib.tfem_status_string(0) == b"ok"
    assert b"bad argument" in lib.tfem_status_string(-1)


def test_argument_validation_happens_before_any_launch():
    """NULL pointers / bad sizes are rejected on the host side, so no GPU is needed."""
    lib = _lib.load()
    assert lib.tfem_tri_p1_geometry_f64(-1, 1, 1, None, None, 3, None, None, None, None, None, None, None, None, None) == -1
    assert lib.tfem_tri_p1_geometry_f64(4, 4, 4, None, None, 3, None, None, None, None, None, None, None, None, None) == -1
    assert lib.tfem_tri_p1_geometry_f64(0, 1, 1, None, None, 3, None, None, None, None, None, None, None, None, None) == 0
    assert lib.tfem_scatter_bilinear_f32(5, None, None, None, None, None) == -1
    assert lib.tfem_quad_reduce_f64(3, 0, 9, None, 0, 0, None, None, None) == -1


def test_product_has_no_cpu_path():
    """Without CUDA the basis constructor must fail loudly, not fall back."""
    import pytest
    import torch

    import pytorch_fem_solver_b200 as tfem

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    mesh = tfem.MeshTri(tfem.meshgen.structured_rectangle(2, 2))
    with pytest.raises(_lib.TfemError):
        tfem.Basis(mesh, tfem.ElementTri(1, 2))


def _header_signatures():
    """{base name: [C parameter type strings]} parsed from include/tfem_b200.h (macro and plain)."""
    import re

    with open(_lib.HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S).replace("\\\n", " ")
    found = {}
    for match in re.finditer(r"(?:int|const char\*)\s+(tfem_\w+?)(_##SUF)?\s*\(([^)]*)\)\s*;", text):
        name, _, params = match.groups()
        params = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        found[name] = params
    return found


def test_binding_arity_and_kinds_match_the_header():
    """A ctypes argtypes list shorter than the C parameter list lets the extra argument through with the
    default (32-bit int) conversion: pointers such as the stream then arrive truncated.  Every entry of
    the binding tables must have exactly one argtype per C parameter, pointer for pointer."""
    signatures = _header_signatures()
    tables = {base: argtypes for base, argtypes in _lib._TYPED.items()}
    tables.update({name: argtypes for name, (argtypes, _) in _lib._UNTYPED.items()})
    assert set(tables) == set(signatures)
    for name, params in signatures.items():
        argtypes = tables[name]
        assert len(argtypes) == len(params), f"{name}: {len(argtypes)} argtypes for {len(params)} C parameters"
        for ctype, param in zip(argtypes, params):
            is_pointer = "*" in param
            bound_pointer = ctype in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ctype, "_type_") and isinstance(ctype._type_, type)
            assert is_pointer == bool(bound_pointer), f"{name}: parameter '{param}' bound as {ctype}"
