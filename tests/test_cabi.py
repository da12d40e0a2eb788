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


def test_argument_validation_of_the_round_two_entry_points():
    """Symbolic phase, edge topology, MLP producer, tiled residual: bad arguments come back as status codes from the
    host side (no launch, no GPU needed)."""
    import ctypes

    lib = _lib.load()
    need = ctypes.c_int64(-1)
    assert lib.tfem_csr_symbolic_workspace(-1, 10, ctypes.byref(need)) == -1
    assert lib.tfem_csr_symbolic_workspace(10, 10, None) == -1
    assert lib.tfem_csr_symbolic_workspace(2**28, 10, ctypes.byref(need)) == -4  # 9 n_el exceeds the 32-bit index range
    assert lib.tfem_csr_symbolic(0, None, 10, None, 0, None, None, None, None, None, None, None, None, None) == -1
    assert lib.tfem_csr_symbolic(4, None, 10, None, 0, None, None, None, None, None, None, None, None, None) == -1
    assert lib.tfem_half_edges_workspace(-1, 1, ctypes.byref(need)) == -1
    assert lib.tfem_half_edges(1, 4, 9, None, None, None, 0, None, None, None, None, None, None) == -1
    assert lib.tfem_edge_cells(1, 4, 9, None, 3, None, None, 12, None, None, None) == -1  # n_sides is 1 or 2
    assert lib.tfem_edge_cells(1, 0, 9, None, 2, None, None, 12, None, None, None) == 0  # no edges: nothing to do
    assert lib.tfem_interior_edge_geometry_f64(1, 4, 9, 8, None, None, None, None, None, None, None, None) == -1
    assert lib.tfem_interior_edge_geometry_f32(1, 0, 9, 8, None, None, None, None, None, None, None, None) == 0
    assert lib.tfem_mlp_value_grad_f64(0, 2, 8, 1, 0, None, None, None, None, None) == 0
    assert lib.tfem_mlp_value_grad_f64(5, 2, 8, 1, 0, None, None, None, None, None) == -1
    assert lib.tfem_mlp_value_grad_f32(-1, 2, 8, 1, 0, None, None, None, None, None) == -1
    assert lib.tfem_mlp_value_grad_bwd_f64(5, 2, 8, 1, 0, None, None, None, None, None, 4, None, None) == -1
    fake = ctypes.addressof(ctypes.create_string_buffer(64))  # never dereferenced: the checks below fail first
    assert lib.tfem_mlp_value_grad_f64(5, 2, 40, 1, 0, fake, fake, fake, fake, None) == -2  # wider than one lane per neuron
    assert lib.tfem_mlp_value_grad_f64(5, 2, 8, 1, 7, fake, fake, fake, fake, None) == -2  # unknown activation
    assert lib.tfem_mlp_value_grad_f64(5, 4, 8, 1, 0, fake, fake, fake, fake, None) == -1  # d is 1..3
    assert lib.tfem_mlp_value_grad_bwd_f64(5, 2, 8, 9, 0, fake, fake, fake, fake, fake, 4, fake, None) == -2  # more than 7 square layers
    assert lib.tfem_weak_residual_tiled_f64(None, None, 3, None, None, 0, None, None, None, None) == -1
    assert lib.tfem_tri_p1_assemble_csr_ex_f64(None, None, 3, None, None, None, 0, None, None, None, None) == -1


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
