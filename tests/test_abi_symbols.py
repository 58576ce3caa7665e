"""CPU: the C-ABI library builds, loads and exports every symbol include/pcs_b200.h declares (no compute calls)."""
import ctypes as ct
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from pycamset_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = (ROOT / "include" / "pcs_b200.h").read_text()
    declared = set(re.findall(r"PCS_API\s+[\w\s\*]+?\b(pcs_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    from pycamset_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in pcs_b200.h but not exported by libpcs_b200.so"


def test_chain_names(lib):
    assert lib.pcs_chain_from_name(b"projection_extrinsic3D_template_points") == 0
    assert lib.pcs_chain_from_name(b"projection_extrinsic3D_rigidTform3d_free_point") == 1
    assert lib.pcs_chain_from_name(b"projection_extrinsic3D_free_point") < 0
    assert b"no CPU fallback" in lib.pcs_last_error()


def test_unknown_chain_raises():
    from pycamset_b200 import problem, _lib
    with pytest.raises(_lib.UnknownChainError):
        problem.chain_id_from_blocks(["projection", "extrinsic3D", "my_custom_block"])


def test_struct_sizes_match_header(lib):
    from pycamset_b200 import _lib
    assert ct.sizeof(_lib.ProblemDesc) == 4 * 2 + 8 + 4 * 4 + 8 * 7
    assert ct.sizeof(_lib.ProblemInfo) == 4 * 6 + 8 * 5
    assert ct.sizeof(_lib.LmOptions) == 8 + 6 * 8
    assert ct.sizeof(_lib.LmStats) == 16 + 5 * 8
