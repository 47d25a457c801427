"""The C-ABI shared library loads and exports every symbol include/botgat.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "botgat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(botgat_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("botgat_graph_create", "botgat_graph_destroy", "botgat_graph_get", "botgat_gat_forward",
                 "botgat_gat_backward", "botgat_edge_stage", "botgat_edge_unstage", "botgat_edge_reduce_dst",
                 "botgat_partition_1d", "botgat_last_error", "botgat_abi_version", "botgat_launch_count"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from bot_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libbotgat.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/botgat.h but not exported"


def test_binding_table_matches_header():
    from bot_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.botgat_abi_version() == _lib.ABI_VERSION
    assert lib.botgat_launch_count() >= 0


def test_struct_sizes_match_header():
    """ctypes mirrors of botgat_fwd_args / botgat_bwd_args: compile a probe against the header and compare sizeof."""
    import subprocess
    import tempfile

    from bot_b200 import _lib

    src = '#include "botgat.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(botgat_fwd_args), sizeof(botgat_bwd_args), sizeof(botgat_graph_info));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(_lib.FwdArgs), ctypes.sizeof(_lib.BwdArgs), ctypes.sizeof(_lib.GraphInfo)]


def test_struct_field_offsets_match_header():
    """Every field of the ctypes mirrors sits at the offset the C compiler gives it (names, order and types)."""
    import subprocess
    import tempfile

    from bot_b200 import _lib

    structs = (("botgat_fwd_args", _lib.FwdArgs), ("botgat_bwd_args", _lib.BwdArgs), ("botgat_graph_info", _lib.GraphInfo))
    lines = ['#include "botgat.h"', "#include <stddef.h>", "#include <stdio.h>", "int main(){"]
    for cname, ct in structs:
        for fname, _ in ct._fields_:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write("\n".join(lines))
        exe = os.path.join(d, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    want = [getattr(ct, fname).offset for _, ct in structs for fname, _ in ct._fields_]
    assert got == want


def test_no_cpu_fallback():
    """Compute entry points refuse CPU tensors instead of silently falling back."""
    import pytest
    import torch

    import bot_b200
    from bot_b200.functional import gat_fused

    g = bot_b200.Graph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.in_degrees()
    with pytest.raises(RuntimeError, match="no CPU path"):
        gat_fused(g, torch.zeros(2, 1, 4), torch.zeros(2, 1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.remove_self_loop()


def test_product_does_not_import_the_oracle():
    """Nothing under bot_b200/ may import oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "bot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
