"""The C-ABI shared library loads and exports every symbol include/pb254.h declares (no compute calls:
there is no GPU in the CPU test tier), and the product binding has no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pb254.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb254_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("pb254_generate_trace", "pb254_prove", "pb254_prove_dev", "pb254_prove_trace", "pb254_verify",
                 "pb254_proof_data", "pb254_proof_free", "pb254_ctx_create", "pb254_last_error",
                 "pb254_config_standard_fast"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from plonky2_bn254_b200 import build, ffi
    build.build_cuda()
    lib = ctypes.CDLL(ffi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_shapes_through_the_abi():
    from plonky2_bn254_b200 import ffi
    L = ffi.Library()
    assert [L.trace_width(k) for k in (0, 1, 2)] == [781, 1295, 427]
    assert [L.input_words(k) for k in (0, 1, 2)] == [20, 36, 8]
    assert [L.num_aux(k) for k in (0, 1, 2)] == [456, 906, 134]
    assert L.trace_rows(1, 1 << 16) == 1 << 16 and L.trace_rows(1024, 1 << 16) == 1 << 19
    assert L.trace_rows(129, 1 << 16) == 1 << 17
    assert L.standard_fast_config().as_tuple() == (1, 4, 2, 84, 16, 4, 5)


def test_no_cpu_fallback():
    """Without a GPU the product context must fail loudly (CUDA error), not fall back to anything."""
    import torch
    from plonky2_bn254_b200 import ffi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ffi.Pb254Error) as e:
        ffi.Context(0)
    assert e.value.code == 4


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "plonky2_bn254_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "oracle/" not in text.replace("the oracle", ""), f


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors of the header's structs (ffi.Config, ffi.Comm, ffi.ProofLayout) have the C compiler's sizes
    and field offsets: a tiny C program including include/pb254.h prints them."""
    import subprocess
    from plonky2_bn254_b200 import ffi
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pb254.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu\\n", sizeof(pb254_config), offsetof(pb254_config, num_query_rounds),'
                   ' offsetof(pb254_config, final_poly_bits), sizeof(pb254_proof_layout));\n'
                   '  printf("%zu %zu %zu %zu %zu\\n", sizeof(pb254_comm), offsetof(pb254_comm, world), offsetof(pb254_comm, user),'
                   ' offsetof(pb254_comm, all_to_all), offsetof(pb254_comm, all_gather));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    cfg, comm = ffi.Config, ffi.Comm
    assert [int(x) for x in out[:4]] == [ctypes.sizeof(cfg), cfg.num_query_rounds.offset, cfg.final_poly_bits.offset,
                                         ctypes.sizeof(ffi.ProofLayout)]
    assert [int(x) for x in out[4:]] == [ctypes.sizeof(comm), comm.world.offset, comm.user.offset,
                                         comm.all_to_all.offset, comm.all_gather.offset]
