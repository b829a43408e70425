"""CPU: the C-ABI library loads, exports every symbol include/wtp_cuda.h declares, the ctypes
mirrors match the C struct layouts, and there is no CPU fallback (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wtp_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wtp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header(pkg, tmp_path):
    """Compile a C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    L = pkg._lib
    structs = {"wtp_force": L.Force, "wtp_spacing": L.Spacing, "wtp_repel_params": L.RepelParams,
               "wtp_repel_result": L.RepelResult, "wtp_trace_entry": L.TraceEntry, "wtp_cloud_metrics": L.CloudMetrics,
               "wtp_timing": L.Timing}
    lines = []
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('printf("\\n");')
    csrc = tmp_path / "layout.c"
    csrc.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "wtp_cuda.h"\nint main(void){' + "".join(lines) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(csrc), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        ct = structs[parts[0]]
        assert int(parts[1]) == C.sizeof(ct), parts[0]
        assert [int(x) for x in parts[2:]] == [getattr(ct, f).offset for f, _ in ct._fields_], parts[0]


def test_header_is_plain_c(tmp_path):
    csrc = tmp_path / "plain.c"
    csrc.write_text('#include "wtp_cuda.h"\nint main(void){return WTP_MAX_K == 256 && WTP_MAX_K_REPEL == 128 ? 0 : 1;}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(csrc), "-o",
                    str(tmp_path / "plain")], check=True)


def test_status_strings_and_version(pkg):
    lib = pkg._lib.load()
    assert lib.wtp_version() >= 100
    assert lib.wtp_status_string(0) == b"ok"
    assert b"fallback" in lib.wtp_status_string(3)


def test_shard_arithmetic_matches_library(pkg):
    lib = pkg._lib.load()
    for n in (0, 1, 7, 1000, 10_000_000, 99_999_999):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = pkg.shard_range(n, r, world)
                assert (b, e) == (lib.wtp_shard_begin(n, r, world), lib.wtp_shard_end(n, r, world))
                assert b == prev and e >= b
                prev = e
            assert prev == n


def test_no_cpu_fallback(pkg):
    """Without a GPU the context cannot be created; with one it can. Either way nothing falls back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(pkg.WtpError):
        pkg.Context(0)
    with pytest.raises(pkg.WtpError):
        pkg.set_topology(pkg.PointCloud(np.random.rand(10, 3)), pkg.KNNTopology, 3)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (tests, smoke and bench only)."""
    pkg_dir = os.path.join(ROOT, "whatsthepoint.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "wtp_oracle" not in text and "import oracle" not in text and "wtpo_" not in text, f
