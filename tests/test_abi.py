"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/sloth_b200.h declares; the host helpers agree with the oracle; and without
a GPU the product fails loudly instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import rust_sloth_b200 as rs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sloth_b200.h")).read()
    return sorted(set(re.findall(r"SLOTH_API\s+[^;(]*?\b(sloth_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = rs.load_library()
    names = header_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/sloth_b200.h but not exported"
    assert sorted(rs.ABI_SYMBOLS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", rs.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    # nothing but the ABI leaks out of the library
    assert all(e.startswith("sloth_") for e in exported), exported - set(names)


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", rs.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rs.SlothError) as e:
        rs.Context.blank(True)
    assert e.value.code == rs.SLOTH_E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rust-sloth_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "import oracle" not in text, os.path.join(dirpath, f)


def test_host_helpers_match_the_oracle():
    for (r, p, y) in [(0.0, np.pi, 0.0), (0.3, 3.3, -0.7), (1e-3, 9.4, 2.0), (-2.0, 0.1, 7.0)]:
        assert np.array_equal(rs.rotation_from_euler(r, p, y), oracle.rotation(r, p, y))
    for (W, H, s) in [(80, 40, 7.534395), (1920, 1080, 0.958196), (3840, 2160, 1.0), (101, 57, 1.367188), (65536, 65536, 1.0), (7, 200, 0.0)]:
        a, b = rs.utransform(W, H, s), oracle.utransform(W, H, s)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for (y0, n) in [(0.0, 360), (0.0, 64), (0.0, 1), (5.0, 100), (-1.0, 7)]:
        assert np.array_equal(rs.turntable_pitches(y0, n), oracle.turntable(y0, n))


def test_default_shader_table():
    assert "".join(rs.default_shader(s) for s in [0.0, 0.2, 0.25, 0.3, 0.45, 0.55, 0.65, 0.75, 0.85, 0.95, 1.0]) == "..::=+*#%@@"
    assert rs.default_shader(1.0001) == " " and rs.default_shader(float("nan")) == " " and rs.default_shader(-5) == "."


def test_flush_formats():
    cells = np.array([ord("@") | 1 << 8 | 2 << 16 | 3 << 24, ord("\n"), ord(" ")], np.uint32)
    assert rs.flush_bytes(cells, False, False, True) == b"@\n \n"
    assert rs.flush_bytes(cells, True, True, True) == (b'<span style="color:rgb(1,2,3)">@<span style="color:rgb(0,0,0)">\n'
                                                      b'<span style="color:rgb(0,0,0)"> ')
    ansi = rs.flush_bytes(cells[:1], True, False, False)
    assert ansi == b"\x1b[1;1H\x1b[48;2;25;25;25m\x1b[38;2;1;2;3m@\x1b[0m"


def test_integration_md_shows_a_binding_for_every_entry_point():
    """INTEGRATION.md is the reference-side binding a maintainer would add: nothing the header declares is missing."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    integ = open(os.path.join(root, "INTEGRATION.md")).read()
    assert sorted(s for s in rs.ABI_SYMBOLS if s not in integ) == []   # ABI_SYMBOLS == the header, see above


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/sloth_b200.h must be consumable by a C compiler (the boundary is a C ABI, no C++ or torch types), and
    a C caller must link: it calls the host-side helpers and checks that context creation fails cleanly without a GPU
    (or succeeds with one)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "sloth_b200.h"
int main(void) {
    float rot[16], ut[16], pitches[8];
    sloth_rotation_from_euler(0.0f, 3.14159274f, 0.0f, rot);
    sloth_utransform(80, 40, 7.5f, ut);
    size_t n = sloth_turntable_pitches(0.0f, 4, pitches, 8);
    sloth_ctx *ctx = NULL;
    int rc = sloth_ctx_create(0, 1, &ctx);
    if (rc == SLOTH_OK) { rc = sloth_ctx_destroy(ctx); if (rc != SLOTH_OK) return 3; }
    else if (rc != SLOTH_E_CUDA || strlen(sloth_last_error()) == 0) return 2;
    printf("%zu %.1f %.1f\n", n, (double)ut[12], (double)ut[13]);
    return 0;
}
''')
    exe = tmp_path / "caller"
    lib_dir = os.path.dirname(rs.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lsloth_b200", f"-Wl,-rpath,{lib_dir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["4", "20.0", "20.0"]      # W/4 and H/2 of Context::update's matrix (context.rs:115-132)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_cli_fails_loudly_without_a_gpu(tmp_path):
    """The drop-in `sloth` binary has no CPU path either: without a device it must exit non-zero with the CUDA error."""
    exe = os.path.join(os.path.dirname(rs.LIB_PATH), "bin", "sloth")
    obj = tmp_path / "t.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    r = subprocess.run([exe, str(obj), "image", "-w", "20", "-h", "10"], capture_output=True, text=True)
    assert r.returncode != 0 and r.stdout == ""
    assert "cuda" in r.stderr.lower()


def test_cli_surface_follows_the_reference_clap_definition():
    """src/inputs.rs:9-85: required arguments, help/version, unknown flags -- all decided before
    any GPU work, so this runs anywhere."""
    exe = os.path.join(os.path.dirname(rs.LIB_PATH), "bin", "sloth")
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True)
    r = run()
    assert r.returncode == 1 and "required arguments were not provided" in r.stderr and "<input filename(s)>" in r.stderr
    r = run("model.obj", "image")
    assert r.returncode == 1 and "-w <width>" in r.stderr
    r = run("model.obj", "image", "-w")
    assert r.returncode == 1 and "requires a value" in r.stderr
    r = run("model.obj", "--frobnicate")
    assert r.returncode == 1 and "wasn't expected" in r.stderr
    r = run("--help")
    assert r.returncode == 0 and r.stdout.startswith("Sloth 0.1") and "image -w <width> [-h <height>] [-j, --webify <frame count>]" in r.stdout
    r = run("-V")
    assert r.returncode == 0 and r.stdout.strip() == "Sloth 0.1"


def test_span_expansion_on_the_host():
    """sloth_expand_spans (the host half of SLOTH_WIRE_SPANS, no GPU involved): run lists against np.repeat, from
    single-cell runs to runs long enough for the streaming-store path, at every alignment of the destination."""
    rng = np.random.default_rng(11)
    for n_cells, n_runs in [(1, 1), (7, 3), (1000, 1), (1000, 1000), (100_003, 37), (3840 * 64, 900), (1 << 20, 5)]:
        starts = np.sort(rng.choice(np.arange(1, n_cells), size=min(n_runs, n_cells) - 1, replace=False)) if n_cells > 1 else np.array([], np.int64)
        starts = np.concatenate([[0], starts]).astype(np.uint32)
        vals = rng.integers(0, 1 << 32, size=starts.size, dtype=np.uint64).astype(np.uint32)
        want = np.repeat(vals, np.diff(np.concatenate([starts, [n_cells]]).astype(np.int64)))
        got = rs.expand_spans(np.stack([starts, vals], axis=1), n_cells)
        assert np.array_equal(got, want), (n_cells, n_runs)
    with pytest.raises(rs.SlothError):
        rs.expand_spans(np.array([[1, 5]], np.uint32), 10)            # must start at cell 0
    with pytest.raises(rs.SlothError):
        rs.expand_spans(np.array([[0, 5], [4, 6], [4, 7]], np.uint32), 10)   # starts must ascend
    with pytest.raises(rs.SlothError):
        rs.expand_spans(np.array([[0, 5], [10, 6]], np.uint32), 10)   # inside the frame


def test_bench_traffic_figures_are_committed_and_labelled():
    """bench.py reports `roofline.traffic` from profiles/traffic.json (DRAM bytes cannot be measured outside ncu): the
    figures of both kernels it reports must be there, labelled with the capture they came from, and that capture's
    summary must be a tracked file."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    for key in ("k_tri_dram_bytes_per_launch", "k_resolve_even_dram_bytes_per_launch", "k_geom3_dram_bytes_per_launch"):
        assert isinstance(t[key], int) and t[key] > 0, key
    assert "ncu --set full" in t["capture"] and t["commit"]
    summary = t["capture"].split("(")[1].rstrip(")")
    assert os.path.exists(os.path.join(root, summary)), summary
    # algorithmic bytes of the bench workload: the dominant kernel must not move more than it is credited with
    assert t["k_tri_dram_bytes_per_launch"] < 40 * 10_025_280
