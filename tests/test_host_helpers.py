"""CPU tests of the C ABI that need no GPU: the library loads, exports every symbol the
header declares, and its host helpers restate pandas / numpy semantics bit-exactly."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from lorenzcycletoolkit_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as G
    G.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lec_b200.h")).read()
    names = set(re.findall(r"\b(lec_[a-z0-9_]+)\s*\(", hdr))
    assert {"lec_create", "lec_destroy", "lec_run_device", "lec_run_host", "lec_gradient_coefs",
            "lec_nearest_index", "lec_last_timing", "lec_timing_reset", "lec_launch_count", "lec_last_transfer", "lec_run_host_raw",
            "lec_diag850_host", "lec_diag850_device", "lec_strerror",
            "lec_last_error", "lec_version"} <= names
    lib = ctypes.CDLL(str(E.library_path()))
    for n in names:
        assert hasattr(lib, n), n
    assert "sm_100a" in E.version()


def test_torch_extension_loads_and_registers_its_operator():
    """csrc/lec_torch_ext.cpp (the thin PyTorch C++ extension over the C ABI) is built in-tree by build(), links to
    the in-tree liblec_b200.so and registers ``torch.ops.lec_b200.run_device`` for the CUDA dispatch key only --
    without a GPU the operator refuses CPU tensors instead of computing anything."""
    import torch
    if "LEC_B200_LIB" in os.environ or os.environ.get("LEC_TORCH_EXT") == "0":
        pytest.skip("extension disabled by the environment")
    assert E.load_torch_extension()
    op = torch.ops.lec_b200.run_device
    assert "Tensor[] fields" in str(op.default._schema) and str(op.default._schema).endswith("-> int")
    z = torch.zeros((1, 1, 2, 2))
    with pytest.raises((RuntimeError, NotImplementedError)):
        op(1, [z] * 5, torch.zeros(56, dtype=torch.uint8), torch.zeros((1, 16), dtype=torch.float64), None, None)


def _build_abi_check(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(str(E.library_path()))
    subprocess.run(["gcc", "-std=c11", "-D_GNU_SOURCE", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe, "-L", libdir, "-llec_b200", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """include/lec_b200.h compiles as C11 with -Wall -Werror; a C program linked against the library (no
    Python, no torch) gets the host helpers' answers and LEC_ERR_CUDA from lec_create without a GPU."""
    import subprocess
    import torch
    exe = _build_abi_check(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert "no GPU: lec_create -> CUDA runtime failure" in out.stdout


@pytest.mark.gpu
def test_c_program_runs_both_host_paths(tmp_path):
    """The same C program on a GPU: lec_run_host and lec_run_host_raw (records stored north to south) give
    identical bits."""
    import subprocess
    exe = _build_abi_check(tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    assert "launches" in out.stdout


def test_step_struct_layout_matches_header():
    assert E.STEP_DTYPE.itemsize == 56
    assert E.STEP_DTYPE.fields["ct_m"][1] == 32 and E.STEP_DTYPE.fields["j1"][1] == 24


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nearest_index_is_pandas_nearest(dtype):
    rng = np.random.default_rng(0)
    for coord in (np.arange(-90, 90.1, 2.5), np.arange(-180, 180, 0.25), np.sort(rng.uniform(-50, 50, 37))):
        coord = coord.astype(dtype)
        idx = pd.Index(coord)
        vals = np.concatenate([rng.uniform(coord[0] - 3, coord[-1] + 3, 300),
                               0.5 * (coord[1:].astype(np.float64) + coord[:-1].astype(np.float64)),   # exact ties
                               coord.astype(np.float64)])
        want = idx.get_indexer(vals, method="nearest")
        got = np.array([E.nearest_index(coord, v) for v in vals])
        assert np.array_equal(got, want)


def test_nearest_index_ties_go_to_larger_coordinate():
    lat25 = np.arange(-90, 90.1, 2.5)
    assert lat25[E.nearest_index(lat25, -31.25)] == -30.0
    lat025 = np.arange(-90, 90.01, 0.25)
    assert lat025[E.nearest_index(lat025, -30.125)] == -30.0


@pytest.mark.parametrize("x", [np.arange(0, 10) * 21600.0,                      # uniform time axis
                               np.array([1000., 2000, 3000, 5000, 7000, 10000, 15000, 20000]),
                               np.deg2rad(np.arange(-35, -19, 2.5, dtype=np.float32)).astype(np.float64)])
def test_gradient_coefs_reproduce_np_gradient(x):
    rng = np.random.default_rng(1)
    f = rng.normal(size=x.size)
    a, b, c = E.gradient_coefs(x)
    fm, fp = np.roll(f, 1), np.roll(f, -1)
    got = a * fm + b * f + c * fp
    want = np.gradient(f, x, edge_order=1)
    assert np.allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(want).max())


def test_gradient_needs_two_points():
    with pytest.raises(ValueError):
        E.gradient_coefs([1.0])


def test_time_stencil_slots_and_edges():
    tsec = np.arange(5) * 21600.0
    st = E.time_stencil(tsec, E.make_steps(5))
    assert list(st["slot_m"]) == [0, 0, 1, 2, 3] and list(st["slot_p"]) == [1, 2, 3, 4, 4]
    assert st["ct_m"][0] == 0 and st["ct_p"][-1] == 0
    assert np.isclose(st["ct_p"][2], 1 / 43200.0) and np.isclose(st["ct_0"][0], -1 / 21600.0)


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = np.arange(4.0)
    with pytest.raises(RuntimeError):
        E.LecEngine(g, g, np.deg2rad(g), np.deg2rad(g), np.cos(np.deg2rad(g)), [1e4, 5e4, 1e5], np.float32)


def test_multiply_shift_division_is_exact_below_2_31():
    """The row kernels decode the block / tile index with ``FastDiv`` (csrc/lec_row_moments.cuh): q = (n * mul) >> shr
    with mul = floor(2^(31+L) / d) + 1, shr = 31 + L, L = ceil(log2 d).  Exact for every numerator below 2^31 (the
    engine refuses larger grids) and every divisor the host can produce; mul must fit 32 bits."""
    rng = np.random.default_rng(7)

    def make(d):
        L = 0
        while (1 << L) < d:
            L += 1
        return ((1 << (31 + L)) // d) + 1, 31 + L

    divisors = list(range(1, 300)) + [int(x) for x in rng.integers(1, 2**31, 3000)] + [2**k for k in range(31)] + [2**31 - 1]
    for d in divisors:
        mul, shr = make(d)
        assert mul < 2**32
        ns = np.concatenate([[0, 1, d - 1, d, min(d + 1, 2**31 - 1), 2**31 - 1, 2**31 - 2],
                             rng.integers(0, 2**31, 40)]).astype(object)
        for n in ns:
            assert (int(n) * mul) >> shr == int(n) // d, (n, d)
