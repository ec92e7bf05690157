"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol declared in
include/ippl_b200.h, refuses to compute without a device (no CPU fallback), and its host-only layout
logic matches the oracle / the reference's golden tables bit for bit."""
import ctypes as C

import numpy as np
import pytest

import ippl_b200 as ib
import oracle


def test_library_exports_every_declared_symbol():
    syms = ib.exported_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(ib.lib(), s)]
    assert missing == []


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = ib.lib().ipplb_ctx_create(C.byref(h), 0, None, 0)
    assert rc == 3  # IPPLB_ERR_NO_DEVICE
    assert b"no CPU fallback" in ib.lib().ipplb_last_error()
    with pytest.raises(ib.IpplbError):
        ib.Context(0)


def test_layout_matches_reference_golden(golden):
    for (n0, n1, n2, nr, per) in golden["layout_cases"]:
        ng = (int(n0), int(n1), int(n2))
        L = ib.Layout(ng, int(nr), periodic=bool(per))
        assert np.array_equal(L.boxes(), golden[f"boxes_{n0}_{n1}_{n2}_{nr}"])
        for my in range(nr):
            assert np.array_equal(L.neighbors(my), golden[f"nb_{n0}_{n1}_{n2}_{nr}_{per}_{my}"])
        L.close()


@pytest.mark.parametrize("nr", [1, 2, 3, 4, 5, 6, 7, 8, 12, 16])
def test_layout_matches_oracle(nr):
    for ng in [(16, 16, 16), (128, 128, 128), (32, 20, 12), (17, 9, 33), (256, 256, 256)]:
        for par in [(1, 1, 1), (1, 0, 1), (0, 0, 1)]:
            if np.prod([ng[d] if par[d] else 1 for d in range(3)]) < nr:
                continue  # FieldLayout::initialize throws for these
            ob = oracle.partition(ng, nr, par)
            L = ib.Layout(ng, nr, parallel=par)
            assert np.array_equal(L.boxes(), ob)
            origin, h = (0.5, -1.0, 0.0), (0.1, 0.2, 0.3)
            assert np.array_equal(L.regions(origin, h), oracle.regions(ng, ob, origin, h))
            if (ob[:, 3:] - ob[:, :3] + 1).min() >= 2:
                for my in range(nr):
                    assert np.array_equal(L.neighbors(my), oracle.neighbors(ng, ob, my))
            L.close()


def test_layout_rejects_too_many_ranks():
    with pytest.raises(ib.IpplbError):
        ib.Layout((2, 1, 1), 4)  # FieldLayout::initialize throws (FieldLayout.hpp:111-117)


def test_facade_drivers_build_and_refuse_to_run_without_a_gpu(tmp_path):
    """The C++ facade (include/ippl/Ippl.h) and the three drivers compile with the host compiler alone, and a driver
    started without a CUDA device stops with an error instead of computing anything on the CPU."""
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "demos"), "-s"])
    for app in ("LandauDamping", "PenningTrap", "BumponTailInstability"):
        assert os.path.exists(os.path.join(root, "demos", app))
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([os.path.join(root, "demos", "LandauDamping"), "8", "8", "8", "1000", "1", "FFT", "1.0", "LeapFrog"],
                         cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr
    assert not (tmp_path / "data").exists()
