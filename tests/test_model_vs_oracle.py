"""CPU checks of the two-pass algorithm + host class (linked against the CPU test double of the engine ABI,
tests/native/mock_ncabi.cpp, which compiles the same csrc/step_logic.cuh the kernels do) against the oracle."""
import numpy as np
import pytest

import scenarios
from helpers import same_bits


def test_c1_golden_seed1(mock_host_lib):
    scenarios.c1_golden(mock_host_lib, "c1_seed1_normalised.npz", 3000, check_every=1)


def test_c1_golden_seed2_raw_flags(mock_host_lib):
    scenarios.c1_golden(mock_host_lib, "c1_seed2_raw.npz", 2000, check_every=1)


def test_control_fixture_through_the_host_class(mock_host_lib):
    """The horizon negative control (tie at step 1030) through the host class + the test double: canonical order, explicit
    raster via nc_read_fires, the (model's) nc_state_signature."""
    st, z = scenarios.c1_long_golden(mock_host_lib, "c1_control_h.npz", driver_draws=0)
    assert int(z["horizon"]) == 1030


def test_synthetic_dense_activity(mock_host_lib):
    st = scenarios.synthetic_vs_oracle(mock_host_lib, 800, 50, 500)
    assert st["deliveries"] > 50_000 and st["loads_dropped"] > 0 and st["hidden_rand"] > 0


def test_synthetic_small_dt_and_run_all(mock_host_lib):
    scenarios.synthetic_vs_oracle(mock_host_lib, 300, 30, 400, dt=0.03125, run_all=True)


def test_lazy_mode_with_window_splitting(mock_host_lib):
    scenarios.lazy_vs_oracle(mock_host_lib, 300, 30, 60, dt=0.5)


def test_edge_cases(mock_host_lib):
    scenarios.edge_cases(mock_host_lib)


def test_detector_offsets_reset(mock_host_lib):
    scenarios.detector_and_reset(mock_host_lib)


def test_host_constructor_matches_reference(mock_host_lib, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g, rnet = scenarios.host_constructor_matches_reference(mock_host_lib, have_ref)
    net = g.export_network()
    assert np.array_equal(net["rowptr"], rnet["rowptr"]) and np.array_equal(net["pre"], rnet["pre"])
    assert same_bits(net["weight"], rnet["weight"]) and same_bits(net["length"], rnet["length"])
    assert same_bits(net["positions"], rnet["positions"]) and np.array_equal(net["flag"], rnet["flag"])


def test_checkpoint_resume(mock_host_lib, tmp_path):
    scenarios.checkpoint_resume(mock_host_lib, tmp_path)
