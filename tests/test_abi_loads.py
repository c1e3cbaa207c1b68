"""The C-ABI library loads, exports every symbol include/neucor_b200.h declares, and has no CPU path."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT
from neurocorrelation_b200 import engine


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "neucor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nc_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(engine.ABI_SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol(native_libs):
    L = C.CDLL(native_libs[0])
    for name in declared_symbols():
        assert hasattr(L, name), name


def test_host_library_loads(native_libs):
    L = C.CDLL(native_libs[1])
    for name in ("nch_create", "nch_run", "nch_run_swept", "nch_import_network", "nch_read_neurons"):
        assert hasattr(L, name), name


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="box has a GPU")
def test_no_cpu_fallback_without_a_gpu(native_libs):
    """Without a CUDA device the engine refuses to exist: there is no CPU execution path to fall back to."""
    L = engine.load()
    assert L.nc_device_count() == 0
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine()
    assert "no usable CUDA device" in str(ei.value)
    import neurocorrelation_b200 as nb
    b = nb.NeuCor(0)
    b.create_neuron(0.0, 0.0, 0.0)
    with pytest.raises(nb.NeuCorError):
        b.run()
