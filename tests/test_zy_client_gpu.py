"""GPU half of tests/test_boundary_client.py: the boundary client — main.cpp's presets plus a renderer frame over the object graph —
linked against the real libraries (host class -> C ABI -> CUDA engine) must print the lines the reference printed
(tests/golden/client_presets.txt).  Runs late (file name): the renderer-frame walk was added after the round's GPU budget was
spent, so the parity tests proper are not held up by it."""
import pytest

from test_boundary_client import GOLDEN, _build_dropin_client, _run_all

pytestmark = pytest.mark.gpu


def test_client_against_the_cuda_engine(native_libs):
    got = _run_all(_build_dropin_client(mock=False))
    assert got == open(GOLDEN).read()
