"""The reference-vector tests of tests/test_audit.py (LimitQueue / SVQueue.Less / truncation, the process excluder, audit-from-cache
scenarios, getValidationMessages, the admission coalescer) re-run against the CUDA library instead of the test backend."""
import pytest

import test_audit as TA
from conftest import has_cuda

pytestmark = pytest.mark.gpu
_NAMES = sorted(n for n in dir(TA) if n.startswith("test_") and callable(getattr(TA, n)))


@pytest.mark.parametrize("name", _NAMES)
def test_audit_vectors_on_the_cuda_library(name, monkeypatch):
    if not has_cuda():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    monkeypatch.setattr(TA, "LIB", None)   # None = gatekeeper_b200/libgk_engine.so
    getattr(TA, name)()
