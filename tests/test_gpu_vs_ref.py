"""`-m gpu`: the CUDA path checked DIRECTLY against the reference's own CPU kernels (oracle/_ref/libafref.so: the templates of
/root/reference/src/backend/cpu/kernel/{reduce,ireduce,scan,scan_by_key}.hpp compiled in place by oracle/Makefile) at the sizes
of the big-shape parity tests -- not only through the restated oracle, which tests/test_oracle_vs_ref.py pins to those kernels on
small shapes.  The test bodies are the ones of tests/test_gpu_parity.py with the checker swapped; skipped where the prebuilt
oracle/_ref did not travel (it is git-ignored, not gpurun-ignored)."""
import pytest

from oracle import pyoracle as po
from tests import test_gpu_parity as P
from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U8, po.F64, po.C32, po.B8])
def test_reduce_dimwise_big_vs_reference_kernels(af, ref_oracle, code):
    P.test_reduce_dimwise_big(af, ref_oracle, code)


@pytest.mark.parametrize("code", [po.F32, po.S32, po.F64, po.C32, po.U8, po.F16])
def test_ireduce_vs_reference_kernels(af, ref_oracle, code):
    P.test_ireduce(af, ref_oracle, code)


@pytest.mark.parametrize("code", [po.F32, po.S32, po.F64, po.S64, po.C32, po.U8, po.B8])
def test_pipe_scan_vs_reference_kernels(af, ref_oracle, code):
    P.test_pipe_scan(af, ref_oracle, code)


@pytest.mark.parametrize("code", [po.F32, po.S32, po.F64])
def test_pipe_scan_by_key_vs_reference_kernels(af, ref_oracle, code):
    P.test_pipe_scan_by_key(af, ref_oracle, code)


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U8, po.C64, po.F16])
def test_reduce_all_vs_reference_kernels(af, ref_oracle, code):
    P.test_reduce_all(af, ref_oracle, code)
