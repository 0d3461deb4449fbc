"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` : oracle vs the reference's golden vectors / oracle/_ref, host logic, C-ABI symbol check.
`-m gpu`       : parity tests proper -- the CUDA path called through the C-ABI, checked by the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.AfOracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import pyoracle
    pyoracle.build()
    if not pyoracle.Oracle.available("reference"):
        pytest.skip("oracle/_ref/libafref.so not built (no /root/reference and no prebuilt copy)")
    return pyoracle.AfOracle("reference")
