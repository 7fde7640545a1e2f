"""GPU tests (-m gpu) of the tcgen05 implicit-GEMM convolution kernels (forward, dgrad, wgrad, transposed conv)."""
import pytest

from tests.tc_cases import CASES, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_tc_conv(name):
    errs = run_case(name)
    for key, val in errs.items():
        assert val < 2e-2, (name, errs)
