"""GPU tier: randomised shape sweep (tools/fuzz_gpu.py) of the drop-in path against the CPU oracle — odd token and
patch counts, every D the tcgen05 kernel accepts, B = 1, ragged and non-prefix masks, Nv > 256."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 7])
def test_random_shapes_match_oracle(seed):
    from tools import fuzz_gpu
    worst, failures = fuzz_gpu.run_cases(30, seed, verbose=False)
    print({"fuzz_seed": seed, "idx_mismatch_total": worst["idx_mismatch_total"], "idx_rows_total": worst["idx_rows_total"]})
    assert not failures, failures
    assert worst["dq"] < 6e-3 and worst["dv"] < 6e-3
