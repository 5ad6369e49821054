"""Accuracy of the straight-line FP64 math used inside the RHS (csrc/fw_math.cuh) against the host libm (numpy),
through the C ABI.  Stated bound: <= 2 ulp on the domain the dynamics reach."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ulp_err(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_hot_loop_math_within_2_ulp(cuda_device):
    import torch
    from tum_adlr_deep_reinforcement_learning_b200.batched import debug_math
    rs = np.random.RandomState(0)
    n = 1 << 20
    x = np.concatenate([rs.uniform(-160, 160, n // 2), rs.uniform(-2, 2, n // 2)])          # M * alpha, |alpha| <= pi
    got = debug_math(0, torch.as_tensor(x).cuda()).cpu().numpy()
    assert _ulp_err(got, np.exp(x)).max() <= 2.0
    assert np.array_equal(got, debug_math(4, torch.as_tensor(x).cuda()).cpu().numpy())      # immediate-coefficient build
    x = np.concatenate([rs.uniform(-0.5, 0.5, n // 2), rs.uniform(-1, 1, n // 2), [0.0, 0.5, -0.5, 1.0, -1.0, 1e-300]])
    got = debug_math(1, torch.as_tensor(x).cuda()).cpu().numpy()
    assert np.array_equal(got, debug_math(5, torch.as_tensor(x).cuda()).cpu().numpy())
    ref = np.arcsin(x)
    nz = ref != 0
    assert _ulp_err(got[nz], ref[nz]).max() <= 2.0 and np.all(got[~nz] == 0)
    xs = np.concatenate([rs.uniform(-30, 30, n), rs.standard_normal(n) * 1e-3, [0.0, 1.0, -1.0, 0.0, -2.0, 3.0]])
    ys = np.concatenate([rs.uniform(-30, 30, n), rs.uniform(-30, 30, n), [0.0, 0.0, 0.0, 2.0, 0.0, 3.0]])
    got = debug_math(2, torch.as_tensor(xs).cuda(), torch.as_tensor(ys).cuda()).cpu().numpy()
    assert np.array_equal(got, debug_math(6, torch.as_tensor(xs).cuda(), torch.as_tensor(ys).cuda()).cpu().numpy())
    ref = np.arctan2(ys, xs)
    nz = ref != 0
    assert _ulp_err(got[nz], ref[nz]).max() <= 2.0
    assert np.all(got[~nz] == 0)
    # 1/sqrt(x) on the range of |airspeed|^2 and of the asin argument reduction (normal range, x > 0)
    x = np.concatenate([np.exp(rs.uniform(np.log(1e-12), np.log(1e12), n)), [1.0, 4.0, 0.25, 400.0]])
    got = debug_math(3, torch.as_tensor(x).cuda()).cpu().numpy()
    assert _ulp_err(got, 1.0 / np.sqrt(x)).max() <= 2.0
    # log and x^p = exp(p log x) as the step-size controller uses them: err^-0.2 and (0.01 / d)^0.2
    x = np.concatenate([np.exp(rs.uniform(np.log(1e-300), np.log(1e300), n)), rs.uniform(0.5, 2.0, n), [1.0, 2.0, 0.5, np.sqrt(2.0)]])
    got = debug_math(7, torch.as_tensor(x).cuda()).cpu().numpy()
    ref = np.log(x)
    nz = ref != 0
    assert _ulp_err(got[nz], ref[nz]).max() <= 2.0 and np.all(got[~nz] == 0)
    # the relative error of exp(p log x) is the ABSOLUTE error of p log x: ~|p log x| ulp.  Error norms and initial-step
    # ratios live in [1e-8, 1e4] (<= 8 ulp there); the far range is bounded too
    for x, bound in ((np.concatenate([np.exp(rs.uniform(np.log(1e-8), np.log(1e4), n)), rs.uniform(1e-3, 10.0, n)]), 8.0),
                     (np.exp(rs.uniform(np.log(1e-30), np.log(1e30), n)), 64.0)):
        for p_ in (-0.2, 0.2):
            pw = torch.full((x.size,), p_, dtype=torch.float64).cuda()
            got = debug_math(8, torch.as_tensor(x).cuda(), pw).cpu().numpy()
            assert _ulp_err(got, np.power(x, p_)).max() <= bound, (p_, bound, _ulp_err(got, np.power(x, p_)).max())
            assert np.array_equal(got, debug_math(9, torch.as_tensor(x).cuda(), pw).cpu().numpy())
    edge = np.array([0.0, 1e-320, np.inf, np.nan])
    got = debug_math(8, torch.as_tensor(edge).cuda(), torch.full((4,), -0.2, dtype=torch.float64).cuda()).cpu().numpy()
    assert got[0] > 1e59 and got[1] > 1e59 and got[2] < 1e-59 and np.isnan(got[3])      # clamped like pow's limits, NaN kept
