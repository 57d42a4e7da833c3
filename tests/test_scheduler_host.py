"""Euler-discrete schedule tables (genima_b200/scheduler.py, host scalars only) against the oracle restatement of
diffusers' EulerDiscreteScheduler and against closed-form facts."""
import math

import numpy as np
import pytest
import torch

from genima_b200.configs import SchedulerConfig
from genima_b200.scheduler import EulerDiscreteSchedule
from oracle.scheduler import EulerDiscreteOracle


@pytest.mark.parametrize("n,expect", [(1, [999]), (4, [999, 749, 499, 249]), (5, [999, 799, 599, 399, 199]),
                                       (10, [999, 899, 799, 699, 599, 499, 399, 299, 199, 99])])
def test_trailing_timesteps(n, expect):
    ts, sig = EulerDiscreteSchedule().set_timesteps(n)
    assert ts.tolist() == expect
    ots, osig = EulerDiscreteOracle().set_timesteps(n)
    assert np.array_equal(ts, ots) and np.array_equal(sig, osig)
    assert sig[-1] == 0.0 and np.all(np.diff(sig) < 0)


def test_sigma_max_and_init_noise_sigma():
    s = EulerDiscreteSchedule()
    s.set_timesteps(5)
    assert abs(s.init_noise_sigma - 14.6146) < 1e-3          # sigma_999 of the SD scaled-linear schedule
    lead = EulerDiscreteSchedule(SchedulerConfig(timestep_spacing="leading"))
    lead.set_timesteps(5)
    assert abs(lead.init_noise_sigma - math.sqrt(float(lead.sigmas.max()) ** 2 + 1)) < 1e-6


def test_unknown_scheduler_classes_fail_loudly():
    with pytest.raises(NotImplementedError):
        EulerDiscreteSchedule(SchedulerConfig(class_name="DPMSolverMultistepScheduler"))
    with pytest.raises(NotImplementedError):
        EulerDiscreteSchedule(SchedulerConfig(prediction_type="v_prediction"))


def test_euler_ancestral_tables_and_step():
    """EulerAncestralDiscreteScheduler (sdxl-turbo): sigma_up^2 + sigma_down^2 = sigma_to^2, no noise on the last step,
    and with zero noise the step is the plain Euler move to sigma_down."""
    from oracle.scheduler import EulerAncestralOracle

    sch = EulerDiscreteSchedule(SchedulerConfig(class_name="EulerAncestralDiscreteScheduler"))
    assert sch.ancestral
    ts, sig = sch.set_timesteps(4)
    o = EulerAncestralOracle()
    ts_o, sig_o = o.set_timesteps(4)
    assert np.array_equal(ts, ts_o) and np.array_equal(sig, sig_o) and sch.init_noise_sigma == o.init_noise_sigma
    g = torch.Generator().manual_seed(0)
    x, eps, noise = (torch.randn(1, 4, 8, 8, generator=g) for _ in range(3))
    for i in range(4):
        up, down = sch.ancestral_sigmas(i)
        assert abs(up * up + down * down - float(sig[i + 1]) ** 2) < 1e-5 * max(1.0, float(sig[i + 1]) ** 2)
        want = x + eps * (down - float(sig[i])) + noise * up
        assert torch.allclose(o.step(eps, i, x, noise), want, rtol=1e-5, atol=1e-5)
    assert sch.ancestral_sigmas(3) == (0.0, 0.0)


def test_euler_step_equals_ddim_eta0_update():
    """with y = x / sqrt(sigma^2 + 1), DDIM(eta=0) on y is the Euler step on x."""
    o = EulerDiscreteOracle()
    _, sig = o.set_timesteps(5)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64) * float(sig[0])
    eps = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    for i in range(5):
        s, s2 = float(sig[i]), float(sig[i + 1])
        x_next = o.step(eps.float(), i, x.float()).double()
        ab, ab2 = 1 / (s * s + 1), 1 / (s2 * s2 + 1)                      # alpha_bar from sigma
        y = x * math.sqrt(ab)
        y2 = math.sqrt(ab2) * (y - math.sqrt(1 - ab) * eps) / math.sqrt(ab) + math.sqrt(1 - ab2) * eps
        assert torch.allclose(x_next * math.sqrt(ab2), y2, rtol=1e-5, atol=1e-5)
        x = x_next


@pytest.mark.parametrize("spacing,offset,alpha_one", [("leading", 1, False), ("trailing", 0, True), ("linspace", 0, False)])
def test_ddim_tables_and_coefficients(spacing, offset, alpha_one):
    """DDIMScheduler (eta = 0) as x' = a x + b eps: timesteps and per-step coefficients against the oracle restatement of
    diffusers' step(); sigmas are all zero so that the Euler-form plumbing scales nothing; clip_sample raises."""
    from oracle.scheduler import DDIMOracle

    cfg = SchedulerConfig(class_name="DDIMScheduler", timestep_spacing=spacing, steps_offset=offset,
                          set_alpha_to_one=alpha_one)
    sch = EulerDiscreteSchedule(cfg)
    assert sch.ddim and not sch.ancestral
    ts, sig = sch.set_timesteps(5)
    o = DDIMOracle(timestep_spacing=spacing, steps_offset=offset, set_alpha_to_one=alpha_one)
    assert np.array_equal(ts, o.set_timesteps(5)[0])
    assert sch.init_noise_sigma == 1.0 and not sig.any() and len(sig) == 6
    g = torch.Generator().manual_seed(0)
    x, eps = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    for i in range(5):
        a, b = sch.ddim_coeffs(i)
        assert torch.allclose(a * x + b * eps, o.step(eps, i, x), rtol=1e-5, atol=1e-5)
        x = o.step(eps, i, x)
    with pytest.raises(NotImplementedError):
        EulerDiscreteSchedule(SchedulerConfig(class_name="DDIMScheduler", clip_sample=True))


def test_ddim_trailing_equals_euler_in_scaled_variable():
    """SURVEY.md Appendix D, now through both schedule objects: with set_alpha_to_one the DDIM(eta = 0) trajectory of
    y = x / sqrt(sigma^2 + 1) is the Euler trajectory of x on the same trailing timesteps."""
    eu = EulerDiscreteSchedule(SchedulerConfig())
    dd = EulerDiscreteSchedule(SchedulerConfig(class_name="DDIMScheduler"))
    ts_e, sig = eu.set_timesteps(5)
    ts_d, _ = dd.set_timesteps(5)
    assert np.array_equal(ts_e, ts_d)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64) * float(sig[0])
    y = x / math.sqrt(float(sig[0]) ** 2 + 1)
    for i in range(5):
        eps = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
        x = x + (float(sig[i + 1]) - float(sig[i])) * eps
        a, b = dd.ddim_coeffs(i)
        y = a * y + b * eps
        assert torch.allclose(y, x / math.sqrt(float(sig[i + 1]) ** 2 + 1), rtol=1e-4, atol=1e-4)


def test_euler_leading_adds_steps_offset():
    """EulerDiscreteScheduler 'leading' spacing: (arange(n) * (T // n))[::-1] + steps_offset (sdxl-base / SD-1.x ship
    leading spacing with steps_offset 1): n = 4 -> [751, 501, 251, 1]."""
    from oracle.scheduler import EulerDiscreteOracle

    s = EulerDiscreteSchedule(SchedulerConfig(timestep_spacing="leading", steps_offset=1))
    ts, sig = s.set_timesteps(4)
    assert ts.tolist() == [751.0, 501.0, 251.0, 1.0]
    o = EulerDiscreteOracle(timestep_spacing="leading", steps_offset=1)
    ots, osig = o.set_timesteps(4)
    assert np.array_equal(ts, ots) and np.array_equal(sig, osig)
    assert EulerDiscreteSchedule(SchedulerConfig(timestep_spacing="leading")).set_timesteps(4)[0].tolist() == [750.0, 500.0, 250.0, 0.0]


def test_scheduler_json_defaults_and_unsupported_flags():
    from genima_b200.checkpoint import scheduler_config_from_json

    # diffusers' defaults when the key is missing: Euler / Euler-ancestral 'linspace', DDIM 'leading'
    assert scheduler_config_from_json({"_class_name": "EulerDiscreteScheduler"}).timestep_spacing == "linspace"
    assert scheduler_config_from_json({"_class_name": "EulerAncestralDiscreteScheduler"}).timestep_spacing == "linspace"
    assert scheduler_config_from_json({"_class_name": "DDIMScheduler"}).timestep_spacing == "leading"
    turbo = {"_class_name": "EulerDiscreteScheduler", "timestep_spacing": "trailing", "steps_offset": 1,
             "use_karras_sigmas": False, "interpolation_type": "linear", "rescale_betas_zero_snr": False,
             "timestep_type": "discrete", "final_sigmas_type": "zero", "sigma_min": None, "sigma_max": None}
    cfg = scheduler_config_from_json(turbo)
    assert cfg.timestep_spacing == "trailing" and cfg.steps_offset == 1
    for k, v in (("use_karras_sigmas", True), ("interpolation_type", "log_linear"), ("rescale_betas_zero_snr", True),
                 ("timestep_type", "continuous"), ("final_sigmas_type", "sigma_min"), ("thresholding", True)):
        with pytest.raises(NotImplementedError, match=k):
            scheduler_config_from_json(dict(turbo, **{k: v}))
