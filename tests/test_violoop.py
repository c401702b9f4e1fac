"""Closed-loop sliding-window VIO (lfvio.violoop: processIMU -> addFeature -> triangulate -> optimization() -> MARGIN_OLD
prior -> slideWindow, estimator.cpp:122-219,1011-1075) with the CPU oracle's numeric steps, and — on a B200 — with the
device library on the same synthetic world: trajectories must coincide, so the ATE of the two arms is the same number.

The world is synthetic (hand-held-like trajectory, 150 tracked bearings per frame at 0.3 px / f = 160, 200 Hz IMU with
the PAL noise values); the absolute ATE describes that scenario, the parity statement is the difference between arms."""
import numpy as np
import pytest

from lfvio import violoop
from oracle import backend as ob
from oracle import loop_ops


def _opts(estimate_td):
    o = ob.default_opts()
    o.row = 960
    o.estimate_td = estimate_td
    return o


def test_closed_loop_oracle_tracks_the_trajectory():
    world = violoop.World(seed=2, n_frames=36)
    loop = violoop.VioLoop(world, loop_ops.oracle_ops(), _opts(0))
    loop.td = world.td_true
    loop.run()
    ate, ks = violoop.trajectory_error(loop)
    assert ks == list(range(10, 36)) and all(i == 8 or i > 0 for i in loop.iterations)
    assert ate < 0.12, ate                                         # ~1 % of the 3.3 m path in this scenario
    assert np.abs(loop.sb[-2, 6:9] - world.bg).max() < 0.01        # gyroscope bias is recovered
    assert loop.prior is not None and loop.prior.n == 6 * 10 + 9 + 6  # (no td block: estimate_td = 0)
    # FeatureManager invariants after sliding: tracks start inside the window and stay contiguous
    assert all(0 <= f.start <= 10 and f.start + len(f.obs) <= 11 for f in loop.features)


def test_closed_loop_noise_free_estimates_td_and_biases():
    world = violoop.World(seed=1, n_frames=30, pixel_noise=0.0)
    loop = violoop.VioLoop(world, loop_ops.oracle_ops(), _opts(1), perturb=0.0)
    loop.run()
    assert abs(loop.td - world.td_true) < 3e-4 and np.abs(loop.sb[-2, 3:6] - world.ba).max() < 0.15
    assert violoop.trajectory_error(loop)[0] < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("estimate_td", [1, 0])
def test_closed_loop_device_equals_oracle(lfvio_mod, gpu, estimate_td):
    world = violoop.World(seed=3, n_frames=40)
    win = lfvio_mod.Window(10, 1000, 16384, device=gpu)
    dopts = lfvio_mod.solver_default_opts()
    dopts.row, dopts.estimate_td = 960, estimate_td
    arms = []
    for ops, opts in ((violoop.device_ops(lfvio_mod, win, gpu), dopts), (loop_ops.oracle_ops(), _opts(estimate_td))):
        loop = violoop.VioLoop(world, ops, opts)
        if not estimate_td:
            loop.td = world.td_true
        arms.append(loop.run())
    dev, ora = arms
    # (termination tests such as |cost change| <= 1e-6 cost can fire one iteration apart when they are borderline)
    assert len(dev.iterations) == len(ora.iterations) and max(abs(a - b) for a, b in zip(dev.iterations, ora.iterations)) <= 1
    ks = sorted(ora.estimates)
    diff = max(np.abs(dev.estimates[k][0] - ora.estimates[k][0]).max() for k in ks)
    ate_d, ate_o = violoop.trajectory_error(dev)[0], violoop.trajectory_error(ora)[0]
    print(f"closed loop estimate_td={estimate_td}: max |p_dev - p_oracle| = {diff:.3e} m, ATE device {ate_d:.5f} m, oracle {ate_o:.5f} m")
    assert diff < 1e-3, diff                      # 30 chained solve / marginalise / slide steps
    assert abs(ate_d - ate_o) < 1e-3
    win.close()
