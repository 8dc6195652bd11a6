"""The kinematic 2-D driver (tools/kinematic_2d.py: icicle's role, kin_cloud_2d_lgrngn.hpp:128-295) on the B200 back-end:
Eulerian donor-cell advection of th / rv + the Lagrangian microphysics, with the fields in host memory, in GPU memory (device
pointers through arrinfo_t) and with step_async overlapping the Eulerian step on its own thread."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def run(b200, steps=120, **kw):
    from kinematic_2d import Kinematic2D
    from tests import support as S
    with S.rng_mode(b200, 0, -1):
        m = Kinematic2D(b200, nx=38, nz=38, sd_conc=48, **kw)
    d0 = m.diagnostics()
    for _ in range(steps):
        m.step()
    return m, d0, m.diagnostics()


def test_eddy_forms_a_cloud_and_closes_the_water_budget(b200):
    m, d0, d1 = run(b200)
    assert d0["cloudy_cells"] == 0 and d0["RH_max"] > 1.05       # icicle's profile starts supersaturated aloft (kin_cloud_2d_lgrngn.hpp:127: "deals with initial supersaturation")
    assert d1["cloudy_cells"] > 30 and d1["rc_max"] > 1e-4, d1            # the updraft branch condenses > 0.1 g/kg
    assert d1["RH_max"] < 1.03, d1                                          # supersaturation stays bounded: condensation keeps up
    assert d1["sd_min"] > 0 and abs(d1["sd_mean"] - d0["sd_mean"]) < 0.05 * d0["sd_mean"], d1     # the non-divergent flow keeps cells populated
    budget = d1["total_water"] + 1e3 * d1["puddle_liquid_volume"] * 0.0     # nothing rains out in 2 minutes
    assert abs(budget - d0["total_water"]) < 2e-5 * d0["total_water"], (d0["total_water"], d1["total_water"])
    assert d1["liquid"] > 10 * d0["liquid"]


def test_device_resident_fields_and_async_step_give_the_same_model(b200):
    """fields kept on the GPU (device-pointer path) and step_async on its own thread: same trajectory as the plain host run up to
    the rounding of the two donor-cell implementations (numpy / torch)"""
    _, _, host = run(b200, steps=60)
    _, _, dev = run(b200, steps=60, device_fields=True, async_step=True)
    for k in ("total_water", "vapour", "th_max", "th_min"):
        assert abs(host[k] - dev[k]) <= 1e-9 * abs(host[k]), (k, host[k], dev[k])
    assert abs(host["liquid"] - dev["liquid"]) <= 1e-3 * host["liquid"]
    assert abs(host["cloudy_cells"] - dev["cloudy_cells"]) <= 3
