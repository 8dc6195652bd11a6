"""Device-pointer fast path of arrinfo_t (SURVEY.md section 7 "Host-pointer API", section 8 f3): a host model whose Eulerian
fields already live in GPU memory passes device pointers; the library tells them apart by unified addressing and copies
device-to-device.  Results must be bit-identical to the same run fed from host arrays, th / rv are written back in place."""
import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def run(b200, device_arrays, scheme=L.as_t.implicit, steps=5):
    import torch
    oi, o, f = S.box_3d(b200, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True, adve=scheme)
    if device_arrays:
        f = {k: torch.from_numpy(v).cuda() for k, v in f.items()}
        torch.cuda.synchronize()
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    for _ in range(steps):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
    host = (lambda v: v.cpu().numpy()) if device_arrays else (lambda v: v.copy())
    return p.get_n(), p.get_attr("rw2"), p.get_attr("x"), p.get_attr("z"), host(f["th"]), host(f["rv"])


@pytest.mark.parametrize("scheme", [L.as_t.implicit, L.as_t.pred_corr])
def test_device_resident_fields_equal_host_fields(b200, scheme):
    a, b = run(b200, False, scheme), run(b200, True, scheme)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert not np.array_equal(a[4], S.box_3d(b200, nx=6, ny=5, nz=8)[2]["th"]), "condensation did not change th - vacuous"


def test_strided_device_arrays_are_refused(b200):
    import torch
    oi, o, f = S.box_3d(b200, nx=6, ny=5, nz=8, sd_conc=8)
    wide = torch.zeros((6, 5, 16), dtype=torch.float64, device="cuda")
    th = wide[:, :, ::2]                       # z stride 2: not contiguous along z
    th.copy_(torch.from_numpy(f["th"]))
    p = b200.factory(L.backend_t.CUDA, oi)
    with pytest.raises(RuntimeError):
        p.init(th, torch.from_numpy(f["rv"]).cuda(), torch.from_numpy(f["rhod"]).cuda(), None, f["Cx"], f["Cy"], f["Cz"])
