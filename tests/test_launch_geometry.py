"""Launch geometry of the persistent stretch kernel (rcg_launch_geometry).

The kernels of W = 1024 .. 8192 keep their twiddles in 128 Tensor Memory columns per CTA and must still run three
CTAs per SM: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for any kernel that allocates Tensor Memory,
and a grid sized from that answer left two thirds of every SM idle (profiles/r3g_*, r3i_*, r3j_*)."""
import pytest

import rocoder_b200 as rb


@pytest.mark.gpu
@pytest.mark.parametrize("window,min_ctas", [(512, 3), (1024, 3), (2048, 3), (4096, 3), (8192, 3), (16384, 1), (32768, 1), (65536, 1)])
def test_stretch_kernel_occupancy(window, min_ctas):
    with rb.Engine(2, 48000, window, 4.0, 1.0, 1, seed=1) as e:
        g = e.launch_geometry()
    assert g["ctas_per_sm"] >= min_ctas, g
    sms = g["max_grid"] // g["ctas_per_sm"] * (2 if window == 65536 else 1)   # W = 65536: the grid counts CTA pairs
    assert g["max_grid"] % g["ctas_per_sm"] == 0 and sms >= 100, g   # SMs x CTAs per SM
    assert g["threads"] * g["ctas_per_sm"] <= 2048 and g["smem_bytes"] * g["ctas_per_sm"] <= 227 * 1024, g
    if window == 16384:
        assert (g["threads"], g["lanes_per_cta"]) == (768, 3), g


def test_launch_geometry_rejects_null():
    from rocoder_b200 import _abi
    assert _abi.lib().rcg_launch_geometry(None, None, None, None, None, None) == _abi.RCG_ERR_INVALID_ARG
