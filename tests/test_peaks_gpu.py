"""Same-run roofline denominators (rtpp_measure_peaks): plausibility on a B200."""
import pytest

import rtpp_b200

pytestmark = pytest.mark.gpu


def test_measured_peaks_are_plausible():
    with rtpp_b200.Ctx(0) as ctx:
        p = ctx.measure_peaks(8 << 20, 3)
    # B200: HBM3e copy 4-8 TB/s, L2 line stream above HBM, sector gather below the stream rate, FP32 below the nominal 74.4 TFLOP/s
    assert 3000 < p["hbm_copy_gbs"] < 8500, p
    assert p["l2_stream_gbs"] > p["hbm_copy_gbs"], p
    assert 500 < p["l2_sector_gather_gbs"] <= p["l2_stream_gbs"] * 1.05, p
    assert 30 < p["fp32_ffma_tflops"] < 80, p


def test_measure_peaks_rejects_bad_size():
    with rtpp_b200.Ctx(0) as ctx:
        with pytest.raises(rtpp_b200.RtppError):
            ctx.measure_peaks(3 << 20, 1)
