"""GPU parity of the textOutline blur: the CUDA convolution
against the oracle's restatement of gst_ttml_blur_image_surface
(/root/reference/plugins/ttml/gstttmlblur.c:28-110). The Gaussian kernel is the reference's
own formula; the convolution semantics are pixman's (not installed: parity unpinned)."""
import numpy as np
import pytest

from helpers import pkg, random_overlay
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,radius,sigma", [
    (64, 48, 1, 0.5), (97, 33, 2, 1.0), (200, 60, 3, 1.5), (321, 75, 5, 2.5),
    (640, 90, 8, 4.0), (128, 128, 12, 6.0), (31, 9, 4, 2.0), (5, 5, 6, 3.0), (300, 40, 0, 1.0),
])
def test_blur_matches_oracle(ctx, w, h, radius, sigma):
    img = random_overlay(w, h, 100 + radius, premultiplied=True, density=0.5)
    want = oracle.blur_argb32(img, radius, sigma)
    got = ctx.blur_argb32(img, radius, sigma)
    assert np.array_equal(got, want), int((got != want).sum())


def test_blur_of_an_outline_like_surface(ctx):
    """What ttmlrender blurs: an outline stroke on a cleared surface padded by the radius
    (gstttmlrender.c:1189-1229)."""
    r = 6
    img = np.zeros((80 + 2 * r, 400 + 2 * r, 4), dtype=np.uint8)
    img[r + 20:r + 24, r + 10:r + 390] = (0, 0, 0, 255)
    img[r + 20:r + 60, r + 10:r + 14] = (40, 40, 40, 200)
    want = oracle.blur_argb32(img, r, r / 2.0)
    got = ctx.blur_argb32(img, r, r / 2.0)
    assert np.array_equal(got, want)
    assert got[:, :, 3].max() < 255 and got[r + 22, r + 200, 3] > 0     # it did blur


def test_blur_rejects_bad_arguments(ctx):
    tb = pkg.ttmlblend
    img = np.zeros((8, 8, 4), dtype=np.uint8)
    assert ctx.lib.fluc_ttmlblend_blur_argb32(ctx.h, img.ctypes.data, 8, 8, 32, 65, 1.0,
                                              img.ctypes.data, 32) == tb.ERROR_INVALID_ARGUMENT
    assert ctx.lib.fluc_ttmlblend_blur_argb32(ctx.h, img.ctypes.data, 8, 8, 32, 2, 0.0,
                                              img.ctypes.data, 32) == tb.ERROR_INVALID_ARGUMENT
