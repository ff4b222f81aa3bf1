"""CPU: the LUT form of the relative-position bias (one_peace_b200/relpos.py) reproduces the reference bucket matrices
exactly — text (bucket 256), audio (bucket 512) and image (w = 14 and 16) — and rejects a scheme it cannot represent."""
import numpy as np
import pytest
import torch

import restated as R
from one_peace_b200 import relpos


@pytest.mark.parametrize("bucket_size,S", [(256, 17), (256, 72), (256, 384), (512, 50), (512, 300)])
def test_token_buckets(bucket_size, S):
    bucket = R.make_token_bucket_position(bucket_size)[:S, :S].numpy()
    r = relpos.build_lut_index(bucket, relpos.text_codes(S))
    assert r is not None
    lut_idx, crow, ccol = r
    assert np.array_equal(lut_idx[crow[:S, None] - ccol[None, :S]], bucket)
    assert lut_idx.size % 4 == 0 and ccol.size % 4 == 0 and lut_idx.size * 4 + S * 5 <= 32768


@pytest.mark.parametrize("w", [14, 16])
def test_image_buckets(w):
    bucket = R.make_image_bucket_position(w).numpy()
    S = w * w + 1
    r = relpos.build_lut_index(bucket, relpos.image_codes(S, w))
    assert r is not None
    lut_idx, crow, ccol = r
    assert np.array_equal(lut_idx[crow[:S, None] - ccol[None, :S]], bucket)
    assert lut_idx.size * 4 + S * 5 <= 32768


def test_rejects_non_difference_scheme():
    rng = np.random.default_rng(0)
    bucket = rng.integers(0, 50, size=(20, 20))
    assert relpos.build_lut_index(bucket, relpos.text_codes(20)) is None
