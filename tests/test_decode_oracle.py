"""CPU: the decode restatement (oracle/decode_oracle.py) against outputs of the reference's own `predictions_to_pred`
bodies (tests/golden/decode_ref.npz, made by tests/golden/make_decode_fixtures.py from /root/reference)."""
import os

import numpy as np

from oracle import decode_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode_ref.npz")


def cases():
    z = np.load(GOLD)
    for key in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        hdr = int(key[1])
        yield key, hdr, z[key + "_in"], z[key + "_pts"], z[key + "_cnt"], z[key + "_meta"]


def test_oracle_matches_reference_fixtures():
    n = 0
    for key, hdr, inp, pts, cnt, meta in cases():
        got = decode_oracle.predictions_to_pred(inp, hdr, 720, 120 if hdr == 7 else 0)
        assert len(got) == len(inp)
        for i, g in enumerate(got):
            if cnt[i] == 0:
                assert g is None, f"{key} lane {i}: the reference skips this lane"
                continue
            assert g is not None and g[0].shape == (cnt[i], 2), f"{key} lane {i}"
            assert np.array_equal(g[0], pts[i, :cnt[i]], equal_nan=True), f"{key} lane {i}: points differ"
            assert np.array_equal(np.float32(g[1]), meta[i]), f"{key} lane {i}: metadata differs"
            n += 1
    assert n > 300
