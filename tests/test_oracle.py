"""CPU tests of the oracle itself: against the committed golden vectors (outputs of the reference's own CUDA op and
of torch's CUDA sort, captured on a B200), against an independent pure-Python restatement, and on hand-made cases."""
import os

import numpy as np
import pytest

from oracle import oracle
from phnet_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases(z):
    return sorted({k.split("/")[0] for k in z.files})


def test_oracle_matches_reference_cuda_op_golden():
    z = np.load(os.path.join(GOLD, "ref_nms_b200.npz"))
    checked = 0
    for nm in _cases(z):
        p, s = z[nm + "/props"], z[nm + "/scores"]
        assert (oracle.order(s) == z[nm + "/order"]).all(), f"{nm}: order differs from torch CUDA sort"
        for key in [k for k in z.files if k.startswith(nm + "/thr") and k.endswith("/keep")]:
            tag = key[: -len("/keep")]
            thr = float(tag.split("/thr")[1].split("_k")[0])
            top_k = int(tag.split("_k")[1])
            for lazy in (False, True):
                keep, num, parent = oracle.nms(p, s, thr, top_k, lazy=lazy)
                assert num == int(z[tag + "/num"]), (tag, lazy)
                assert (keep == z[tag + "/keep"]).all(), (tag, lazy)
                assert (parent == z[tag + "/parent"]).all(), (tag, lazy)
                checked += 1
    assert checked >= 300


def test_sort_model_matches_torch_cuda_sort_golden():
    z = np.load(os.path.join(GOLD, "torch_cuda_sort.npz"))
    names = _cases(z)
    assert len(names) >= 200
    for nm in names:
        s = z[nm + "/scores_bits"].view(np.float32)
        assert (oracle.order(s, oracle.SORT_TORCH_CUDA) == z[nm + "/order"]).all(), nm
        assert (oracle.order(s, oracle.SORT_STABLE_RADIX) == z[nm + "/order_stable"]).all(), nm


@pytest.mark.parametrize("n_off", [36, 72])
def test_c_oracle_matches_pure_python_restatement(n_off):
    for seed in range(3):
        p, s = synth.edge_frame(n_off, seed=seed)
        p, s = p.numpy()[:48], s.numpy()[:48]
        for top_k in (0, 1, 4, 48):
            a = oracle.nms(p, s, 50.0, top_k)
            b = oracle.nms_py(p, s, 50.0, top_k)
            assert a[1] == b[1] and (a[0] == b[0]).all() and (a[2] == b[2]).all(), (seed, top_k)
    props, scores = synth.make_frames(1, 70, n_off, seed=4)
    a = oracle.nms(props[0].numpy(), scores[0].numpy(), 30.0, 4)
    b = oracle.nms_py(props[0].numpy(), scores[0].numpy(), 30.0, 4)
    assert a[1] == b[1] and (a[0] == b[0]).all() and (a[2] == b[2]).all()


def test_literal_equals_lazy_on_big_frames():
    props, scores = synth.make_frames(2, 1000, 72, seed=3)
    for f in range(2):
        for top_k in (0, 4, 1000):
            a = oracle.nms(props[f].numpy(), scores[f].numpy(), 50.0, top_k, lazy=False)
            b = oracle.nms(props[f].numpy(), scores[f].numpy(), 50.0, top_k, lazy=True)
            assert a[1] == b[1] and (a[0] == b[0]).all() and (a[2] == b[2]).all()


def _lane(n_off, start_y, length, x):
    row = np.zeros(5 + n_off, dtype=np.float32)
    row[2], row[4] = start_y, length
    row[5:] = x
    return row


def test_handmade_predicate_cases():
    n = 72
    a = _lane(n, 0.0, 72, 100.0)
    assert oracle.pred(a, a, n, 50.0)                                   # identical lanes: distance 0 < 50 * 72
    assert not oracle.pred(a, a, n, 0.0)                                # 0 < 0 is false
    assert oracle.pred(a, _lane(n, 0.0, 72, 149.9), n, 50.0)            # mean |dx| just under the threshold
    assert not oracle.pred(a, _lane(n, 0.0, 72, 150.0), n, 50.0)        # exactly at the threshold: strict <
    lo, hi = _lane(n, 0.0, 10, 100.0), _lane(n, 0.5, 10, 100.0)         # rows 0..9 vs rows 36..45: disjoint y-ranges
    assert not oracle.pred(lo, hi, n, 50.0)
    assert oracle.lane_bounds(_lane(n, 0.0, 0.0, 0), n) == (0, -1)      # length 0: -1 + 0.5 - 1 = -1.5 -> -1 (the `- (x<0)` trick)
    assert oracle.lane_bounds(_lane(n, 0.0, 0.4, 0), n) == (0, -1)
    assert oracle.lane_bounds(_lane(n, 0.0, 1.0, 0), n) == (0, 0)
    assert oracle.lane_bounds(_lane(n, 1.0, 200.0, 0), n) == (71, 270)  # end is clamped per pair, not per lane
    assert oracle.lane_bounds(_lane(n, float("nan"), 10.0, 0), n)[0] == -2147483648   # B200 F2I.F64(NaN) = INT_MIN
    assert oracle.lane_bounds(_lane(n, 1e12, 10.0, 0), n)[0] == 2147483647            # saturation
    # negative start in [-5,-1]: the header columns enter the sum (unsigned char counter starts at 5+start >= 0)
    neg_a, neg_b = _lane(n, -0.03, 20, 100.0), _lane(n, -0.03, 20, 100.0)
    assert oracle.lane_bounds(neg_a, n)[0] == -1
    neg_b[4] = 20.0
    neg_b[3] = 1e6                                                      # start_x differs hugely but is column 3, not summed
    assert oracle.pred(neg_a, neg_b, n, 50.0)
    neg_b[4] = 1e6                                                      # column 4 (length) IS summed when start == -1 ...
    neg_a2 = neg_a.copy()
    neg_a2[4] = 1e6                                                     # ... unless both lanes carry the same value
    assert oracle.pred(neg_a2, neg_b, n, 50.0)
    # start <= -6: the counter wraps past the row, the loop is skipped, dist = 0 < thr * len  ->  suppressed
    w_a, w_b = _lane(n, -0.1, 30, 0.0), _lane(n, -0.1, 30, 700.0)
    assert oracle.lane_bounds(w_a, n)[0] == -6
    assert oracle.pred(w_a, w_b, n, 50.0)
    nan_x = _lane(n, 0.0, 72, 100.0)
    nan_x[40] = float("nan")
    assert not oracle.pred(a, nan_x, n, 50.0)                           # NaN distance compares false
    assert not oracle.pred(a, a, n, float("nan"))


def test_handmade_collect_cases():
    n_off, N = 36, 10
    props = np.zeros((N, 5 + n_off), dtype=np.float32)
    props[:, 4] = n_off
    props[:, 5:] = (np.arange(N, dtype=np.float32) * 10.0)[:, None]     # lanes 10 px apart
    scores = np.linspace(0.9, 0.1, N).astype(np.float32)
    keep, num, parent = oracle.nms(props, scores, 25.0, 4)              # lane i suppresses i+1, i+2
    assert num == 4 and list(keep[:4]) == [0, 3, 6, 9] and (keep[4:] == 0).all()
    assert list(parent) == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4]
    keep, num, parent = oracle.nms(props, scores, 25.0, 2)              # top_k stops the scan: the rest stays untouched
    assert num == 2 and list(keep[:2]) == [0, 3] and list(parent) == [1, 1, 1, 2, 2, 2, 0, 0, 0, 0]
    keep, num, parent = oracle.nms(props, scores, 25.0, 0)              # top_k == 0 never stops and reports 0
    assert num == 0 and list(keep[:4]) == [0, 3, 6, 9]
    keep, num, parent = oracle.nms(props, scores, 25.0, 100)
    assert num == 4
    # last writer wins: lane 2 is covered by kept lane 0 (slot 1) and, with a wider threshold, again by a later one
    keep, num, parent = oracle.nms(props, scores[::-1].copy(), 25.0, 4)  # reversed scores: order 9, 8, ...
    assert list(keep[:4]) == [9, 6, 3, 0]
    # ties: stable order above 32 elements, ATen's bitonic network at or below 32
    tied = np.full(40, 0.5, dtype=np.float32)
    assert list(oracle.order(tied)) == list(range(40))
    z = np.load(os.path.join(GOLD, "torch_cuda_sort.npz"))
    s = z["N20_v1/scores_bits"].view(np.float32)
    assert (oracle.order(s) == z["N20_v1/order"]).all()
    assert (oracle.order(s) != oracle.order(s, oracle.SORT_STABLE_RADIX)).any()   # the network really is unstable


def test_batched_and_n_valid():
    props, scores = synth.make_frames(6, 120, 36, seed=1)
    nv = np.array([0, 1, 33, 120, 64, 7], dtype=np.int32)
    keep, num, parent = oracle.nms_batched(props.numpy(), scores.numpy(), nv, 50.0, 4, threads=3)
    for f in range(6):
        n = int(nv[f])
        k, m, p = oracle.nms(props[f, :n].numpy(), scores[f, :n].numpy(), 50.0, 4) if n else (np.zeros(0), 0, np.zeros(0))
        assert num[f] == m and (keep[f, :n] == k).all() and (parent[f, :n] == p).all()
        assert (keep[f, n:] == 0).all() and (parent[f, n:] == 0).all()


def test_argument_errors():
    with pytest.raises(RuntimeError):
        oracle.nms(np.zeros((4, 5 + 251), np.float32), np.zeros(4, np.float32), 50.0, 4)
    with pytest.raises(RuntimeError):
        oracle.nms(np.zeros((64000, 6), np.float32), np.zeros(64000, np.float32), 50.0, 4)
