"""Shared helpers for the parity tests: run the product op (C-ABI, CUDA) and the oracle on the same inputs."""
import numpy as np
import torch

from oracle import oracle


def oracle_batched(props, scores, thr, top_k, n_valid=None, sort_model=0, lazy=True):
    p = props.detach().cpu().numpy()
    s = scores.detach().cpu().numpy()
    nv = None if n_valid is None else n_valid.detach().cpu().numpy()
    return oracle.nms_batched(p, s, nv, float(thr), int(top_k), sort_model=sort_model, lazy=lazy)


def assert_same(got, want, ctx=""):
    keep, num, parent = [t.detach().cpu().numpy() for t in got]
    wk, wn, wp = want
    num = np.asarray(num).reshape(-1)
    wn = np.asarray(wn).reshape(-1)
    keep = keep.reshape(len(wn), -1)
    parent = parent.reshape(len(wn), -1)
    wk = np.asarray(wk).reshape(len(wn), -1)
    wp = np.asarray(wp).reshape(len(wn), -1)
    bad = np.nonzero(num != wn)[0]
    assert bad.size == 0, f"{ctx}: num_to_keep differs in frames {bad[:8]}: got {num[bad[:8]]} want {wn[bad[:8]]}"
    bad = np.nonzero((keep != wk).any(axis=1))[0]
    assert bad.size == 0, f"{ctx}: keep differs in frames {bad[:8]}; first: got {keep[bad[0]][:12]} want {wk[bad[0]][:12]}"
    bad = np.nonzero((parent != wp).any(axis=1))[0]
    assert bad.size == 0, f"{ctx}: parent differs in frames {bad[:8]} ({(parent[bad[0]] != wp[bad[0]]).sum()} entries in the first)"
