"""CPU: the `*.lines.txt` writer (phnet_b200/results.py) against text produced by the reference's own generate_pred /
generate_predV2 (tests/golden/results_ref.json, made by tests/golden/make_results_fixtures.py from /root/reference)."""
import json
import os

import numpy as np

from phnet_b200 import results

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "results_ref.json")


def test_frame_text_matches_reference_writer(tmp_path):
    cases = json.load(open(GOLD))
    assert len(cases) == 8
    for c in cases:
        pts, n = np.array(c["points"], dtype=np.float64), np.array(c["npoints"])
        assert results.frame_lines(pts, n, c["size"], c["fmt"]) == c["text"]
    c = cases[0]
    pts = np.array([c["points"], c["points"]], dtype=np.float64)
    n = np.array([c["npoints"], c["npoints"]])
    paths = results.write_clip(pts, n, str(tmp_path), "clipA", ["00001", "00002"], c["size"], c["fmt"])
    assert [os.path.basename(p) for p in paths] == ["00001.lines.txt", "00002.lines.txt"]
    assert open(paths[1]).read() == c["text"] and os.path.dirname(paths[0]).endswith("clipA")
