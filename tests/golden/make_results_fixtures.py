"""Generates tests/golden/results_ref.json by running the REFERENCE's `generate_pred` / `generate_predV2`
(evaluation/generate_lane.py:31-61) -- bodies cut out of the file under /root/reference with `ast`, because the module
itself imports cv2 / yaml -- on stub lanes, in a scratch directory, and recording the text they write.

    python tests/golden/make_results_fixtures.py        (needs /root/reference; CPU only)
"""
import ast
import json
import os
import tempfile
import textwrap

import numpy as np

SRC = "/root/reference/evaluation/generate_lane.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "results_ref.json")


class StubLane:
    def __init__(self, points):
        self.points = points


def main():
    src = open(SRC).read()
    fns = {n.name: textwrap.dedent(ast.get_source_segment(src, n)) for n in ast.walk(ast.parse(src))
           if isinstance(n, ast.FunctionDef) and n.name in ("generate_pred", "generate_predV2")}
    ns = {"os": os, "np": np}
    for f in fns.values():
        exec(f, ns)
    rng = np.random.default_rng(7)
    cases = []
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for fmt, fn, sub, size in (("vil", "generate_pred", "evaluation/txt/pred_txt", (1080, 1920)),
                                       ("openlane", "generate_predV2", "evaluation/txt4OL/pred_txt", (1280, 1920))):
                for c in range(4):
                    K, n_off = 6, 36 if fmt == "vil" else 72
                    npts = rng.integers(0, n_off + 1, size=K)
                    npts[0], npts[1], npts[2] = 2, 3, 0
                    pts = np.zeros((K, n_off, 2))
                    for k in range(K):
                        pts[k, :npts[k], 0] = rng.uniform(-0.2, 1.3, npts[k])
                        pts[k, :npts[k], 1] = np.sort(rng.uniform(0, 1, npts[k]))
                    lanes = [StubLane(pts[k, :npts[k]].copy()) for k in range(K) if npts[k] > 1]   # get_lanes drops <= 1 point
                    info = {"name": f"clip{c}", "ImgName": ["00000", f"{c:05d}"], "size": size}
                    ns[fn](info, lanes, 1)
                    text = open(os.path.join(sub, info["name"], info["ImgName"][1] + ".lines.txt")).read()
                    cases.append({"fmt": fmt, "size": list(size), "points": pts.tolist(),
                                  "npoints": [int(v) if v > 1 else 0 for v in npts], "text": text})
        finally:
            os.chdir(cwd)
    json.dump(cases, open(OUT, "w"))
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
