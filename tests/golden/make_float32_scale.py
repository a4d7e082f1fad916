"""Generates tests/golden/float32_scale.json: int64(float32(a) * pct) for pct in {1, 0.7},
computed with numpy.float32 (IEEE-754 binary32, round-to-nearest convert and multiply,
truncating conversion back) — the arithmetic of core.go:656-659.  The first nine rows are
SURVEY.md Appendix B."""
import json
import os

import numpy as np

VALS = [8000, 10000, 110, 16777217, 33554433, 17179869184, 17033068544, 67108864001,
        1099511640121, 137438953471, 549755813889, 96000, 128000, 1098437885953]
rows = [dict(alloc=str(a), pct_1=str(int(np.float32(a) * np.float32(1.0))),
             pct_07=str(int(np.float32(a) * np.float32(0.7)))) for a in VALS]
json.dump(rows, open(os.path.join(os.path.dirname(__file__), "float32_scale.json"), "w"), indent=1)
