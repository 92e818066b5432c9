"""Generates tests/golden/two_magnets_8000_solution.npz: the CPU oracle's full solve of the bench
workload (BASELINE configs[1], 2 x 20x20x10 = 8000 cells, N = 24000), so that the GPU parity test can
compare the full-size solution without spending ~1.5 min of CPU per test run.

    python tests/golden/make_two_magnets_8000_solution.py          (needs ~30 GB of host memory)
"""
import hashlib
import json
import os
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from bench import make_workload  # noqa: E402
from oracle import demag_ref  # noqa: E402

w = make_workload("two_magnets_8000")
stages = {}
t0 = time.perf_counter()
x = demag_ref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], H_ext=w["H_ext"], split=32,
                              timings=stages, threads=os.cpu_count())
t = time.perf_counter() - t0
out = pathlib.Path(__file__).with_name("two_magnets_8000_solution.npz")
np.savez_compressed(out, x_local=x)
meta = {"workload": "two_magnets_8000", "desc": w["desc"], "seconds": round(t, 2), "stages_s": {k: round(v, 2) for k, v in stages.items()},
        "cores": os.cpu_count(), "sha256_x_local": hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(),
        "max_abs_x": float(np.abs(x).max())}
pathlib.Path(__file__).with_name("two_magnets_8000_solution.json").write_text(json.dumps(meta, indent=1) + "\n")
print(json.dumps(meta))
