"""CPU: bench.py's helpers -- both arms describe the workload with ONE config dict (same_config),
the object-API collection of the end-to-end leg is the array workload, the reference arm pins its host threads."""

from __future__ import annotations

import json
import os
import pathlib
import subprocess
import sys

import numpy as np

import bench
from magpylib_material_response_b200.demag import _CellTable

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_collection_of_the_e2e_leg_is_the_array_workload():
    for name in ("two_magnets_1000", "soft_cube_512"):
        w = bench.make_workload(name)
        coll = bench.make_collection(name)
        tab = _CellTable(coll._source_groups())
        pos, quat, dim = tab.geometry()
        np.testing.assert_allclose(pos, w["pos"], rtol=0, atol=1e-18)
        np.testing.assert_allclose(np.abs(np.sum(quat * w["quat"], axis=1)), 1.0, atol=1e-15)
        np.testing.assert_allclose(dim, w["dim"], rtol=1e-15)
        np.testing.assert_array_equal(tab.pol_local(), w["pol"])
        np.testing.assert_array_equal(tab.susceptibilities(None), w["chi"])
        np.testing.assert_array_equal(tab.H_ext(), w["H_ext"])
        np.testing.assert_array_equal(bench.collection_polarizations(coll), w["pol"])


def test_reference_arm_line_and_same_config(tmp_path):
    env = dict(os.environ, OMP_NUM_THREADS="1", MMR_REF_BUDGET_S="20")  # torchrun's export must not throttle the arm
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "two_magnets_216"], capture_output=True, text=True, env=env, check=True, timeout=300)
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["steps"] == 1
    assert line["config"] == bench.bench_config("two_magnets_216", bench.make_workload("two_magnets_216"), "lu")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["full_size"]["n_cells"] == 216
    assert set(cb["full_size"]["stages_s"]) == {"assembly_s", "build_Q_s", "solve_s"}
    assert line["host_threads_env"]["OMP_NUM_THREADS"] == str(os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_sample_sizes_family():
    assert bench.sample_sizes("two_magnets_8000")[:3] == [8000, 5832, 4096]
    assert bench.sample_sizes("soft_cube_64000")[0] == 64000
