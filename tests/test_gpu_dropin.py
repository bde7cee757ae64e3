"""The C++ drop-in (graph-embed_b200/host/include/embed.hpp) compiled against the reference's own
interface and run on the GPU, and the embedMultilevel out-parameters of ge_embed."""
import os
import subprocess

import numpy as np
import pytest

from helpers import load_hier_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin_demo_runs(capi):
    exe = os.path.join(ROOT, "graph-embed_b200", "lib", "ge_dropin_demo")
    if not os.path.exists(exe):
        from graph_embed_b200 import build
        build.build_all()
    for args in (["48", "2"], ["32", "3"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "embedded! in time" in r.stdout and "embedVia ok" in r.stdout
        assert "getting base coords" in r.stdout  # the reference's progress lines (src/embed.cpp:583)


def test_cpp_dropin_demo_writes_the_reference_files(capi, tmp_path):
    """--out: writeCoords (include/export.hpp:23) and the plot inputs of examples/embedder.cpp:230-289."""
    exe = os.path.join(ROOT, "graph-embed_b200", "lib", "ge_dropin_demo")
    r = subprocess.run([exe, "20", "3", "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = (tmp_path / "coords.txt").read_text().splitlines()
    assert len(rows) == 400 and all(len(l.split()) == 3 for l in rows)
    xyz = np.array([[float(v) for v in l.split()] for l in (tmp_path / "coords.temp").read_text().splitlines()])
    assert xyz.shape == (400, 3) and np.isfinite(xyz).all()
    head = (tmp_path / "part.temp").read_text().splitlines()
    assert head[0].split()[0] == "400" and int(head[0].split()[1]) == len(head[1].split())
    assert len((tmp_path / "mat.temp").read_text().splitlines()) == 2 * 2 * 20 * 19


def test_embed_level1_outputs_match_oracle(ctx, capi, oracle):
    """r_A / coords_A out-parameters of embedMultilevel (src/embed.cpp:580-581): recomputed here by
    running the oracle's radii step on the coordinates of a short embed of the coarser levels."""
    As, Ps, _ = load_hier_golden()
    x, _, r_A, coords_A = ctx.embed(As, Ps, 2, seed=9, coarse_iterations=20, level_iterations=3,
                                    return_level1=True)
    sub, _, r_Ac, coords_Ac = ctx.embed(As[1:], Ps[1:], 2, seed=9, coarse_iterations=20,
                                        level_iterations=3, return_level1=True)
    cA, rA = oracle.radii(sub, 2, As[1], Ps[1], coords_Ac, r_Ac)
    assert np.abs(coords_A - cA).max() < 1e-9 and np.abs(r_A - rA).max() < 1e-9
