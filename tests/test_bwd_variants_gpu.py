"""The backward's alternative code paths, each in its own process (the switches are read once per process):
MOLKGNN_BWD_DBUF=0 -- single-buffered coefficient arrays (what a batch whose arrays do not fit twice gets),
MOLKGNN_BWD_PIPE=1 -- the software-pipelined kernel k_conv_bwd_pipe (opt-in),
MOLKGNN_BWD_MERGE=0 -- two block groups also where all blocks fit one pass (layer 0),
MOLKGNN_BWD_FUSE=0 -- the two block groups of a layer as two launches instead of two passes of one launch.
Each runs the golden / oracle parity tests of the stack and the tile-vs-SIMT comparison at the bench size."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"MOLKGNN_BWD_DBUF": "0"}, {"MOLKGNN_BWD_PIPE": "1"}, {"MOLKGNN_BWD_MERGE": "0"},
                                 {"MOLKGNN_BWD_FUSE": "0"}],
                         ids=["single_buffered", "pipelined", "two_passes_everywhere", "two_launches_per_layer"])
def test_backward_variant(env):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    e = dict(os.environ)
    e.update(env)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_conv_gpu.py", "tests/test_fullsize_gpu.py", "-k",
           "golden or vs_oracle or tile_and_simt or edge_case or bitwise"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
