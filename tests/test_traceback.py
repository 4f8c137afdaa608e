"""K5 (alignment path) parity against the reference's edlibAlign TASK_PATH vectors (make_golden_edlib_path.py)."""
import gzip
import json
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN


def _vectors():
    with gzip.open(os.path.join(GOLDEN, "edlib_path_vectors.json.gz"), "rt") as f:
        return json.load(f)


def _check(ctx, vec):
    dist, end, ops, flags = ctx.edlib_path_batch([c["q"] for c in vec], [c["t"] for c in vec], [c["mode"] for c in vec])
    n_h = 0
    for i, c in enumerate(vec):
        n_h += 1 if c["hirschberg"] else 0   # above edlib's 1 MiB switch: divide-and-conquer path
        assert flags[i] == 0, i
        assert int(dist[i]) == c["dist"], (i, c["mode"], len(c["q"]), len(c["t"]))
        if c["end"] is not None:
            assert int(end[i]) == c["end"], (i, "end")
        assert ops[i] == c["aln"], (i, c["mode"], len(c["q"]), len(c["t"]), "alignment path differs")
        # a path is a valid edit script: consumes the whole query and the aligned target prefix with `dist` edits
        if c["aln"]:
            assert sum(1 for o in ops[i] if o != 2) == len(c["q"])
            assert sum(1 for o in ops[i] if o != 0) == c["dist"]
    return n_h


def test_traceback_kernel_source_matches_reference(sim_lib):
    ctx = rb.Context(0, lib=sim_lib)
    vec = _vectors()
    vec = vec[:23] + vec[23::4] + [c for c in vec if c["hirschberg"]][:2]
    assert _check(ctx, vec) >= 2
    ctx.close()


@pytest.mark.gpu
def test_traceback_cuda_matches_reference_golden():
    ctx = rb.Context(0)
    assert _check(ctx, _vectors()) >= 1
    ctx.close()
