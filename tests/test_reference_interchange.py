"""Checkpoint interchange with the reference's own NeRFPipeline (the Instant-NGP pipeline of the
reference cannot be constructed here: tiny-cuda-nn is absent). The reference package is also called
`atmonr`, so it runs in a child process (tests/golden/make_golden.py: import_reference). Only
possible in the build container; skipped where /root/reference does not exist (the GPU box)."""

import json
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT, FakeDataset, tiny_scene

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

CHILD = r"""
import sys, json, torch
from types import SimpleNamespace
sys.path.insert(0, sys.argv[1])
from make_golden import import_reference
import_reference()
from atmonr.pipelines.nerf import NeRFPipeline
cfg = json.loads(sys.argv[2])
ds = SimpleNamespace(config={"ray_origin_height": 20000}, scale=1.0, offset=torch.zeros(3, dtype=torch.float64),
                     max_i=0.3, get_point_preprocessor=lambda name: (lambda p: p))
torch.manual_seed(3)
pipe = NeRFPipeline(cfg, ds)
if sys.argv[3] == "save":
    torch.save(pipe.state_dict(), sys.argv[4])
else:
    pipe.load_state_dict(torch.load(sys.argv[4]))      # strict: every key and shape must match
    sd = pipe.state_dict()
    torch.save({m: {k: v.clone() for k, v in sd[m].items()} for m in sd}, sys.argv[5])
"""


def _child(*args):
    r = subprocess.run([sys.executable, "-c", CHILD, os.path.join(ROOT, "tests", "golden"), *args],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_nerf_checkpoint_interchanges_with_the_reference(tmp_path):
    from atmonr.pipelines.factory import get_pipeline
    cfg = json.load(open(os.path.join(ROOT, "configs", "nerf.json")))["pipeline"]
    cfg["mlp_hidden_dim"] = 32
    ref_pt, mine_pt, back_pt = (str(tmp_path / n) for n in ("ref.pt", "mine.pt", "back.pt"))
    # reference -> this package
    _child(json.dumps(cfg), "save", ref_pt)
    ref_sd = torch.load(ref_pt)
    pipe = get_pipeline(cfg, FakeDataset(tiny_scene(h=2, w=2, n_views=3)))
    pipe.load_state_dict(ref_sd)
    mine = pipe.state_dict()
    assert set(mine) == set(ref_sd) == {"coarse", "fine"}
    for mode in mine:
        assert list(mine[mode]) == list(ref_sd[mode])                  # same names, same order (fc1..fc11)
        for k in mine[mode]:
            assert torch.equal(mine[mode][k], ref_sd[mode][k]), (mode, k)
    # this package -> reference
    torch.manual_seed(8)
    pipe2 = get_pipeline(cfg, FakeDataset(tiny_scene(h=2, w=2, n_views=3)))
    torch.save(pipe2.state_dict(), mine_pt)
    _child(json.dumps(cfg), "load", mine_pt, back_pt)
    back = torch.load(back_pt)
    for mode in back:
        for k, v in pipe2.state_dict()[mode].items():
            assert torch.equal(back[mode][k], v), (mode, k)
