import functools
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@functools.lru_cache(maxsize=None)
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def synth_ckpt(model="edge_n", nc=80, img=640, p2=False, p6=False, anchors=1, seed=7, obj_shift=0.0):
    from oracle import model_ref
    meta = model_ref.make_meta(model, nc, img, use_p2=p2, use_p6=p6, anchors=anchors)
    return model_ref.synth_checkpoint(meta, seed=seed, obj_bias_shift=obj_shift)


FWD_CASES = ["fwd_edge_n_64_nc3", "fwd_edge_n_320_nc80", "fwd_edge_n_96_p2p6_a2", "fwd_edge_m_64_nc3", "fwd_ms_n_64_nc4_p6"]


def case_ckpt(name):
    k = kat()[name]
    return synth_ckpt(k["model"], k["nc"], k["img"], k["p2"], k["p6"], k["anchors"], k["seed"]), k
