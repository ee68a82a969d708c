"""Import the *real* reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/segtran_oracle.py header).  The GPU box has no
/root/reference; callers must check ``available()`` first.  Recipe probed in SURVEY.md App. B:
two ``sys.modules`` stubs (``train_util``, ``timm.models``) for names the shells import but
never call, and a redirect of the hard-coded ``device='cuda'`` literal (segtran3d.py:464) when
no GPU is present.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from argparse import Namespace

import torch

REF_ROOT = os.environ.get("SEGTRAN_REFERENCE", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "networks", "segtran_shared.py"))


_loaded = {}


def load():
    """Returns a namespace with the reference modules: shared, seg3d, seg2d."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    if "train_util" not in sys.modules:
        m = types.ModuleType("train_util")
        m.batch_norm = None                      # imported at segtran3d.py:15, never called
        sys.modules["train_util"] = m
    if "timm" not in sys.modules:
        t = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        for n in ("tf_efficientnetv2_s_in21k", "tf_efficientnetv2_m_in21k", "tf_efficientnetv2_l_in21k"):
            setattr(tm, n, None)                 # segtran2d.py:13
        t.models = tm
        sys.modules["timm"] = t
        sys.modules["timm.models"] = tm
    with contextlib.redirect_stdout(io.StringIO()):
        import networks.segtran_shared as shared
        import networks.segtran3d as seg3d
        import networks.segtran2d as seg2d
    ns = Namespace(shared=shared, seg3d=seg3d, seg2d=seg2d)
    _loaded["ns"] = ns
    return ns


@contextlib.contextmanager
def cuda_literal_to_cpu():
    """segtran3d.py:464 builds ``torch.tensor(..., device='cuda')``; redirect on a CPU-only host."""
    if torch.cuda.is_available():
        yield
        return
    orig = torch.tensor

    def patched(*a, **kw):
        if kw.get("device") == "cuda":
            kw["device"] = "cpu"
        return orig(*a, **kw)

    torch.tensor = patched
    try:
        yield
    finally:
        torch.tensor = orig


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def encoder_config(shared, *, dims, num_modes=4, num_attractors=16, pos_dim=3, qk_have_bias=True,
                   dropout=0.0, trans_output_type="private"):
    """A SegtranConfig sufficient to build SegtranFusionEncoder stand-alone (what
    Segtran3dConfig.update_config + set_fpn_layers would derive, segtran_shared.py:158-196)."""
    cfg = shared.SegtranConfig()
    cfg.num_translayers = len(dims) - 1
    cfg.translayer_dims = list(dims)
    cfg.translayer_compress_ratios = [1] * len(dims)
    cfg.trans_in_dim = dims[0]
    cfg.trans_out_dim = dims[-1]
    cfg.min_feat_dim = min(dims)
    cfg.num_modes = num_modes
    cfg.num_attractors = num_attractors
    cfg.pos_dim = pos_dim
    cfg.qk_have_bias = qk_have_bias
    cfg.hidden_dropout_prob = dropout
    cfg.attention_probs_dropout_prob = dropout
    cfg.trans_output_type = trans_output_type
    cfg.mid_type = "shared"
    cfg.pos_code_type = "lsinu"
    cfg.use_squeezed_transformer = True
    cfg.tie_qk_scheme = "shared"
    return cfg


def build_encoder(cfg, seed=0):
    """Reference SegtranFusionEncoder with the reference's init sequence
    (init_weights -> tie_qk -> add_identity_bias; segtran3d.py:246-249)."""
    ns = load()
    shared = ns.shared
    torch.manual_seed(seed)
    with quiet():
        enc = shared.SegtranFusionEncoder(cfg, "Fusion")
        init = shared.SegtranInitWeights(cfg)
        enc.apply(init.init_weights)
        enc.apply(init.tie_qk)
        enc.apply(init.add_identity_bias)
    return enc
