"""Shared test helpers (tolerances, fixture loading, oracle drivers)."""
import os

import torch

from oracle import segtran_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def rel_err(a, b):
    """max|a-b| / max|b| — the '1e-3 rel' of BASELINE.json's north_star, as used in every parity test."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rms_rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def oracle_encoder(fx, x=None, dtype=torch.float32, collect=None):
    p = {"voxel_fusion." + k: v.to(dtype) for k, v in fx["state_dict"].items()}
    x = fx["x"] if x is None else x
    return O.fusion_encoder(p, "voxel_fusion.", x.to(dtype), fx["voxels_pos"].to(dtype), fx["vmask"], fx["dims"],
                            fx["num_modes"], collect=collect), p
