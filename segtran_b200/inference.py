"""Sliding-window inference of a whole volume: drop-in for the reference's
``test_util3d.test_single_case`` (code/test_util3d.py:93-184) — same signature, same window enumeration, same padding —
with the per-patch "sigmoid -> accumulate -> count" update and the final "average -> BraTS consistency -> threshold"
running as two library kernels (csrc/sx_infer.cu) and the two tri-linear resizes as the library's per-axis kernels.
No CPU fallback: the volume must live on the GPU."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops


def _resize(x, size):
    """F.interpolate(x, size, mode='trilinear', align_corners=False) via the library's per-axis kernels (no-op if equal)."""
    if tuple(x.shape[2:]) == tuple(size):
        return x
    return ops.resize_linear(x.float(), tuple(int(s) for s in size))


def test_single_case(net, image, orig_patch_size, input_patch_size, batch_size, stride_xy, stride_z, task_name, net_type,
                     num_classes):
    """image [C,H,W,D] (CUDA) -> (preds_hard, preds_soft), exactly as the reference's function of the same name."""
    ops._req_cuda(image)
    C, H, W, D = image.shape
    dx, dy, dz = orig_patch_size
    h_pad, w_pad, d_pad = max(dx - H, 0), max(dy - W, 0), max(dz - D, 0)
    add_pad = (h_pad + w_pad + d_pad) > 0
    hl_pad, hr_pad = h_pad // 2, h_pad - h_pad // 2
    wl_pad, wr_pad = w_pad // 2, w_pad - w_pad // 2
    dl_pad, dr_pad = d_pad // 2, d_pad - d_pad // 2
    if add_pad:
        image = F.pad(image, (dl_pad, dr_pad, wl_pad, wr_pad, hl_pad, hr_pad), mode='constant', value=0)
    C, H2, W2, D2 = image.shape
    sx = math.ceil((H2 - dx) / stride_xy) + 1
    sy = math.ceil((W2 - dy) / stride_xy) + 1
    sz = math.ceil((D2 - dz) / stride_z) + 1
    K = int(num_classes)
    dev = image.device
    preds_soft = torch.zeros((K, H2, W2, D2), device=dev, dtype=torch.float32)
    cnt = torch.zeros((H2, W2, D2), device=dev, dtype=torch.float32)
    st = ops._stream

    for x in range(sx):
        xs = min(stride_xy * x, H2 - dx)
        yzs_batch, test_patches = [], []
        for y in range(sy):
            ys = min(stride_xy * y, W2 - dy)
            for z in range(sz):
                zs = min(stride_z * z, D2 - dz)
                test_patches.append(image[:, xs:xs + dx, ys:ys + dy, zs:zs + dz])
                yzs_batch.append((ys, zs))
                if len(test_patches) == batch_size or (y == sy - 1 and z == sz - 1):
                    test_batch = _resize(torch.stack(test_patches, dim=0), input_patch_size)
                    with torch.no_grad():
                        scores_raw = net(test_batch)
                    if net_type == 'unet':
                        scores_raw = scores_raw[1]
                    scores_raw = _resize(scores_raw, orig_patch_size).float().contiguous()
                    for i, (ys_i, zs_i) in enumerate(yzs_batch):       # sequential launches: overlapping windows never race
                        L.call("sx_sw_accumulate", scores_raw[i].data_ptr(), K, dx, dy, dz, preds_soft.data_ptr(),
                               cnt.data_ptr(), H2, W2, D2, xs, ys_i, zs_i, st())
                    test_patches, yzs_batch = [], []

    brats = task_name == 'brats'
    hard = torch.empty((K, H2, W2, D2) if brats else (H2, W2, D2), device=dev, dtype=torch.float32)
    L.call("sx_sw_finalize", preds_soft.data_ptr(), cnt.data_ptr(), K, H2 * W2 * D2, 1 if brats else 0, hard.data_ptr(), st())
    preds_hard = hard if brats else hard.long()
    if add_pad:
        sl = (slice(hl_pad, hl_pad + H), slice(wl_pad, wl_pad + W), slice(dl_pad, dl_pad + D))
        preds_hard = (preds_hard[(slice(None),) + sl] if brats else preds_hard[sl]).clone()
        preds_soft = preds_soft[(slice(None),) + sl].clone()
    return preds_hard, preds_soft
