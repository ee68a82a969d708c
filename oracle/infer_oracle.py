"""CPU restatement of the reference's sliding-window inference (TEST INFRASTRUCTURE ONLY — see segtran_oracle.py header).

Follows code/test_util3d.py:93-184 (test_single_case) and code/dataloaders/datasets3d.py:43-61
(make_brats_pred_consistent, is_conservative=False), plain PyTorch on CPU.  Pinned by tests/golden/infer_sw.pt, which
oracle/gen_golden.py produces with the reference's own function.  One deliberate difference: the reference's padding
branch passes its pads to F.pad in the wrong dimension order (test_util3d.py:119-120: the tuple starts with the LAST
dim, so the H/W/D pads land on W/H/C) and cannot run; this restatement pads H, W, D as the surrounding code intends."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def make_brats_pred_consistent(preds_soft):            # datasets3d.py:53-59
    out = preds_soft.clone()
    out[2] = torch.max(preds_soft[1:], dim=0)[0]       # If TC (or ET) then WT
    out[3] = torch.max(preds_soft[[1, 3]], dim=0)[0]   # If ET then TC
    return out


def test_single_case(net, image, orig_patch_size, input_patch_size, batch_size, stride_xy, stride_z, task_name, net_type,
                     num_classes):
    C, H, W, D = image.shape
    dx, dy, dz = orig_patch_size
    h_pad, w_pad, d_pad = max(dx - H, 0), max(dy - W, 0), max(dz - D, 0)          # :99-114
    add_pad = (h_pad + w_pad + d_pad) > 0
    hl, wl, dl = h_pad // 2, w_pad // 2, d_pad // 2
    if add_pad:
        image = F.pad(image, (dl, d_pad - dl, wl, w_pad - wl, hl, h_pad - hl), mode='constant', value=0)
    C, H2, W2, D2 = image.shape
    sx = math.ceil((H2 - dx) / stride_xy) + 1                                     # :125-127
    sy = math.ceil((W2 - dy) / stride_xy) + 1
    sz = math.ceil((D2 - dz) / stride_z) + 1
    preds_soft = torch.zeros((num_classes,) + tuple(image.shape[1:]))
    cnt = torch.zeros_like(image[0])
    for x in range(sx):                                                           # :132-162
        xs = min(stride_xy * x, H2 - dx)
        yzs, patches = [], []
        for y in range(sy):
            ys = min(stride_xy * y, W2 - dy)
            for z in range(sz):
                zs = min(stride_z * z, D2 - dz)
                patches.append(image[:, xs:xs + dx, ys:ys + dy, zs:zs + dz])
                yzs.append((ys, zs))
                if len(patches) == batch_size or (y == sy - 1 and z == sz - 1):
                    batch = F.interpolate(torch.stack(patches, 0), size=input_patch_size, mode='trilinear', align_corners=False)
                    with torch.no_grad():
                        scores = net(batch)
                    if net_type == 'unet':
                        scores = scores[1]
                    scores = F.interpolate(scores, size=orig_patch_size, mode='trilinear', align_corners=False)
                    probs = torch.sigmoid(scores)
                    for i, (ys_i, zs_i) in enumerate(yzs):
                        preds_soft[:, xs:xs + dx, ys_i:ys_i + dy, zs_i:zs_i + dz] += probs[i]
                        cnt[xs:xs + dx, ys_i:ys_i + dy, zs_i:zs_i + dz] += 1
                    patches, yzs = [], []
    preds_soft = preds_soft / cnt.unsqueeze(0)                                    # :164
    if task_name == 'brats':                                                      # :165-170
        preds_soft = make_brats_pred_consistent(preds_soft)
        preds_hard = torch.zeros_like(preds_soft)
        preds_hard[1:] = (preds_soft[1:] >= 0.5)
        preds_hard[0] = (preds_hard[1:].sum(dim=0) == 0)
    else:
        preds_hard = torch.argmax(preds_soft, dim=0)                              # :173
    if add_pad:                                                                   # :175-178
        preds_hard = preds_hard[..., hl:hl + H, wl:wl + W, dl:dl + D].clone()
        preds_soft = preds_soft[:, hl:hl + H, wl:wl + W, dl:dl + D].clone()
    return preds_hard, preds_soft
