"""BASELINE configs[2]: HiP-AD stage-2 training step, ResNet-50 + FPN, bs per GPU, batch-sharded DDP
(``MMDistributedDataParallel``, apis/mmdet_train.py:97-102 of the reference; NCCL all-reduce of ~98 M parameters' gradients).

TEST / BENCH INFRASTRUCTURE.  Only the aggregation path is this repository's product; the rest of the step is the
reference's own decoder (harness/decoder.py) behind a torchvision ResNet-50 and a minimal FPN stand-in (SURVEY.md §7
step 6 allows it: what is exercised is the DFA forward + backward inside a real autograd graph and DDP's bucketed
gradient all-reduce), with a scalar surrogate loss over every decoder output (targets / Hungarian matching are out of
scope) and an AdamW step.  The backbone runs under bf16 autocast and hands fp32 feature maps to
``feature_maps_format`` like the reference's ``extract_feat`` (sparse_detector.py:66-91, ``auto_fp16(out_fp32=True)``).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from harness import decoder as HD  # noqa: E402


class ResNet50FPN(nn.Module):
    """torchvision ResNet-50 (random init) + FPN: 4 levels of 256 channels at strides 4/8/16/32
    (the reference: mmdet ResNet depth 50 out_indices (0,1,2,3) + FPN num_outs 4, hipad_b2d_stage2.py:113-137)."""

    def __init__(self, out_channels=256):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.stages = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.lateral = nn.ModuleList([nn.Conv2d(c, out_channels, 1) for c in (256, 512, 1024, 2048)])
        self.output = nn.ModuleList([nn.Conv2d(out_channels, out_channels, 3, padding=1) for _ in range(4)])

    def forward(self, img):
        bs, cams = img.shape[:2]
        x = self.stem(img.flatten(0, 1))
        feats = []
        for stage in self.stages:
            x = stage(x)
            feats.append(x)
        lat = [l(f) for l, f in zip(self.lateral, feats)]
        for i in range(3, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[-2:], mode="nearest")
        outs = [o(l) for o, l in zip(self.output, lat)]
        return [o.float().unflatten(0, (bs, cams)) for o in outs]


class TrainModel(nn.Module):
    def __init__(self, decoder, share_gradient=True):
        super().__init__()
        self.backbone = ResNet50FPN()
        self.decoder = decoder
        self.share_gradient = share_gradient

    def forward(self, img, metas):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            levels = self.backbone(img)
        ops = self.decoder._hipad_ops
        fm = ops.feature_maps_format(levels)
        if self.share_gradient and hasattr(ops, "share_feature_gradient"):
            fm = [ops.share_feature_gradient(fm[0]), fm[1], fm[2]]      # one dense feature gradient per step
        out = self.decoder(img, fm, metas)
        loss = img.new_zeros(())
        for t in HD.flatten_outputs(out).values():
            if t.dtype.is_floating_point and t.requires_grad:
                loss = loss + t.float().pow(2).mean()
        return loss


def make_batch(bs, hw, step, device, seed=0):
    rng = np.random.default_rng(seed * 1000 + step)
    img = torch.from_numpy(rng.standard_normal((bs, 6, 3, hw[0], hw[1]), dtype=np.float32)).to(device)
    proj, wh = HD.camera_matrices(hw)
    metas = dict(
        projection_mat=torch.from_numpy(np.tile(proj[None], (bs, 1, 1, 1))).to(device),
        image_wh=torch.from_numpy(np.tile(wh[None], (bs, 1, 1))).to(device),
        timestamp=torch.full((bs,), 0.5 * step, dtype=torch.float64, device=device),
        img_metas=[dict(T_global=np.eye(4), T_global_inv=np.eye(4), timestamp=0.5 * step) for _ in range(bs)],
        target_point=torch.from_numpy(np.tile(np.array([[0.0, 30.0]], dtype=np.float32), (bs, 1))).to(device),
        gt_ego_fut_cmd=torch.from_numpy(np.tile(np.eye(6, dtype=np.float32)[3:4], (bs, 1))).to(device),
    )
    return img, metas


def time_train_step(variant="ours_module", bs=4, hw=(352, 640), steps=5, warmup=2, device=None, world=1, rank=0,
                    share_gradient=True):
    """Times `steps` optimizer steps (forward + backward + DDP gradient all-reduce + AdamW) with CUDA events.
    Returns a dict; `ms_per_step` is this rank's (the caller takes the MAX over ranks)."""
    import torch.distributed as dist
    import hipad_b200
    device = device or torch.device("cuda", torch.cuda.current_device())
    dec = HD.build_decoder(variant, hw=hw, device=device, seed=0)
    dec.train()
    model = TrainModel(dec, share_gradient=share_gradient and variant != "reference").to(device)
    model.train()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    net = model
    if world > 1:
        net = nn.parallel.DistributedDataParallel(model, device_ids=[device.index], broadcast_buffers=False,
                                                  find_unused_parameters=True)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-3)
    ops = hipad_b200.ops
    batches = [make_batch(bs, hw, s, device, seed=rank) for s in range(4)]
    HD.reset(dec)

    def one(step):
        img, metas = batches[step % 4]
        metas = dict(metas, timestamp=metas["timestamp"] + 2.0 * (step // 4))
        opt.zero_grad(set_to_none=True)
        loss = net(img, metas)
        loss.backward()
        opt.step()
        return loss

    for s in range(warmup):
        one(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rec = ops.KernelTimer() if hasattr(ops, "KernelTimer") else None
    ev = []
    if rec is not None:
        rec.__enter__()
    for s in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = one(warmup + s)
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    if rec is not None:
        rec.__exit__(None, None, None)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    res = {"variant": variant, "bs_per_gpu": bs, "input_hw": list(hw), "steps": steps, "ms_per_step": round(ms, 2),
           "trainable_params_M": round(n_params / 1e6, 2), "loss": float(loss),
           "backbone": "torchvision ResNet-50 + FPN stand-in, bf16 autocast, random init",
           "optimizer": "AdamW step inside the timed region", "shared_feature_gradient": model.share_gradient}
    if rec is not None:
        k = rec.summary()
        res["dfa_kernel_ms_per_step"] = {kk: round(v / steps, 3) for kk, v in k.items()}
        res["dfa_share_of_step"] = round(sum(k.values()) / steps / ms, 4)
    if world > 1:
        # what the gradient all-reduce costs when nothing overlaps it: one flat fp32 buffer of the same size
        flat = torch.zeros(n_params, device=device)
        for _ in range(2):
            dist.all_reduce(flat)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            dist.all_reduce(flat)
        b.record()
        torch.cuda.synchronize()
        ar = a.elapsed_time(b) / 3
        res["allreduce_standalone_ms"] = round(ar, 3)
        res["allreduce_bus_gbs"] = round(2 * (world - 1) / world * n_params * 4 / (ar * 1e-3) / 1e9, 1)
    del net, model, opt
    torch.cuda.empty_cache()
    return res
