#!/usr/bin/env python
"""cfg 5 (BASELINE.json configs[4]): one FPN Mask R-CNN training
step (ResNet-50-FPN, random init, synthetic COCO-shaped batch; forward + losses +
backward + MomentumSGD update) with the pooling path swapped between

    b200         this package's fused kernels (one launch for box 7x7 + mask 14x14)
    torchvision  torchvision.ops.roi_align batched per level (library arm)
    per_roi      the reference's dispatch: one op call per RoI and size
                 (fpn_roi_mask_head.py:57-63,74-78), torchvision's kernel as the op

Prints one JSON line per back end: iter/s of the whole step (CUDA events, max of
nothing -- single GPU) and the pooling calls alone (fwd + bwd on the step's own
features and sampled RoIs).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chainer_maskrcnn_b200.model.fpn_maskrcnn_train_chain import (  # noqa: E402
    FPNMaskRCNNTrainChain, MaskRCNN, synthetic_batch)


def time_pooling(head, feats, rois, levels, scales, reps=10):
    feats = [f.detach().float().clone().requires_grad_(True) for f in feats]

    def once():
        if head.pooling == "b200":
            pb, pm = head.pool(feats, rois, levels, scales, train=True)
        else:
            pb = head._pool(feats, rois, levels, scales, head.roi_size_box)
            pm = head._pool(feats, rois, levels, scales, head.roi_size_mask)
        torch.autograd.backward([pb, pm], [torch.ones_like(pb), torch.ones_like(pm)])
        for f in feats:
            f.grad = None
    for _ in range(2):
        once()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pooling", default="b200,torchvision,per_roi")
    ap.add_argument("--sampling-ratio", type=int, default=2)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the dense layers (pooling stays fp32)")
    ap.add_argument("--tf32", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.backends.cudnn.allow_tf32 = args.tf32
    imgs, bboxes, labels, masks = synthetic_batch(args.batch, args.height, args.width, seed=5, device=dev)
    imgs = imgs.contiguous(memory_format=torch.channels_last)
    for pooling in args.pooling.split(","):
        torch.manual_seed(0)
        model = MaskRCNN(80, pooling=pooling, sampling_ratio=args.sampling_ratio)
        chain = FPNMaskRCNNTrainChain(model).to(dev).to(memory_format=torch.channels_last)
        chain.train()
        opt = torch.optim.SGD(chain.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)   # train.py:107-109

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp):
                loss = chain(imgs, bboxes, labels, masks, 1.0)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        last = chain.last
        pool_ms = time_pooling(model.head, last["features"], last["indices_and_rois"], last["levels"],
                               model.extractor.spatial_scales)
        print(json.dumps({
            "metric": "maskrcnn_fpn_train_step", "value": 1e3 / ms, "unit": "iter/s", "ms_per_iter": ms,
            "pooling": pooling, "pooling_fwd_bwd_ms_isolated": pool_ms,
            "config": {"workload": "cfg5 ResNet-50-FPN Mask R-CNN training step, random init, synthetic batch",
                       "batch": args.batch, "image": [args.height, args.width],
                       "sampled_rois": int(last["n_sample"]), "proposals": int(last["n_proposals"]),
                       "sampling_ratio": args.sampling_ratio,
                       "dense_dtype": "bf16 autocast" if args.amp else ("tf32" if args.tf32 else "fp32"),
                       "step": "forward + 5 losses + backward + MomentumSGD(wd 5e-4) update"},
            "loss": float(loss.detach()), "iters": args.iters, "warmup": args.warmup,
            "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
        del chain, model, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
