#!/usr/bin/env python3
"""Time the policy CNN's rollout forward (no grad, train-mode BatchNorm) and update step per path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bbgpu.network import BlockBlastNetwork
torch.backends.cudnn.benchmark = True
dev = "cuda"
for n in (32768, 65536, 131072):
    x = (torch.rand(n, 4, 8, 8, device=dev) < 0.4).float().contiguous(memory_format=torch.channels_last)
    for fused in (False, True):
        net = BlockBlastNetwork().to(dev).to(memory_format=torch.channels_last).set_fused_bn(fused)
        net.train()
        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return net.trunk(x)
        for _ in range(3):
            fwd()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            fwd()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print("batch %6d fused_bn=%-5s forward %.2f ms" % (n, fused, (t1 - t0) / 5 * 1e3), flush=True)
        del net
