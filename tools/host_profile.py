"""Host-side cost of enqueueing one stage-1 step: cProfile over 3 steps (the GPU runs asynchronously underneath)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from genhancer_b200 import optim

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
step, clip_vis, dit, vae = bench.build_models(336, dev)
trainable = list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
groups = optim.flatten(trainable)
opt = optim.FusedAdamW(groups, lr=1e-4, engine_managed=[dit])
x = torch.rand(32, 3, 336, 336, device=dev)


def train_step():
    loss = step(x)
    loss.backward()
    opt.step(1.0)
    opt.zero_grad()


for _ in range(3):
    train_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    train_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / 3:.1f} ms/step, with sync {1e3 * (t2 - t0) / 3:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    train_step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
