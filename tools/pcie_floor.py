"""How close is bench.py's e2e step to the PCIe floor? Times the step's H2D copies alone (pinned -> device, same sizes,
same three-buffer split and as one packed buffer) and the D2H of the survivor lists."""
import json
import time

import torch

H2D = [8294400, 1280000, 120004]     # depth 1920x1080 f32, 10k x 128 B entity data, 12 B header + 10k x 12 B entity draws
D2H = 1937188
dev = torch.device("cuda:0")
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in H2D]
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for n in H2D]
hp = torch.empty(sum(H2D), dtype=torch.uint8).pin_memory()
dp = torch.empty(sum(H2D), dtype=torch.uint8, device=dev)
ho = torch.empty(D2H, dtype=torch.uint8).pin_memory()
do = torch.empty(D2H, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def split():
    with torch.cuda.stream(s_in):
        for h, d in zip(hs, ds):
            d.copy_(h, non_blocking=True)


def packed():
    with torch.cuda.stream(s_in):
        dp.copy_(hp, non_blocking=True)


def both():
    packed()
    with torch.cuda.stream(s_out):
        ho.copy_(do, non_blocking=True)


out = {"h2d_split_us": run(split), "h2d_packed_us": run(packed), "h2d_packed_plus_d2h_us": run(both)}
out["h2d_GBs_packed"] = sum(H2D) / out["h2d_packed_us"] / 1e3
print(json.dumps(out))
