"""Soak test of the peer-memory gradient exchange: many optimiser steps over changing ragged groups on all GPUs
of the box (torchrun), with random per-rank host delays to skew the ranks against each other; at the end every
replica must hold the same bits and no exchange may have timed out.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/peer_soak.py --steps 20000
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import parallel as Pl

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20000)
ap.add_argument("--batch", type=int, default=128)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
params, bn = bench.load_weights()
eng = sstg.TrainEngine(params, bn, dev)
tr = Pl.DistributedTrainer(eng, exchange="peer", max_scenes=a.batch)
groups = []
for g in range(16):
    n = bench.synth_ragged_counts(a.batch if g % 4 else max(world, a.batch // 8), 300 + g)   # some tiny groups: empty shards
    o, p, _ = bench.synth_batch(int(n.sum()), 1, 500 + g)
    groups.append((n, o.to(dev), p.to(dev)))
rng = np.random.default_rng(1000 + rank)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for i in range(a.steps):
    n, o, p = groups[i % len(groups)]
    tr.step(n, o, p, a.batch, 1e-5)
    if rng.random() < 0.002:
        time.sleep(rng.random() * 0.003)            # skew this rank by up to 3 ms
    if i % 2000 == 1999:
        tr.peer.check()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tr.peer.check()
blob = torch.cat([eng.params, eng.bn])
both = [torch.empty_like(blob) for _ in range(world)]
dist.all_gather(both, blob)
same = all(torch.equal(both[0], b) for b in both)
finite = bool(torch.isfinite(blob).all())
if rank == 0:
    print(f"peer soak: world {world}, {a.steps} steps in {dt:.2f} s ({dt / a.steps * 1e6:.1f} us/step), "
          f"replicas bit-identical: {same}, finite: {finite}, epoch {tr.peer.epoch}")
assert same and finite
tr.peer.close()
dist.destroy_process_group()
