"""worker for tests/test_dist_cpu.py: world_size-2 gloo run of the sample-split + reduce plumbing, with the CPU oracle
standing in for the per-rank renderer (there is no GPU in the CPU test tier)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
from turner_b200 import dist as tdist, scenes  # noqa: E402

rank, local_rank, world = tdist.init_process_group("gloo")
sc = scenes.fixture("cornell_box")
o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
W, D, M, PPS = 48, 3, 3, 5
begin, stride = tdist.sample_split(rank, world)
cfg = ob.make_cfg(sc, W, D, M, PPS, rng_mode=1, seed=11, sample_begin=begin, sample_stride=stride)
part, _, st = o.render(cfg)
assert st.num_prim_rays == W * cfg.height * tdist.local_sample_count(PPS, rank, world)
acc = torch.from_numpy(part.copy())
rays = torch.tensor([st.num_rays], dtype=torch.int64)
tdist.reduce_accum(acc, root=0)
dist.reduce(rays, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    full, _, fst = o.render(ob.make_cfg(sc, W, D, M, PPS, rng_mode=1, seed=11))
    assert int(rays.item()) == fst.num_rays, (int(rays.item()), fst.num_rays)
    assert np.allclose(acc.numpy(), full, rtol=1e-5, atol=1e-6)
    print("DIST_OK world=%d rays=%d" % (world, fst.num_rays))
dist.barrier()
dist.destroy_process_group()
