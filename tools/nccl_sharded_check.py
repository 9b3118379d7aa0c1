"""Multi-GPU check of ttcr_b200.distributed on real GPUs (NCCL): run under torchrun, one rank per GPU.
Every rank gets the same receiver times as a single-GPU solve of the same sources (bit for bit)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d
from ttcr_b200.distributed import raytrace_sharded

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 96
x = np.linspace(0.0, 20.0, n)
X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
s = ((1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z)).astype(np.float32)
rng = np.random.default_rng(7)
src = rng.uniform(0.5, 19.5, (5, 3))
rcv = rng.uniform(0.5, 19.5, (33, 3))
g = Grid3d(x, x, x, n_threads=2, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32, device=local)
tt, its = raytrace_sharded(g, src, rcv, s if rank == 0 else None)
g1 = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32, device=local)
g1.set_slowness(s)
ref = np.stack([g1.raytrace(src[i:i + 1], rcv) for i in range(len(src))])
ok = np.array_equal(tt, ref)
print(f"rank {rank}/{dist.get_world_size()}: sharded == local: {ok}, niter {its[:, 0].tolist()}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
