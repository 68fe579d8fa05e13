"""torchrun worker: parallel tempering over NCCL, one rank per GPU, checked against the oracle.
Every rank steps its block of slots on its GPU; after every tempering step rank r compares its local
configurations with a single-process oracle ladder (configurations move between fixed slots there)."""
import os
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from isingmontecarlo_b200 import MODE_FAST, lattices  # noqa: E402
from isingmontecarlo_b200.tempering import TemperingContainer  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
n_chains, n_betas = 2, 4 * world
betas = np.linspace(0.4, 2.0, n_betas)
S = n_chains * n_betas
keys = 0x55E00000 + np.arange(S, dtype=np.uint64)
edges = lattices.two_d_periodic_mixed(4)
tc = TemperingContainer(edges, 1.0, 0.0, 16, betas, n_chains=n_chains, rng_keys=keys, pt_key=0xABC, mode=MODE_FAST, device=local)
ladder = [[po.SseOracle(edges, 1.0, 0.0, 16, key=int(keys[c * n_betas + k])) for k in range(n_betas)] for c in range(n_chains)]
cursors = [0] * n_chains
swaps = 0
for step in range(10):
    tc.timesteps(2)
    for c in range(n_chains):
        for k in range(n_betas):
            ladder[c][k].timesteps(2, float(betas[k]), MODE_FAST)
    tc.tempering_step()
    for c in range(n_chains):
        s, cursors[c] = po.pt_step(ladder[c], betas, 0xABC + c, cursors[c])
        swaps += s
    g = tc.graph
    n, cut, cur, st = g.get_n(), g.get_cutoff(), g.rng_cursors(), g.state_ref()
    for s_local, slot in enumerate(tc.slots()):
        ref = ladder[slot // n_betas][slot % n_betas]
        assert int(n[s_local]) == ref.n and int(cut[s_local]) == ref.cutoff and int(cur[s_local]) == ref.cursor, (rank, step, slot)
        assert np.array_equal(st[s_local], ref.state()) and np.array_equal(g.dump_ops(s_local), ref.dump_ops())
    assert tc.get_total_swaps() == swaps
assert swaps > 0 and tc.verify()
states, energy = tc.timesteps_sample(6, 2, 3)
assert energy.shape == (S,) and np.all(np.isfinite(energy))
print(f"rank {rank}/{world} ok swaps={swaps}")
dist.barrier()
dist.destroy_process_group()
