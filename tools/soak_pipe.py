"""Long bit-exact comparison of the multi-warp COUNTER builds (pipeline = 1: three roles, 2: two roles) with the one-warp
build (pipeline = 0) on low-occupancy batches: same keys, same sweeps, states / n / cutoffs / stream positions / operator
strings must be identical after every chunk.  Usage: python tools/soak_pipe.py [sweeps]"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_COUNTER, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
CASES = [("square L=64 h=0 beta ladder", lattices.square_periodic(64, -1.0), 3.04, 0.0, np.geomspace(0.25, 16.0, 512).repeat(2), 1024),
         ("triangular L=24 h=0.2 beta=8", lattices.triangular_periodic(24, 1.0), 1.0, 0.2, 8.0, 600)]
for name, edges, gamma, h, betas, R in CASES:
    keys = 0x50AC0000 + np.arange(R, dtype=np.uint64)
    nv = 1 + max(max(a, b) for (a, b), _ in edges)
    gs = []
    for pl in (0, 1, 2):
        g = QmcIsingGraph(edges, gamma, h, nv, keys, betas, mode=MODE_COUNTER)
        g.set_option("pipeline", pl)
        gs.append(g)
    done = 0
    for chunk in [1, 2, 5] + [16] * ((sweeps - 8) // 16):
        for g in gs:
            g.timesteps(chunk, betas)
        done += chunk
        ref = gs[0]
        for pl, g in zip((1, 2), gs[1:]):
            assert np.array_equal(ref.state_ref(), g.state_ref()), (name, pl, done, "state")
            assert np.array_equal(ref.get_n(), g.get_n()) and np.array_equal(ref.get_cutoff(), g.get_cutoff()), (name, pl, done, "n / cutoff")
            assert np.array_equal(ref.rng_cursors(), g.rng_cursors()), (name, pl, done, "cursor")
            for r in range(0, R, max(1, R // 24)):
                assert zlib.crc32(ref.dump_ops(r).tobytes()) == zlib.crc32(g.dump_ops(r).tobytes()), (name, pl, done, r)
    assert all(gs[1].verify(r) for r in range(0, R, 50))
    print(f"{name}: {done} sweeps, three builds identical, <n> = {gs[0].get_n().mean():.0f}, launches {gs[1].launch_count()}")
    for g in gs:
        g.close()
