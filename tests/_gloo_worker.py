"""world_size-2 gloo worker for tests/test_host_cpu.py: shards clips over ranks exactly as bench.py does."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.tim_oracle import TIMOracle                     # noqa: E402
from tim_b200.config import TIMConfig                       # noqa: E402
from tim_b200.dist import gather_rows, max_over_ranks, shard_range   # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = TIMConfig(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                    num_layers=2, num_feats=6)
    B, Qv, Qa = 5, 3, 2
    sd = synth_state_dict(cfg, 0)
    inp = synth_inputs(cfg, B, Qv, Qa, 99)
    lo, hi = shard_range(B, rank, world)
    o = TIMOracle(cfg, sd, np.float32)
    mine = o.forward(inp["vis"][lo:hi], inp["aud"][lo:hi], inp["times"][lo:hi], Qv, Qa)
    act = gather_rows(torch.from_numpy(mine["action"]), rows_per_clip=Qv, B=B)
    t = max_over_ranks(float(rank + 1))
    if rank == 0:
        full = o.forward(inp["vis"], inp["aud"], inp["times"], Qv, Qa)
        assert np.array_equal(act.numpy(), full["action"])
        assert t == float(world)
        print("GLOO_SHARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
