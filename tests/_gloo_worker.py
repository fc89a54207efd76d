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
    # the training leg's one collective: gradients in ONE flat buffer, one all-reduce (sum / world) - against per-tensor averaging
    from tim_b200.dist import FlatGrads
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3))
    flat = FlatGrads(net.named_parameters())
    assert all(p.grad.data_ptr() == flat.view(n).data_ptr() for n, p in net.named_parameters())
    assert all(lo % FlatGrads.ALIGN_ELEMS == 0 for lo, _ in flat.offsets.values())
    x = torch.randn((4, 7), generator=torch.Generator().manual_seed(10 + rank))
    net(x).square().sum().backward()                        # autograd accumulates INTO the flat views
    mine = {n: p.grad.clone() for n, p in net.named_parameters()}
    assert flat.buffer.abs().sum() > 0
    flat.all_reduce()
    for n, p in net.named_parameters():
        parts = [torch.zeros_like(mine[n]) for _ in range(world)]
        dist.all_gather(parts, mine[n])
        assert torch.allclose(p.grad, sum(parts) / world, rtol=1e-6, atol=1e-7), n
        assert p.grad.data_ptr() == flat.view(n).data_ptr()
    # optimizer.zero_grad() (set_to_none=True, what both reference train loops call) detaches param.grad from the flat buffer:
    # (a) reattach() before the backward (what patch_model does) makes autograd accumulate into the views again;
    # (b) a backward WITHOUT it leaves foreign gradients, which all_reduce() copies in and re-binds before the collective.
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for variant in ("reattach_before_backward", "foreign_grads"):
        opt.zero_grad()
        assert all(p.grad is None for p in net.parameters())
        if variant == "reattach_before_backward":
            assert flat.reattach() == len(list(net.parameters())) and float(flat.buffer.abs().sum()) == 0.0
        net(x).square().sum().backward()
        mine = {n: p.grad.clone() for n, p in net.named_parameters()}
        flat.all_reduce()
        for n, p in net.named_parameters():
            parts = [torch.zeros_like(mine[n]) for _ in range(world)]
            dist.all_gather(parts, mine[n])
            assert torch.allclose(p.grad, sum(parts) / world, rtol=1e-6, atol=1e-7), (variant, n)
            assert p.grad.data_ptr() == flat.view(n).data_ptr(), (variant, n)
    if rank == 0:
        print("GLOO_FLATGRAD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
