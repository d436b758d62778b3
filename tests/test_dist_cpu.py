"""CPU (gloo, world_size 2): the bucketed gradient all-reduce used for data-parallel training."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gdr_net_b200 import GDRN as G
from gdr_net_b200.config import a6_config
from gdr_net_b200.dist import SEGMENTS, GradAllReducer, segment_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_segments_cover_flat_buffer_contiguously():
    (model, _opt) = G.build_model_optimizer(a6_config(device="cpu"))
    named = list(model.named_parameters())
    b = segment_bounds(named, SEGMENTS)
    total = sum(p.numel() for _, p in named)
    spans = sorted(b.values())
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0
    # reverse-forward order of completion: pnp first ... stem last
    assert b["pnp_net"][0] > b["rot_head_net"][0] > b["backbone.layer4"][0] > b["backbone.conv1"][0]
    assert (b["pnp_net"][1] - b["pnp_net"][0]) * 4 > 36e6 and total * 4 > 140e6  # 140.2 MB fp32 per step


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    named = [("backbone.conv1.weight", torch.zeros(10)), ("backbone.bn1.weight", torch.zeros(3)),
             ("backbone.layer1.0.conv1.weight", torch.zeros(7)), ("backbone.layer2.0.conv1.weight", torch.zeros(5)),
             ("backbone.layer3.0.conv1.weight", torch.zeros(4)), ("backbone.layer4.0.conv1.weight", torch.zeros(6)),
             ("rot_head_net.features.0.weight", torch.zeros(8)), ("pnp_net.fc1.weight", torch.zeros(9))]
    flat = torch.arange(52, dtype=torch.float32) * (rank + 1)
    red = GradAllReducer(flat, named, use_side_stream=False)
    # the engine calls the hook in reverse-forward order
    for stage in ("pnp_net", "rot_head_net", "backbone.layer4", "backbone.layer3", "backbone.layer2", "backbone.stem"):
        red(None, stage)
    red.finish()
    q.put((rank, flat.tolist(), red.bytes_reduced))  # plain lists: no shared-memory tensor handles that die with the worker
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_mean_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(2)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    expect = torch.arange(52, dtype=torch.float32) * 1.5  # mean of x*1 and x*2
    for rank, flat, nbytes in res:
        assert torch.allclose(torch.tensor(flat), expect)
        assert nbytes == 52 * 4
