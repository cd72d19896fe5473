"""CPU (gloo, world_size 2 and 4) tests of the multi-GPU host logic: episode sharding and the data-parallel gradient
all-reduce that replaces nn.DataParallel's reduce_add (meta_tuning_sun_m/train_meta.py:128-129)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sun_oracle as O
from sunb200.dist import GradAllReducer, broadcast_module_state, shard_episodes, shard_range


def test_shard_range_covers_everything():
    for n in (600, 8, 75, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(600, 3, 8) == (225, 300)        # BASELINE.json configs[1]: 75 episodes per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _episode_grads(sd, xs, xq, way, query):
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    ep = xs.shape[0]
    logits = O.meta_baseline_forward(params, xs, xq, training=True, bn_state=O.BNState()).view(-1, way)
    loss = O.cross_entropy(logits, O.make_nk_label(way, query, ep))
    loss.backward()
    names = [k for k, v in params.items() if isinstance(v, torch.Tensor) and v.requires_grad]
    return names, [params[k].grad for k in names], loss.item()


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    way, shot, query, ep = 3, 1, 1, world                       # one episode per rank (8 episodes on 8 GPUs in SUN-M)
    sd = O.init_meta_baseline_state_dict(5)
    data = O.make_episode_images(900, ep * way, shot + query)
    xs, xq = O.split_shot_query(data, way, shot, query, ep)
    my_s, my_q = shard_episodes(xs, xq, rank, world)              # the DataParallel scatter along the episode axis
    names, grads, loss = _episode_grads(sd, my_s, my_q, way, query)
    holders = [torch.nn.Parameter(torch.zeros_like(g)) for g in grads]
    red = GradAllReducer(holders)
    red.attach()
    for h, g in zip(holders, grads):
        h.grad.copy_(g)
    red.all_reduce_mean()
    # broadcast: rank 0 state is authoritative
    lin = torch.nn.Linear(4, 4)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
    broadcast_module_state(lin)
    assert float(lin.weight[0, 0]) == 1.0
    # rank 0's BatchNorm statistics are authoritative after sync_bn_buffers (parameters stay untouched)
    from sunb200.dist import sync_bn_buffers
    bn = torch.nn.BatchNorm2d(3)
    with torch.no_grad():
        bn.running_mean.fill_(float(rank + 1))
        bn.weight.fill_(float(rank + 5))
    sync_bn_buffers(bn)
    assert float(bn.running_mean[0]) == 1.0 and float(bn.weight[0]) == float(rank + 5)
    if rank == 0:
        torch.save({"names": names, "avg": [h.grad.clone() for h in holders], "loss": loss}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_gradients_match_shard_average(tmp_path, world):
    out = str(tmp_path / "avg.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    way, shot, query, ep = 3, 1, 1, world
    sd = O.init_meta_baseline_state_dict(5)
    data = O.make_episode_images(900, ep * way, shot + query)
    xs, xq = O.split_shot_query(data, way, shot, query, ep)
    per_shard = [_episode_grads(sd, xs[r:r + 1], xq[r:r + 1], way, query) for r in range(world)]
    for i, name in enumerate(got["names"]):
        ref = sum(per_shard[r][1][i] for r in range(world)) / world
        assert torch.allclose(got["avg"][i], ref, rtol=1e-5, atol=1e-7), name
