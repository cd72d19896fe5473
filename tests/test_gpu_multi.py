"""-m gpu: two data-parallel ranks of the SUN-M meta-tuning step against the REFERENCE's 2-shard gradient average
(tests/golden/train_step_sunm.npz, written by the real reference modules run shard by shard on the CPU).

Each rank is one process running its 60-image shard with the overlapped in-backward gradient all-reduce (GradComm,
sunb200/dist.py).  With >= 2 GPUs the ranks sit on different devices and talk NCCL; on a 1-GPU box both ranks share
cuda:0 and use gloo (NCCL refuses two ranks on one device), so the test never skips and the same code path runs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, backend, overlap):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here,):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from sunb200.dist import GradAllReducer
    from test_gpu_train import sunm_shard_grads
    model, logits, loss = sunm_shard_grads(rank, device=f"cuda:{dev}")
    if overlap:
        model.encoder.enable_data_parallel()          # per-stage all-reduce inside the native backward, on a side stream
        red = GradAllReducer([model.temp])
    else:
        red = GradAllReducer(model.parameters())      # one flat bucket after backward
    loss.backward()
    red.all_reduce_mean()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters()}
    torch.save({"grads": grads, "loss": loss.item()}, os.path.join(out_dir, f"rank{rank}_{int(overlap)}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False], ids=["overlapped", "flat"])
def test_two_rank_average_vs_reference(tmp_path, golden_dir, overlap):
    import torch.multiprocessing as mp
    from test_gpu_train import compare_grads
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), backend, overlap), nprocs=2, join=True)
    g = np.load(os.path.join(golden_dir, "train_step_sunm.npz"))
    r0 = torch.load(tmp_path / f"rank0_{int(overlap)}.pt")
    r1 = torch.load(tmp_path / f"rank1_{int(overlap)}.pt")
    for s, r in ((0, r0), (1, r1)):
        assert abs(r["loss"] - float(g[f"s{s}.loss"])) <= 0.05 * float(g[f"s{s}.loss"])
    for n in r0["grads"]:                              # both ranks hold the same averaged gradients
        assert torch.equal(r0["grads"][n], r1["grads"][n]), n
    compare_grads(r0["grads"], g, "avg.", f"N=2 {backend} {'overlapped' if overlap else 'flat'}")
