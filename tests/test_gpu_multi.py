"""-m gpu, needs >= 2 GPUs (skipped otherwise): the overlapped in-backward gradient all-reduce (GradComm) gives the same
averaged gradients as a single flat all-reduce after backward (GradAllReducer), and both ranks end up identical."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import torch.nn.functional as F
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import models
    import utils.few_shot as fs
    import sun_oracle as O
    from sunb200.dist import GradAllReducer, shard_episodes
    sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
    way, shot, query, ep = 3, 1, 2, 2
    data = O.make_episode_images(500, ep * way, shot + query).cuda()
    xs, xq = fs.split_shot_query(data, way, shot, query, ep_per_batch=ep)
    xs, xq = shard_episodes(xs, xq, rank, world)
    label = fs.make_nk_label(way, query, xs.shape[0]).cuda()
    grads = {}
    for mode in ("flat", "flat_again", "overlap"):
        model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={})
        model.load_state_dict(sd)
        model = model.cuda().train()
        if mode == "overlap":
            model.encoder.enable_data_parallel()
            red = GradAllReducer([model.temp])
        else:
            red = GradAllReducer(model.parameters())
        loss = F.cross_entropy(model(xs, xq).view(-1, way), label)
        red.attach()
        red.flat.zero_()
        loss.backward()
        red.all_reduce_mean()
        torch.cuda.synchronize()
        grads[mode] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    zero_grad = ("patch_embed2.proj.bias", "patch_embed2.norm.bn.bias", "patch_embed3.proj.bias", "patch_embed3.norm.bn.bias")
    def compare(x, y):
        rels = []
        for n in grads[x]:
            if n.endswith(zero_grad):      # analytically zero under batch-stat BN: pure rounding noise, no relative error
                continue
            a, b = grads[x][n], grads[y][n]
            rels.append((((a - b).norm() / (a.norm() + 1e-12)).item(), n))
        rels.sort(reverse=True)
        return rels
    noise, rels = compare("flat", "flat_again"), compare("flat", "overlap")
    torch.save({"worst": rels[0][0], "median": rels[len(rels) // 2][0], "top": rels[:4],
                "noise_worst": noise[0][0], "noise_median": noise[len(noise) // 2][0],
                "probe": grads["overlap"]["encoder.stage3.2.attn.proj.weight"].cpu()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_overlapped_allreduce_matches_flat_allreduce(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    # Independent runs of the same step differ by the fp32 atomic-add order (BN statistics, split-K wgrad), which flips a few
    # bf16 roundings downstream.  The overlapped schedule must sit inside that run-to-run noise (measured with two flat runs),
    # far below the bf16-vs-fp32 tolerance of the step itself.
    print(f"flat vs flat: median {r0['noise_median']:.2e} worst {r0['noise_worst']:.2e}; "
          f"flat vs overlap: median {r0['median']:.2e} worst {r0['worst']:.2e}; top {r0['top']}")
    assert r0["median"] <= 3 * r0["noise_median"] + 1e-3 and r0["worst"] <= 3 * r0["noise_worst"] + 1e-2, r0["top"]
    assert torch.equal(r0["probe"], r1["probe"])          # ranks hold identical averaged gradients
