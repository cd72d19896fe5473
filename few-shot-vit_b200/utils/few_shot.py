"""Episode index utilities (reference: test_phase/utils/few_shot.py:4-16).  Pure view/split/arange work on the
caller's device -- index arithmetic only, bit-exact by construction."""
import torch


def split_shot_query(data, way, shot, query, ep_per_batch=1):
    """data [ep*way*(shot+query), *img] in sampler order -> (x_shot [ep,way,shot,*img], x_query [ep,way*query,*img]).
    Image ((e*way + c)*(shot+query) + j) is shot j of class c when j < shot, else query c*query + (j-shot)."""
    img = tuple(data.shape[1:])
    grid = data.reshape(ep_per_batch, way, shot + query, *img)
    x_shot = grid.narrow(2, 0, shot).contiguous()
    x_query = grid.narrow(2, shot, query).contiguous().reshape(ep_per_batch, way * query, *img)
    return x_shot, x_query


def make_nk_label(n, k, ep_per_batch=1):
    """int64 labels: class id repeated k times, tiled once per episode."""
    return torch.arange(n).repeat_interleave(k).repeat(ep_per_batch)
