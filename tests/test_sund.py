"""SUN-D head: the oracle restatement against the reference's own Network.py methods (CPU, where the reference checkout is
present), and -m gpu the device kernel against the oracle (which calls cv2.EMD, the solver the reference calls)."""
import importlib.util
import os
import sys
import types

import pytest
import torch

import sun_oracle as O

REF = "/root/reference/meta_tuning_sun_d/Models/models"


def _nodes(n_img, seed, n=13, C=512):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n_img, C, 1, 1, generator=g)
    return (base + 0.7 * torch.randn(n_img, C, n, 1, generator=g)).contiguous()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_oracle_matches_reference_network_methods():
    """get_weight_vector / normalize_feature / get_similiarity_map of the real DeepEMD class (device-agnostic methods,
    SURVEY.md Appendix E) == the oracle restatement."""
    pytest.importorskip("cv2")
    for name in ("qpth", "qpth.qp"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["qpth.qp"].QPFunction = object
    pkg = types.ModuleType("refsund")
    pkg.__path__ = [REF]
    sys.modules["refsund"] = pkg
    for stub in ("resnet", "nest", "visformer"):
        m = types.ModuleType(f"refsund.{stub}")
        m.ResNet = m.nest_micro_80 = m.nest_nano_80 = m.visformer_small_80 = object
        sys.modules[f"refsund.{stub}"] = m
    for f in ("emd_utils", "Network"):
        spec = importlib.util.spec_from_file_location(f"refsund.{f}", os.path.join(REF, f + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refsund.{f}"] = mod
        spec.loader.exec_module(mod)
    Net = sys.modules["refsund.Network"].DeepEMD
    args = types.SimpleNamespace(backbone="none", norm="center", metric="cosine", solver="opencv", temperature=12.5)
    net = Net.__new__(Net)
    torch.nn.Module.__init__(net)
    net.args = args
    proto, query = _nodes(5, 1), _nodes(7, 2)
    assert torch.allclose(net.get_weight_vector(query, proto), O.sund_weight_vector(query, proto), atol=1e-5)
    sim_ref = net.get_similiarity_map(net.normalize_feature(proto), net.normalize_feature(query))
    assert torch.allclose(sim_ref, O.sund_similarity_map(proto, query), atol=1e-6)
    # the per-pair solver call of the reference (emd_utils.emd_inference_opencv) == the oracle's flows
    emd = sys.modules["refsund.emd_utils"].emd_inference_opencv
    w1, w2 = O.sund_weight_vector(query, proto), O.sund_weight_vector(proto, query)
    sim = O.sund_similarity_map(proto, query)
    _, flows = O.sund_emd_logits(proto, query, return_flows=True)
    _, f00 = emd(1 - sim[3, 2], w1[3, 2], w2[2, 3])
    assert torch.allclose(torch.from_numpy(f00), flows[3, 2], atol=1e-6)


def test_oracle_emd_is_a_valid_transport_plan():
    pytest.importorskip("cv2")
    proto, query = _nodes(3, 3), _nodes(4, 4)
    logits, flows = O.sund_emd_logits(proto, query, return_flows=True)
    assert logits.shape == (4, 3) and (flows >= -1e-3).all()      # cv2.EMD works in float32
    assert torch.allclose(flows.sum(dim=(-1, -2)), torch.full((4, 3), 13.0), atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("W,Q,n", [(5, 15, 13), (5, 75, 13), (3, 4, 25), (2, 3, 1)])
def test_emd_head_kernel_vs_cv2(W, Q, n):
    pytest.importorskip("cv2")
    from sunb200 import sund
    proto, query = _nodes(W, 10 + n, n=n), _nodes(Q, 20 + n, n=n)
    ref, ref_flows = O.sund_emd_logits(proto, query, return_flows=True)
    got, flows = sund.emd_logits(proto.cuda(), query.cuda(), return_flows=True)
    got, flows = got.cpu(), flows.cpu()
    # the optimal objective (hence the logit) is unique; plans may differ on degenerate ties, so compare logits and feasibility
    assert (got - ref).abs().max().item() < 2e-3, (got - ref).abs().max().item()
    assert (flows >= -1e-6).all()
    # marginals: cv2.EMD works in float32 (its own plans miss the marginals by ~1e-4)
    assert torch.allclose(flows.sum(-1), ref_flows.sum(-1), atol=2e-3) and torch.allclose(flows.sum(-2), ref_flows.sum(-2), atol=2e-3)
    assert (got.argmax(1) == ref.argmax(1)).all()


@pytest.mark.gpu
def test_sund_episode_through_the_encoder():
    """5-way 1-shot 3-query SUN-D episode with 13 grid crops per image: node features from the native encoder, logits from
    the device EMD head, against the oracle's encoder + cv2.EMD."""
    pytest.importorskip("cv2")
    from sunb200 import sund
    sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    model = sund.DeepEMD(mode="encoder").cuda().eval()
    model.encoder.load_state_dict(enc_sd)
    n_img, crops = 5 + 15, 4                         # 4 crops per image keep the CPU oracle quick
    g = torch.Generator().manual_seed(77)
    protos = torch.randn(5, 1, 3, 80, 80, generator=g)
    imgs = torch.cat([protos + 0.5 * torch.randn(5, 1, 3, 80, 80, generator=g),
                      (protos + 0.5 * torch.randn(5, 3, 3, 80, 80, generator=g)).reshape(15, 1, 3, 80, 80)])[:, 0]
    x = (imgs[:, None] + 0.3 * torch.randn(n_img, crops, 3, 80, 80, generator=g)).contiguous()
    with torch.no_grad():
        data = model(x.cuda())                                       # [20, 512, 4, 1]
        model.mode = "meta"
        logits = model((data[:5].unsqueeze(0), data[5:])).cpu()
        _, feat = O.encoder_forward(sd, x.reshape(-1, 3, 80, 80), "encoder.")
        ref_nodes = feat.reshape(n_img, crops, 512).permute(0, 2, 1).unsqueeze(-1)
        ref = O.sund_emd_logits(ref_nodes[:5], ref_nodes[5:])
    label = torch.arange(5).repeat_interleave(3)
    print("SUN-D logits max |delta|", (logits - ref).abs().max().item(), "acc", (logits.argmax(1) == label).float().mean().item())
    assert (logits - ref).abs().max().item() < 0.25
    assert (logits.argmax(1) == ref.argmax(1)).float().mean().item() >= 0.93
