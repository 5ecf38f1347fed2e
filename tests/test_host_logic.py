"""CPU: host-side logic of the drop-in module and the data-parallel helpers (no kernels are launched)."""
import random

import pytest
import torch

from scat_b200 import dp, synth
from tests.util import StubBackbone, make_opt


def _net(**kw):
    from scat_b200.hand_net import EncoderTransformer
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    return EncoderTransformer(make_opt(**kw), mean, backbone=StubBackbone())


def test_state_dict_keys_match_reference_layout():
    # exact key set and shapes (probed from the reference state_dict)
    for heads in (8, 4):
        net = _net(heads=heads)
        sd = net.state_dict()
        expect = synth.head_param_shapes(heads)
        assert [k for k in sd if k != "positionalEncoding.pe"] == list(expect)
        for k, shape in expect.items():
            assert tuple(sd[k].shape) == shape, k
        assert tuple(sd["positionalEncoding.pe"].shape) == (1, 21, 784)
        assert [n for n, _ in net.named_parameters()] == list(expect)
        assert len(net.head_parameters()) == 35
        assert sum(p.numel() for p in net.head_parameters()) == {8: 3795099, 4: 1687899 + 0}.get(heads, None) or heads == 4


def test_full_backbone_keys_present():
    from scat_b200.hand_net import EncoderTransformer
    net = EncoderTransformer(make_opt(), torch.zeros(1, 66))
    keys = net.state_dict().keys()
    for k in ("main_encoder.conv1.weight", "main_encoder.layer2.3.conv3.weight", "main_encoder.fc1.weight",
              "main_encoder.layer1.0.downsample.0.weight", "main_encoder.bn1.running_mean"):
        assert k in keys
    assert "main_encoder.fc.weight" not in keys


def test_attributes_and_mask_rate_window():
    net = _net(mask_rate=0.2)
    for attr in ("main_encoder", "conv1x1_channel_reduction", "transformer", "positionalEncoding", "mask_token",
                 "regressor", "mean_params", "pl", "iteration", "pos_embed", "mask_rate", "full_content"):
        assert hasattr(net, attr)
    assert "mean_params" not in dict(net.named_buffers())          # plain attribute, hand_net.py:321
    for rate, n in ((0.05, 0), (0.1, 2), (0.2, 4), (0.5, 10), (0.9, 18), (0.95, 0)):
        net.mask_rate = rate
        assert len(net._draw_mask()) == n
    net.mask_rate = 0.2
    random.seed(0)
    assert net._draw_mask() == [10, 19, 17, 14]
    net.mask_rate = 0.0
    random.seed(3); a = random.random(); random.seed(3)
    net._draw_mask()
    assert random.random() == a                                     # no RNG consumed when masking is off


def test_parameter_containers_refuse_eager_fallback():
    net = _net()
    with pytest.raises(RuntimeError, match="parameters only"):
        net.transformer.layers[0][0](torch.zeros(1, 21, 784))
    with pytest.raises(RuntimeError):
        net.positionalEncoding(torch.zeros(1, 21, 784))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            net.pl = False
            net.forward_features(torch.zeros(2, 1024), torch.zeros(2, 512, 28, 28))


def test_head_config_descriptor():
    from scat_b200.functional import HeadConfig
    d = HeadConfig(heads=4, iteration=2, pos_embed=False, n_masked=10, pl_reg=True, precision="fp32").desc(7)
    assert (d.batch, d.n_tokens, d.channels, d.token_dim, d.heads, d.iteration, d.pos_embed, d.n_masked, d.pl_reg,
            d.precision, d.main_feat_dim, d.n_out) == (7, 21, 512, 784, 4, 2, 0, 10, 1, 0, 1024, 66)


def test_shard_batch_partitions():
    for gb, w in ((768, 8), (96, 1), (10, 4), (3, 8)):
        spans = [dp.shard_batch(gb, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_flat_grad_bucket_views():
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    b = dp.FlatGradBucket(ps)
    assert b.offsets == [0, dp.ALIGN] and b.flat.numel() == 2 * dp.ALIGN       # every tensor on a 128-byte boundary
    assert b.payload_bytes() == 68 and b.nbytes() == 8 * dp.ALIGN
    b.flat.fill_(2.0)
    assert torch.all(ps[0].grad == 2.0) and ps[1].grad.data_ptr() == b.flat[dp.ALIGN:].data_ptr()
    assert b.all_reduce() is None                                   # single process: no-op


def test_flatten_parameters_is_value_preserving_and_idempotent():
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(1, 1, 7))]
    vals = [p.detach().clone() for p in ps]
    flat = dp.flatten_parameters(ps)
    A = dp.ALIGN
    assert flat.numel() == 3 * A and all(torch.equal(p.data, v) for p, v in zip(ps, vals))
    assert all(p.data_ptr() % (4 * A) == flat.data_ptr() % (4 * A) for p in ps)
    assert float(flat[12:A].abs().sum()) == 0.0               # gaps are zero
    again = dp.flatten_parameters(ps)
    assert again.data_ptr() == flat.data_ptr() and again.numel() == 3 * A
    flat[A] = 42.0
    assert ps[1].data[0] == 42.0
    bucket = dp.FlatGradBucket(ps)
    g = dp.flat_gradients(ps)
    assert g is not None and g.data_ptr() == bucket.flat.data_ptr() and g.numel() == 3 * A
    ps[1].grad = torch.zeros(5)                               # a replaced .grad breaks the flat layout
    assert dp.flat_gradients(ps) is None
    ps[1].grad = None
    assert dp.flat_gradients(ps) is None


def test_head_adam_has_no_cpu_path():
    from scat_b200.optim import HeadAdam
    with pytest.raises(RuntimeError):
        HeadAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-4)
    with pytest.raises(ValueError):
        HeadAdam([torch.nn.Parameter(torch.zeros(4))], lr=-1.0)


def test_coarse_module_state_dict_matches_reference_layout():
    """EncoderTransformerCoarse (hand_net.py:216-259): key set, order and shapes as probed from the reference by
    oracle/make_golden.py coarse (which asserts the same list against the unmodified reference's state_dict)."""
    from scat_b200.hand_net import EncoderTransformerCoarse
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    net = EncoderTransformerCoarse(make_opt(pl_reg=False), mean, backbone=StubBackbone())
    expect = synth.coarse_param_shapes()
    sd = net.state_dict()
    assert [k for k in sd if k != "positionalEncoding.pe"] == list(expect)
    for k, shape in expect.items():
        assert tuple(sd[k].shape) == shape, k
    assert len(net.head_parameters()) == 35
    # C-ABI slot order: post-attention LayerNorm first, then qkv / out projection (include/scat_b200.h)
    hp = net.head_parameters()
    assert hp[2] is net.transformer.layers[0][1].norm.weight and hp[4] is net.transformer.layers[0][0].to_qkv.weight
    assert tuple(hp[33].shape) == (3, 1027)
    with pytest.raises(RuntimeError, match="does not require grad"):       # pl_reg: the reference fails under no_grad too
        EncoderTransformerCoarse(make_opt(pl_reg=True), mean, backbone=StubBackbone()).forward_features(None, None)
