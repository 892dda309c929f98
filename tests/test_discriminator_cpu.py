"""CPU: the Discriminator mirror (stylegan2/discriminator.py) against golden vectors of the UNMODIFIED reference class
(tests/golden/make_golden_disc.py).  The module code is the product's; only the two op entry points are replaced by the
oracle's CPU statements of the same ops, so this checks the host logic -- module tree, state_dict keys, padding rules,
minibatch-stddev, scaling -- without a GPU (tests/test_model_gpu.py runs it on the kernels)."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture()
def cpu_ops(sg2, oracle, monkeypatch):
    model = importlib.import_module("stylegan-for-facerec_b200.stylegan2.model")
    fused = importlib.import_module("stylegan-for-facerec_b200.stylegan2.op.fused_act")
    monkeypatch.setattr(model, "upfirdn2d", lambda x, k, up=1, down=1, pad=(0, 0): oracle.upfirdn2d(x, k, up, down, pad))
    monkeypatch.setattr(model, "fused_leaky_relu", lambda x, b, ns=0.2, s=2 ** 0.5: oracle.fused_leaky_relu(x, b, ns, s))
    monkeypatch.setattr(fused.FusedLeakyReLU, "forward",
                        lambda self, x: oracle.fused_leaky_relu(x, self.bias, self.negative_slope, self.scale))
    return model


def test_discriminator_matches_reference_golden(sg2, cpu_ops):
    import make_golden_disc as MD
    g = np.load(os.path.join(GOLDEN, "disc.npz"))
    keys = json.load(open(os.path.join(GOLDEN, "disc_state_dict_keys.json")))
    for name, size, cm, batch in MD.CASES:
        D = sg2.Discriminator(size, channel_multiplier=cm).eval()
        assert [[k, list(v.shape)] for k, v in D.state_dict().items()] == keys[name], name
        D.load_state_dict(MD.seeded_state_dict(D, name), strict=True)
        with torch.no_grad():
            y = D(MD.images(name, batch, size))
        ref = torch.from_numpy(g[name + "/out"])
        assert y.shape == ref.shape
        assert (y - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item()), (name, (y - ref).abs().max().item())


def test_minibatch_stddev_and_channel_table(sg2):
    D = importlib.import_module("stylegan-for-facerec_b200.stylegan2.discriminator")
    table = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}      # model.py:549-559, x multiplier
    for cm in (1, 2):
        for res, c in table.items():
            assert D.feature_channels(res, cm) == (c if res <= 32 else c * cm)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 6, 3, 5, generator=g)
    y = D.minibatch_stddev(x, 4, 1)
    # the reference's statement of it (model.py:655-668)
    s = x.view(4, -1, 1, 6, 3, 5)
    s = torch.sqrt(s.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdims=True).squeeze(2).repeat(4, 1, 3, 5)
    assert y.shape == (8, 7, 3, 5) and torch.equal(y[:, :6], x) and torch.allclose(y[:, 6:], s, atol=1e-6)
    assert D.minibatch_stddev(x[:2], 4, 1).shape == (2, 7, 3, 5)             # batch below the group size: one group
    with pytest.raises(RuntimeError, match="multiple"):
        D.minibatch_stddev(x[:6], 4, 1)
    with pytest.raises(ValueError):
        sg2.Discriminator(48)
