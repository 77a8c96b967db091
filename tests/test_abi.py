"""CPU: the C-ABI library loads and exports every symbol include/riqn_b200.h declares; the host package
fails loudly (no fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "riqn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(riqn_\w+)\s*\(", text)))


def test_header_symbols_exported():
    from rainbow_iqn_apex_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/riqn_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert lib.riqn_version() == 1


def test_no_silent_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rainbow_iqn_apex_b200 import _lib, Agent
    from helpers import make_args
    with pytest.raises(_lib.RiqnError):
        _lib.require_device()
    with pytest.raises(_lib.RiqnError):
        Agent(make_args(torch.device("cpu")), 18, None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rainbow_iqn_apex_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_layout_matches_reference_names():
    """Host logic on CPU tensors only (no kernel is launched): names, shapes, parameter order, arena views."""
    from rainbow_iqn_apex_b200.model import DQN
    from oracle import network as net
    from helpers import make_args
    d = DQN(make_args(torch.device("cpu")), 18)
    sd = d.state_dict()
    shapes = net.layer_shapes(18)
    assert set(sd) == set(shapes)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == s, k
    # parameter order == reference registration order (optimiser state indices in checkpoints, agent.py:43-47)
    names = [n for n, _ in d.named_parameters()]
    expect = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
              "iqn_fc.weight", "iqn_fc.bias"]
    for l in ("fcnoisy_h_v", "fcnoisy_h_a", "fcnoisy_z_v", "fcnoisy_z_a"):
        expect += [f"{l}.{p}" for p in ("weight_mu", "weight_sigma", "bias_mu", "bias_sigma")]
    assert names == expect
    assert sum(p.numel() for p in d.parameters()) == 6725894      # SURVEY.md section 2.2 census
    # [h_v | h_a] adjacent in the arena; grads are views of the gradient arena and survive zero_grad
    hv, ha = d.fcnoisy_h_v.weight_mu, d.fcnoisy_h_a.weight_mu
    assert ha.data_ptr() == hv.data_ptr() + 4 * hv.numel()
    d.zero_grad()
    assert all(p.grad.data_ptr() == d._flat_grad.data_ptr() + 4 * p._riqn_offset for p in d.parameters())
    # load_state_dict keeps the views bound
    p0 = d.conv1.weight.data_ptr()
    d.load_state_dict({k: torch.randn(s) for k, s in shapes.items()})
    assert d.conv1.weight.data_ptr() == p0 == d._flat.data_ptr() + 4 * d.conv1.weight._riqn_offset
