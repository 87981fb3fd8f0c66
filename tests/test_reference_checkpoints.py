"""A checkpoint as the reference writes it -- e3nn 0.4.4 modules under DistributedDataParallel: ``module.`` prefix
(reference inference.py:46-53), Wigner-3j buffers of the compiled tensor products, generated sub-modules -- loads
strictly into the B200 drop-in; a truly missing or unknown parameter still fails."""
import pytest
import torch

from e3_layers import configs
from e3_layers.utils import build


@pytest.mark.parametrize("name", ["config_energy_force", "config_dipole"])
def test_reference_style_state_dict_loads_strictly(name):
    cfg = getattr(configs, name)()
    model = build(cfg.model_config)
    own = {k: v.clone() for k, v in model.state_dict().items()}
    ref = {}
    for k, v in own.items():
        ref["module." + k] = torch.randn_like(v) if v.is_floating_point() else v
    some = next(k for k in own if k.endswith("weight"))
    stem = "module." + some.rsplit(".", 1)[0]
    ref[stem + "._compiled_main_left_right._w3j_1_1_0"] = torch.zeros(3, 3, 1)       # e3nn TensorProduct constants
    ref[stem + "._compiled_main_right._w3j_2_1_1"] = torch.zeros(5, 3, 3)
    ref[stem + "._compiled_main.dummy"] = torch.zeros(1)
    res = model.load_state_dict(ref)                                                  # strict
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref["module." + k])
    bad = dict(ref)
    bad["module.not_a_layer.weight"] = torch.zeros(2)
    with pytest.raises(RuntimeError):
        model.load_state_dict(bad)
    short = dict(ref)
    short.pop("module." + some)
    with pytest.raises(RuntimeError):
        model.load_state_dict(short)
