"""Atomic dipoles (reference ``config_dipole.py``): n_dim 32, l_max 2, 5 blocks, ``1x1o`` head."""
from functools import partial

from ..data import computeEdgeIndex
from ..nn import PointwiseLinear
from ._common import ELEMENTS, skeleton
from .layer_configs import featureModel


def get_config(spec=None):
    config, data, model = skeleton(
        learning_rate=1e-2, batch_size=256, metric_key="validation_loss", max_epochs=int(1e6),
        early_stopping_patiences={"validation_loss": 20}, early_stopping_lower_bounds={"LR": 1e-6},
        loss_coeffs={"dipole": [1e3, "MSELoss"]}, metrics_components={"dipole": ["mae"]},
        lr_scheduler_patience=2, lr_scheduler_factor=0.8)
    model.n_dim, model.l_max, model.r_max, model.num_layers, model.node_attrs = 32, 2, 5.0, 5, "16x0e"
    num_types = 18
    data.n_train, data.n_val, data.train_val_split, data.shuffle = 811113, 202778, "random", True
    data.path = "multipole.hdf5"
    data.type_names = ELEMENTS[:num_types]
    data.preprocess = [partial(computeEdgeIndex, r_max=model.r_max)]
    if spec and "profiling" in spec:
        data.n_train, data.n_val = 2048, 256
    features = "+".join(f"{model.n_dim}x{l}e+{model.n_dim}x{l}o" for l in range(model.l_max + 1))
    net = featureModel(n_dim=model.n_dim, l_max=model.l_max, edge_spherical="1x0e+1x1o+1x2e",
                       node_attrs=model.node_attrs, edge_radial="8x0e", num_types=num_types,
                       num_layers=model.num_layers, r_max=model.r_max)
    net.layers = list(net.layers) + [("dipole_output", {"module": PointwiseLinear,
                                                        "irreps_in": (features, "node_features"),
                                                        "irreps_out": ("1x1o", "dipole")})]
    model.update(net)
    return config
