"""QM9 total energy (reference ``config_energy.py``): n_dim 64, l_max 3 features with l<=2
spherical harmonics, 5 blocks, r_max 4, energy only."""
from functools import partial

from ..data import computeEdgeIndex
from ._common import ELEMENTS, skeleton
from .layer_configs import addEnergyOutput, featureModel

SHIFTS = [-620.4502, -16.4435, -620.4502, -620.4502, -620.4502, -620.4502, -1036.0271, -1489.8005, -2046.9702,
          -2717.4263]


def get_config(spec=None):
    config, data, model = skeleton(
        learning_rate=1e-2, batch_size=128, metric_key="validation_loss", max_epochs=int(1e6),
        early_stopping_patiences={"validation_loss": 20}, early_stopping_lower_bounds={"LR": 1e-6},
        loss_coeffs={"total_energy": [1e3, "MSELoss"]}, metrics_components={"total_energy": ["mae"]},
        lr_scheduler_patience=1, lr_scheduler_factor=0.8)
    model.n_dim, model.l_max, model.r_max, model.num_layers = 64, 3, 4.0, 5
    model.node_attrs, model.jit = "20x0e", True
    num_types = 10
    data.n_train, data.n_val, data.train_val_split, data.shuffle = 120000, 10831, "random", True
    data.path = "/opt/shared-data/qm9.hdf5"
    data.type_names = ELEMENTS[:num_types]
    data.key_map = {"Z": "species", "R": "pos", "U0": "total_energy"}
    data.preprocess = [partial(computeEdgeIndex, r_max=model.r_max)]
    net = featureModel(n_dim=model.n_dim, l_max=model.l_max, edge_spherical="1x0e+1x1o+1x2e",
                       node_attrs=model.node_attrs, edge_radial="8x0e", num_types=num_types,
                       num_layers=model.num_layers, r_max=model.r_max, normalize=False)
    model.update(addEnergyOutput(net, SHIFTS))
    return config
