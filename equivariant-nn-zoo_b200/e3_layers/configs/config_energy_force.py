"""Energy + forces (reference ``config_energy_force.py``): n_dim 64, l_max 2, 5 blocks, r_max 5,
forces = -dE/dpos through ``GradientOutput``."""
import ast
from functools import partial

from ..data import computeEdgeIndex
from ._common import ELEMENTS, skeleton
from .layer_configs import addEnergyOutput, addForceOutput, featureModel

SHIFTS = [-3.7204, -2.2483, -3.7204, -3.7204, -3.7204, -3.7204, -7.6108, -4.0182, -5.2651, -3.7204, -3.7204, -3.7204,
          -3.7204, -3.7204, -3.7204, -3.7204, -3.2213, -3.7204, -3.7204, -3.7204]


def get_config(spec=None):
    config, data, model = skeleton(
        epoch_subdivision=5, learning_rate=1e-2, batch_size=64, metric_key="training_loss", max_epochs=int(1e6),
        early_stopping_patiences={"training_loss": 20}, early_stopping_lower_bounds={"LR": 1e-6},
        loss_coeffs={"energy": [1e3, "MSELoss"], "forces": [3e4, "MSELoss"]},
        metrics_components={"energy": ["mae"], "forces": ["mae"]}, lr_scheduler_patience=1, lr_scheduler_factor=1.0)
    model.n_dim, model.l_max, model.r_max, model.num_layers = 64, 2, 5.0, 5
    model.jit, model.node_attrs = True, "16x0e"
    num_types = 20
    data.n_train, data.n_val, data.train_val_split, data.shuffle = 2560000, 171180, "random", True
    data.path = "/opt/shared-data/proteindata_cz/protein_E_and_F.hdf5"
    data.type_names = ELEMENTS[:num_types]
    data.preprocess = [partial(computeEdgeIndex, r_max=model.r_max)]
    if spec:  # a dict literal of dotted overrides, e.g. "{'model_config.n_dim': 32}" (parsed safely, D10)
        config.update_from_flattened_dict(ast.literal_eval(spec) if isinstance(spec, str) else spec)
    net = featureModel(n_dim=model.n_dim, l_max=model.l_max, edge_spherical="1x0e+1x1o+1x2e",
                       node_attrs=model.node_attrs, edge_radial="8x0e", num_types=num_types,
                       num_layers=model.num_layers, r_max=model.r_max)
    net = addForceOutput(addEnergyOutput(net, SHIFTS, output_key="energy"))
    model.update(net)
    return config
