"""Score model for protein C-alpha traces (reference ``config_diffusion_CA.py``; also registered
as ``config_diffusion_protein``, the name the reference's README uses): n_dim 64, 8 blocks with
LayerNormalization, 32 radial functions + relative sequence-position embedding, neighbour list as
the first model layer (radius 8 A on scaled coordinates OR same-chain |i-j| < 5 OR 2 % random)."""
from functools import partial

from e3b200 import ops

from ..data import computeEdgeIndex, computeEdgeVector
from ..nn import Concat, PointwiseLinear, RadialBasisEncoding, RelativePositionEncoding, symmetricCutoff
from ..utils import getScaler, insertAfter, replace
from ._common import skeleton
from .config_diffusion import time_conditioning
from .layer_configs import featureModel


# same chain and |i - j| < 5, or 2 % of all ordered pairs (reference config_diffusion_CA.py:58-64); a PairCriteria is
# evaluated inside the neighbour-list sweep and is also a callable with the reference's criteria(data, edge_index) protocol
criteria = ops.PairCriteria(segment_key="chain_id", max_separation=5, p_random=0.02)


def get_config(spec=""):
    config, data, model = skeleton(learning_rate=1e-2, batch_size=4, grad_acc=4, config_spec=spec,
                                   lr_scheduler_patience=1, lr_scheduler_factor=0.8, grad_clid_norm=1.0,
                                   diffusion_keys={"CA": 3})
    model.n_dim, model.l_max, model.r_max, model.num_layers = 64, 2, 5.0, 8
    model.edge_radial, model.node_attrs, model.jit = "32x0e", "32x0e", True
    num_types = 21
    data.n_train, data.n_val, data.std = 0.9, 0.1, 25.83
    data.scaler = getScaler([("CA", ("shift", "mean")), ("CA", ("scale", 1 / data.std))])
    data.inverse_scaler = getScaler([("CA", ("scale", data.std))])
    data.train_val_split, data.shuffle = "random", True
    data.path = [f"/mnt/vepfs/hb/protein_new/{i}" for i in range(8)]
    data.key_map = {}
    features = "+".join(f"{model.n_dim}x{l}e+{model.n_dim}x{l}o" for l in range(model.l_max + 1))
    net = featureModel(n_dim=model.n_dim, l_max=model.l_max, edge_spherical="1x0e+1x1o+1x2e",
                       node_attrs=model.node_attrs, edge_radial=model.edge_radial, num_types=num_types,
                       num_layers=model.num_layers, r_max=model.r_max, avg_num_neighbors=100, normalize=True)
    layers = replace(net.layers, "edge_vector", ("edge_vector", partial(computeEdgeVector, key="CA")))
    rel_pos = ("relative_position", {
        "module": RelativePositionEncoding, "segment": ("1x0e", "chain_id"), "id": ("1x0e", "id"),
        "irreps_out": (model.edge_radial, "rel_pos_embed"),
        "radial_encoding": {"module": RadialBasisEncoding, "r_max": 150, "cutoff": symmetricCutoff,
                            "trainable": True, "one_over_r": False}})
    layers = [rel_pos] + layers
    layers = insertAfter(layers, "radial_basis", ("concat1", {
        "module": Concat, "rel_pos": (model.edge_radial, "rel_pos_embed"),
        "edge_radial": (model.edge_radial, "edge_radial"), "irreps_out": (model.edge_radial, "edge_radial")}))
    layers = time_conditioning(layers, model.n_dim, model.node_attrs)
    for key in config.diffusion_keys:
        layers.append((f"score_{key}", {"module": PointwiseLinear, "irreps_in": (features, "node_features"),
                                        "irreps_out": ("1x1o", f"score_{key}")}))
    if "no_edge_layer" not in (spec or ""):
        layers = [("edge_index", partial(computeEdgeIndex, r_max=8.0 / data.std, key="CA", criteria=criteria))] + layers
    net.layers = layers
    model.update(net)
    return config
