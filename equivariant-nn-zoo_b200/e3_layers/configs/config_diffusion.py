"""Score model for small molecules (reference ``config_diffusion.py``): n_dim 32, 4 blocks,
complete graphs with bond-type one-hot mixed into the radial embedding and a time embedding mixed
into the node attributes; head = ``1x1o`` score, or d(nll)/d(pos) with the 'nll' spec."""
from functools import partial

from ..data import computeEdgeIndex
from ..nn import Broadcast, Concat, OneHotEncoding, PointwiseLinear, RadialBasisEncoding
from ..utils import insertAfter
from ._common import ELEMENTS, skeleton
from .layer_configs import addEnergyOutput, addForceOutput, featureModel


def time_conditioning(layers, n_dim, node_attrs):
    """t -> Bessel embedding -> broadcast to nodes -> mixed into node_attrs"""
    emb = f"{n_dim}x0e"
    layers = insertAfter(layers, "embedding", ("time_encoding", {
        "module": RadialBasisEncoding, "r_max": 1.0, "trainable": True, "irreps_in": ("1x0e", "t"),
        "one_over_r": False, "irreps_out": (emb, "time_encoding")}))
    layers = insertAfter(layers, "time_encoding", ("graph2node", {
        "module": Broadcast, "irreps_in": (emb, "time_encoding"), "irreps_out": (emb, "time_encoding"), "to": "node"}))
    return insertAfter(layers, "graph2node", ("concat2", {
        "module": Concat, "node_attrs": (node_attrs, "node_attrs"), "time_encoding": (emb, "time_encoding"),
        "irreps_out": (node_attrs, "node_attrs")}))


def get_config(spec=""):
    spec = spec or ""
    config, data, model = skeleton(learning_rate=1e-2, batch_size=128, grad_clid_norm=1.0, grad_acc=1,
                                   lr_scheduler_patience=1, lr_scheduler_factor=0.8, config_spec=spec)
    model.n_dim, model.l_max, model.num_layers = 32, 2, 4
    model.edge_radial, model.node_attrs, model.r_max, model.jit = "8x0e", "16x0e", 8.0, True
    num_types = 18
    data.n_train, data.n_val, data.std = 120000, 10831, 1.4
    data.r_max = model.r_max / data.std
    data.train_val_split, data.shuffle, data.path = "random", True, "qm9_edge.hdf5"
    data.type_names = ELEMENTS[:num_types]
    data.key_map = {"Z": "species", "R": "pos", "U": "total_energy", "edge_attr": "bond_type"}
    data.preprocess = [partial(computeEdgeIndex, r_max=9999)]
    if "profiling" in spec:
        data.n_train, data.n_val = 2048, 256
    features = "+".join(f"{model.n_dim}x{l}e+{model.n_dim}x{l}o" for l in range(model.l_max + 1))
    net = featureModel(n_dim=model.n_dim, l_max=model.l_max, edge_spherical="1x0e+1x1o+1x2e",
                       node_attrs=model.node_attrs, edge_radial=model.edge_radial, num_types=num_types,
                       num_layers=model.num_layers, r_max=model.r_max / data.std)
    layers = insertAfter(net.layers, "radial_basis", ("bond_onehot", {
        "module": OneHotEncoding, "num_types": 4, "irreps_in": ("1x0e", "bond_type"),
        "irreps_out": ("4x0e", "bond_type_onehot")}))
    layers = insertAfter(layers, "bond_onehot", ("concat1", {
        "module": Concat, "bondtype": ("4x0e", "bond_type_onehot"), "edge_radial": (model.edge_radial, "edge_radial"),
        "irreps_out": (model.edge_radial, "edge_radial")}))
    net.layers = time_conditioning(layers, model.n_dim, model.node_attrs)
    if "nll" in spec:
        net = addForceOutput(addEnergyOutput(net, shifts=None, output_key="nll"), y="nll", gradients="score")
    else:
        net.layers = list(net.layers) + [("score_output", {"module": PointwiseLinear,
                                                           "irreps_in": (features, "node_features"),
                                                           "irreps_out": ("1x1o", "score")})]
    model.update(net)
    return config
