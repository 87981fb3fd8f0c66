"""Model-config builders of the reference, restated as plain node lists with module NAMES
(oracle; test infrastructure only).  ``oracle.ref_layers.build`` resolves the names.

Follows ``e3_layers/configs/layer_configs.py:10-166`` (featureModel, embedCategorial,
addEnergyOutput, addForceOutput) and the model sections of ``config_energy_force.py:37-76``,
``config_energy.py:34-80``, ``config_dipole.py:35-79``, ``config_diffusion.py:32-119``,
``config_diffusion_CA.py:90-191``.  All sizes are parameters so that tests can run reduced
widths; defaults are the reference's.
"""
import copy
import functools

from .irreps import Irreps
from .ref_layers import tp_path_exists


def _insert_after(lst, key, item):
    for i, (k, _) in enumerate(lst):
        if k == key:
            return lst[: i + 1] + [item] + lst[i + 1 :]
    raise ValueError(key)


def feature_model(n_dim, l_max, edge_radial, num_types, num_layers, r_max, node_attrs,
                  edge_spherical=None, avg_num_neighbors=10, normalize=False, pos_key="pos"):
    """layer_configs.py:10-101"""
    node_features = "+".join(f"{n_dim}x{n}e+{n_dim}x{n}o" for n in range(l_max + 1))
    if edge_spherical is None:
        edge_spherical = "+".join(f"1x{n}e" if n % 2 == 0 else f"1x{n}o" for n in range(l_max + 1))
    layers = []
    if pos_key == "pos":
        layers.append(("edge_vector", "computeEdgeVector"))
    else:
        layers.append(("edge_vector", ("computeEdgeVector", {"key": pos_key})))
    layers.append(("onehot", {"module": "OneHotEncoding", "num_types": num_types,
                              "irreps_out": (f"{num_types}x0e", "onehot"), "irreps_in": ("1x0e", "species")}))
    layers.append(("embedding", {"module": "PointwiseLinear", "irreps_in": (f"{num_types}x0e", "onehot"),
                                 "irreps_out": (node_attrs, "node_attrs")}))
    layers.append(("node_features", {"module": "PointwiseLinear", "irreps_in": (f"{num_types}x0e", "onehot"),
                                     "irreps_out": (f"{n_dim}x0e", "node_features")}))
    layers.append(("spharm_edges", {"module": "SphericalEncoding", "irreps_out": (edge_spherical, "edge_spherical"),
                                    "irreps_in": ("1x1o", "edge_vector")}))
    layers.append(("radial_basis", {"module": "RadialBasisEncoding", "r_max": r_max, "trainable": True,
                                    "polynomial_degree": 6, "irreps_in": ("1x0e", "edge_length"),
                                    "irreps_out": (edge_radial, "edge_radial")}))
    conv = {"module": "FactorizedConvolution", "avg_num_neighbors": avg_num_neighbors, "use_sc": True,
            "invariant_layers": 3, "invariant_neurons": n_dim}
    cur = Irreps(f"{n_dim}x0e")
    full = Irreps(node_features)
    for i in range(num_layers):
        nxt = Irreps([(mul, ir) for mul, ir in full if tp_path_exists(cur, edge_spherical, ir)])
        layers.append((f"layer{i}", {
            "module": "MessagePassing", "resnet": False, "convolution": copy.deepcopy(conv),
            "nonlinearity_type": "gate",
            "nonlinearity_scalars": {"e": "silu", "o": "tanhlu"},
            "nonlinearity_gates": {"e": "silu", "o": "tanhlu"},
            "normalize": normalize,
            "node_attrs": node_attrs,
            "input_features": [str(cur), "node_features"],
            "edge_radial": edge_radial,
            "edge_spherical": edge_spherical,
            "output_features": [str(nxt), "node_features"],
        }))
        cur = nxt
    meta = {"n_dim": n_dim, "l_max": l_max, "num_types": num_types, "node_features": node_features,
            "edge_spherical": edge_spherical, "edge_radial": edge_radial, "node_attrs": node_attrs}
    return layers, meta


def add_energy_output(layers, meta, shifts=None, output_key="total_energy"):
    """layer_configs.py:121-147"""
    layers = list(layers)
    layers.append(("output_linear", {"module": "PointwiseLinear", "irreps_in": (meta["node_features"], "node_features"),
                                     "irreps_out": ("1x0e", "energy")}))
    if shifts is not None:
        layers.append(("rescale", {"module": "PerTypeScaleShift", "num_types": meta["num_types"], "shifts": shifts,
                                   "scales": None, "irreps_in": ("1x0e", "energy"), "irreps_out": ("1x0e", "energy"),
                                   "species": ("1x0e", "atom_types")}))
    layers.append(("reduce", {"module": "Pooling", "reduce": "sum", "irreps_in": ("1x0e", "energy"),
                              "irreps_out": ("1x0e", output_key)}))
    return layers


def add_force_output(layers, gradients="forces", y="energy", sign=-1.0):
    """layer_configs.py:150-166"""
    return {"module": "GradientOutput", "func": {"module": "SequentialGraphNetwork", "layers": layers},
            "x": ("1x1o", "pos"), "y": ("1x0e", y), "gradients": ("1x1o", gradients), "sign": sign}


ENERGY_FORCE_SHIFTS = [-3.7204, -2.2483, -3.7204, -3.7204, -3.7204, -3.7204, -7.6108, -4.0182,
                       -5.2651, -3.7204, -3.7204, -3.7204, -3.7204, -3.7204, -3.7204, -3.7204,
                       -3.2213, -3.7204, -3.7204, -3.7204]
ENERGY_SHIFTS = [-620.4502, -16.4435, -620.4502, -620.4502, -620.4502, -620.4502, -1036.0271,
                 -1489.8005, -2046.9702, -2717.4263]


def config_energy_force(n_dim=64, l_max=2, r_max=5.0, num_layers=5, node_attrs="16x0e", num_types=20):
    """config_energy_force.py:37-76"""
    layers, meta = feature_model(n_dim=n_dim, l_max=l_max, edge_spherical="1x0e+1x1o+1x2e",
                                 node_attrs=node_attrs, edge_radial="8x0e", num_types=num_types,
                                 num_layers=num_layers, r_max=r_max)
    shifts = (ENERGY_FORCE_SHIFTS * ((num_types + 19) // 20))[:num_types]
    layers = add_energy_output(layers, meta, shifts, output_key="energy")
    return add_force_output(layers)


def config_energy(n_dim=64, l_max=3, r_max=4.0, num_layers=5, node_attrs="20x0e", num_types=10):
    """config_energy.py:34-80"""
    layers, meta = feature_model(n_dim=n_dim, l_max=l_max, edge_spherical="1x0e+1x1o+1x2e",
                                 node_attrs=node_attrs, edge_radial="8x0e", num_types=num_types,
                                 num_layers=num_layers, r_max=r_max, normalize=False)
    layers = add_energy_output(layers, meta, (ENERGY_SHIFTS * ((num_types + 9) // 10))[:num_types])
    return {"module": "SequentialGraphNetwork", "layers": layers}


def config_dipole(n_dim=32, l_max=2, r_max=5.0, num_layers=5, node_attrs="16x0e", num_types=18):
    """config_dipole.py:35-79"""
    layers, meta = feature_model(n_dim=n_dim, l_max=l_max, edge_spherical="1x0e+1x1o+1x2e",
                                 node_attrs=node_attrs, edge_radial="8x0e", num_types=num_types,
                                 num_layers=num_layers, r_max=r_max)
    layers.append(("dipole_output", {"module": "PointwiseLinear", "irreps_in": (meta["node_features"], "node_features"),
                                     "irreps_out": ("1x1o", "dipole")}))
    return {"module": "SequentialGraphNetwork", "layers": layers}


def _time_and_attr_layers(layers, n_dim, node_attrs):
    """config_diffusion.py:84-102 == config_diffusion_CA.py:149-166"""
    layers = _insert_after(layers, "embedding", ("time_encoding", {
        "module": "RadialBasisEncoding", "r_max": 1.0, "trainable": True, "irreps_in": ("1x0e", "t"),
        "one_over_r": False, "irreps_out": (f"{n_dim}x0e", "time_encoding")}))
    layers = _insert_after(layers, "time_encoding", ("graph2node", {
        "module": "Broadcast", "irreps_in": (f"{n_dim}x0e", "time_encoding"),
        "irreps_out": (f"{n_dim}x0e", "time_encoding"), "to": "node"}))
    layers = _insert_after(layers, "graph2node", ("concat2", {
        "module": "Concat", "node_attrs": (node_attrs, "node_attrs"),
        "time_encoding": (f"{n_dim}x0e", "time_encoding"), "irreps_out": (node_attrs, "node_attrs")}))
    return layers


def config_diffusion(n_dim=32, l_max=2, num_layers=4, edge_radial="8x0e", node_attrs="16x0e",
                     r_max=8.0, std=1.4, num_types=18, nll=False):
    """config_diffusion.py:32-119"""
    layers, meta = feature_model(n_dim=n_dim, l_max=l_max, edge_spherical="1x0e+1x1o+1x2e",
                                 node_attrs=node_attrs, edge_radial=edge_radial, num_types=num_types,
                                 num_layers=num_layers, r_max=r_max / std)
    layers = _insert_after(layers, "radial_basis", ("bond_onehot", {
        "module": "OneHotEncoding", "num_types": 4, "irreps_in": ("1x0e", "bond_type"),
        "irreps_out": ("4x0e", "bond_type_onehot")}))
    layers = _insert_after(layers, "bond_onehot", ("concat1", {
        "module": "Concat", "bondtype": ("4x0e", "bond_type_onehot"), "edge_radial": (edge_radial, "edge_radial"),
        "irreps_out": (edge_radial, "edge_radial")}))
    layers = _time_and_attr_layers(layers, n_dim, node_attrs)
    if nll:
        layers = add_energy_output(layers, meta, shifts=None, output_key="nll")
        return add_force_output(layers, y="nll", gradients="score")
    layers.append(("score_output", {"module": "PointwiseLinear", "irreps_in": (meta["node_features"], "node_features"),
                                    "irreps_out": ("1x1o", "score")}))
    return {"module": "SequentialGraphNetwork", "layers": layers}


def config_diffusion_CA(n_dim=64, l_max=2, num_layers=8, edge_radial="32x0e", node_attrs="32x0e",
                        r_max=5.0, num_types=21, with_edge_index_layer=False, std=25.83, criteria=None):
    """config_diffusion_CA.py:90-191.  The neighbour-list layer (:190-191) draws unseeded
    random edges (D9); pass ``with_edge_index_layer=True`` and a deterministic ``criteria``
    to include it, otherwise supply ``edge_index`` in the input batch."""
    layers, meta = feature_model(n_dim=n_dim, l_max=l_max, edge_spherical="1x0e+1x1o+1x2e",
                                 node_attrs=node_attrs, edge_radial=edge_radial, num_types=num_types,
                                 num_layers=num_layers, r_max=r_max, avg_num_neighbors=100, normalize=True,
                                 pos_key="CA")
    rel = ("relative_position", {
        "module": "RelativePositionEncoding", "segment": ("1x0e", "chain_id"), "id": ("1x0e", "id"),
        "irreps_out": (edge_radial, "rel_pos_embed"),
        "radial_encoding": {"module": "RadialBasisEncoding", "r_max": 150, "cutoff": "symmetricCutoff",
                            "trainable": True, "one_over_r": False}})
    layers = [rel] + layers
    layers = _insert_after(layers, "radial_basis", ("concat1", {
        "module": "Concat", "rel_pos": (edge_radial, "rel_pos_embed"), "edge_radial": (edge_radial, "edge_radial"),
        "irreps_out": (edge_radial, "edge_radial")}))
    layers = _time_and_attr_layers(layers, n_dim, node_attrs)
    layers.append(("score_CA", {"module": "PointwiseLinear", "irreps_in": (meta["node_features"], "node_features"),
                                "irreps_out": ("1x1o", "score_CA")}))
    if with_edge_index_layer:
        layers = [("edge_index", ("computeEdgeIndex", {"r_max": 8.0 / std, "key": "CA", "criteria": criteria}))] + layers
    return {"module": "SequentialGraphNetwork", "layers": layers}
