"""Model builders shared by the configs (reference ``e3_layers/configs/layer_configs.py``):
the feature model (embeddings -> SH -> radial basis -> N interaction blocks) and the energy /
force heads.  A config is a ``ConfigDict`` with ``module`` and ``layers = [(key, node | callable)]``."""
from copy import deepcopy

from e3b200.irreps import Irreps

from ..data import computeEdgeVector
from ..nn import (FactorizedConvolution, GradientOutput, MessagePassing, OneHotEncoding, PerTypeScaleShift,
                  PointwiseLinear, Pooling, RadialBasisEncoding, SequentialGraphNetwork, SphericalEncoding)
from ..utils import ConfigDict, tp_path_exists


def embedCategorial(num_types, irreps_in, irreps_out):
    one_hot = f"{num_types}x0e"
    return {
        "onehot": {"module": OneHotEncoding, "num_types": num_types, "irreps_out": (one_hot, "onehot"),
                   "irreps_in": irreps_in},
        "embedding": {"module": PointwiseLinear, "irreps_in": (one_hot, "onehot"), "irreps_out": irreps_out},
    }


def featureModel(n_dim, l_max, edge_radial, num_types, num_layers, r_max, node_attrs, edge_spherical=None,
                 avg_num_neighbors=10, normalize=False):
    node_features = "+".join(f"{n_dim}x{l}e+{n_dim}x{l}o" for l in range(l_max + 1))
    if edge_spherical is None:
        edge_spherical = "+".join(f"1x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(l_max + 1))

    config = ConfigDict()
    for name, value in (("n_dim", n_dim), ("l_max", l_max), ("edge_radial", edge_radial), ("num_types", num_types),
                        ("num_layers", num_layers), ("r_max", r_max), ("node_features", node_features),
                        ("edge_spherical", edge_spherical), ("node_attrs", node_attrs)):
        config[name] = value
    config.module = SequentialGraphNetwork

    layers = [("edge_vector", computeEdgeVector)]
    layers += list(embedCategorial(num_types, ("1x0e", "species"), (node_attrs, "node_attrs")).items())
    layers.append(("node_features", {"module": PointwiseLinear, "irreps_in": (f"{num_types}x0e", "onehot"),
                                     "irreps_out": (f"{n_dim}x0e", "node_features")}))
    layers.append(("spharm_edges", {"module": SphericalEncoding, "irreps_out": (edge_spherical, "edge_spherical"),
                                    "irreps_in": ("1x1o", "edge_vector")}))
    layers.append(("radial_basis", {"module": RadialBasisEncoding, "r_max": r_max, "trainable": True,
                                    "polynomial_degree": 6, "irreps_in": ("1x0e", "edge_length"),
                                    "irreps_out": (edge_radial, "edge_radial")}))

    block = {
        "module": MessagePassing, "resnet": False, "nonlinearity_type": "gate",
        "nonlinearity_scalars": {"e": "silu", "o": "tanhlu"}, "nonlinearity_gates": {"e": "silu", "o": "tanhlu"},
        "normalize": normalize,
        "convolution": {"module": FactorizedConvolution, "avg_num_neighbors": avg_num_neighbors, "use_sc": True,
                        "invariant_layers": 3, "invariant_neurons": n_dim},
        "node_attrs": node_attrs, "edge_radial": edge_radial, "edge_spherical": edge_spherical,
    }
    every = Irreps(node_features)
    current = Irreps(f"{n_dim}x0e")
    for i in range(num_layers):
        reachable = Irreps([b for b in every if tp_path_exists(current, edge_spherical, b.ir)])
        layer = deepcopy(block)
        layer["input_features"] = [str(current), "node_features"]
        layer["output_features"] = [str(reachable), "node_features"]
        layers.append((f"layer{i}", layer))
        current = reachable
    config.layers = layers
    return config


def addEnergyOutput(config, shifts=None, output_key="total_energy"):
    head = [("output_linear", {"module": PointwiseLinear, "irreps_in": (config.node_features, "node_features"),
                               "irreps_out": ("1x0e", "energy")})]
    if shifts is not None:
        head.append(("rescale", {"module": PerTypeScaleShift, "num_types": config.num_types, "shifts": shifts,
                                 "scales": None, "irreps_in": ("1x0e", "energy"), "irreps_out": ("1x0e", "energy"),
                                 "species": ("1x0e", "atom_types")}))
    head.append(("reduce", {"module": Pooling, "reduce": "sum", "irreps_in": ("1x0e", "energy"),
                            "irreps_out": ("1x0e", output_key)}))
    config.layers = list(config.layers) + head
    return config


def addForceOutput(config, gradients="forces", y="energy", sign=-1.0):
    inner = config.to_dict()
    wrapped = ConfigDict({k: v for k, v in inner.items() if k not in ("layers", "module")})
    wrapped.func = {"module": inner["module"], "layers": inner["layers"]}
    wrapped.update({"module": GradientOutput, "x": ("1x1o", "pos"), "y": ("1x0e", y),
                    "gradients": ("1x1o", gradients), "sign": sign})
    return wrapped
