"""Pieces shared by the config files: element symbols (stands in for ``ase.atom.atomic_numbers``),
default trainer settings."""
from ..utils import ConfigDict

ELEMENTS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe").split()


def skeleton(**top):
    config, data, model = ConfigDict(), ConfigDict(), ConfigDict()
    config.data_config, config.model_config = data, model
    config.use_ema, config.ema_decay, config.ema_use_num_updates = True, 0.99, True
    config.optimizer_name, config.lr_scheduler_name = "Adam", "ReduceLROnPlateau"
    for k, v in top.items():
        config[k] = v
    return config, data, model
