from .config_dict import ConfigDict
from .utils import (keyMap, tp_path_exists, build, activations, activation_name, pruneArgs, insertAfter, replace,
                    _countParameters, getScaler, setSeed)
