from .config_dipole import get_config as config_dipole
from .config_energy import get_config as config_energy
from .config_energy_force import get_config as config_energy_force
from .config_diffusion import get_config as config_diffusion
from .config_diffusion_CA import get_config as config_diffusion_CA

# the name used by the reference's README / notebooks and by BASELINE.json
config_diffusion_protein = config_diffusion_CA
