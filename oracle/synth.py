"""Re-export of the synthetic-input generator (it lives in the product package because bench.py needs it too)."""
from osteosarcoma_diffusionmodel_b200.synthetic import *  # noqa: F401,F403
from osteosarcoma_diffusionmodel_b200.synthetic import CONFIG_YAML_DIMS, SMOKE_DIMS, SCENARIO_CONDITIONS  # noqa: F401
