"""Drop-in import path of the reference (``from utils.validation import BiologicalValidator``, main.py:274):
resolves to the B200-native validator for the three hot-path methods."""
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator  # noqa: F401
