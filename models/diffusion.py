"""Drop-in import path of the reference: ``from models.diffusion import BiologyAwareDiffusionModel``
(main.py:154, utils/train.py:457, utils/generate.py:270) resolves to the B200-native class."""
from osteosarcoma_diffusionmodel_b200.diffusion import (  # noqa: F401
    BiologyAwareDiffusionModel,
    ConditionalEmbedding,
    DiffusionUNet,
    TimeEmbedding,
)
