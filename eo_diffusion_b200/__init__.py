"""eo_diffusion_b200: B200-native (sm_100a) drop-in for EO_Diffusion's reverse-sampling
hot path -- `UNetModel`, `EODiffusion`, `DDIMSampler` with the reference signatures,
executed by the hand-written CUDA library libeo_b200.so (C ABI: include/eo_b200.h)."""
from .unet import UNetModel  # noqa: F401
from .diffusion import EODiffusion  # noqa: F401
from .ddim import DDIMSampler  # noqa: F401
from .checkpoint import load_checkpoint, make_checkpoint  # noqa: F401

__all__ = ["UNetModel", "EODiffusion", "DDIMSampler", "load_checkpoint", "make_checkpoint"]
