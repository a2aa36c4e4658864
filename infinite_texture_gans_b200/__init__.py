"""B200-native patch-by-patch Generator inference for Infinite_Texture_GANs (local padding).

Public surface (mirrors the reference's models/generators.py, models/layers.py and utils.py for this path):

    from infinite_texture_gans_b200 import generators, layers, utils
    netG = generators.ResidualPatchGenerator(...).to('cuda'); netG.load_state_dict(sd); netG.eval()
    img = utils.sample_from_gen_PatchByPatch_test(netG, z_dim=128, output_resolution_height=H, output_resolution_width=W)

All arithmetic runs in libitg_b200.so (csrc/, C ABI in include/itg.h); there is no CPU or PyTorch fallback.
"""
from . import bands, config, engine, generators, halo, layers, ops, packing, utils  # noqa: F401
from ._lib import ItgError, LIB_PATH  # noqa: F401
from .config import GenConfig  # noqa: F401
from .generators import ResidualPatchGenerator  # noqa: F401
from .layers import LocalPadder  # noqa: F401

__version__ = "0.1.0"
