"""``networks.acai_vanilla`` -- the import path the reference's plugin loader resolves for ``ae_class: VanillaACAI``
(networks/net_config.py:2-7, kwatsch/get_trainer.py:67-68).  Implementation: superresolution_aniso_mri_b200."""
from superresolution_aniso_mri_b200.networks.acai_vanilla import (  # noqa: F401
    BatchNormHolder, ConvHolder, VanillaACAI, build_decoder, build_encoder)


def create_decoder(args):
    """networks/acai_vanilla.py:105-109."""
    import math
    scales = int(round(math.log(args['width'] // args['latent_width'], 2)))
    return build_decoder(scales, args['depth'], args['latent'], args['colors']).to(args['device'])
