from micro_diffusion_b200.models.utils import (DATA_TYPES, DistLoss, get_2d_sincos_pos_embed,  # noqa: F401
                                               text_encoder_embedding_format)
