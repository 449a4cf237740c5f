from micro_diffusion_b200.models.dit import DiT, MicroDiT_Tiny_2, MicroDiT_XL_2  # noqa: F401
