from micro_diffusion_b200.models.model import LatentDiffusion, PrecomputedLatentStubs, create_latent_diffusion  # noqa: F401
