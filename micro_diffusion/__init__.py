"""Drop-in namespace: keeps the reference's import paths (`micro_diffusion.models.model.create_latent_diffusion`
is the Hydra `_target_` of every reference config, configs/res_256_pretrain.yaml:10; train.py:9 imports
`micro_diffusion.models.utils.text_encoder_embedding_format`) and routes them to the B200 implementation."""
