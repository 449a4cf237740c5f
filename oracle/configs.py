"""Small architectures used by the parity tests (TEST INFRASTRUCTURE ONLY).

P  -- the BASELINE.json "Tiny" gate: DiT(dim=128, depth=2, head_dim=32, mixer 2x128)  
S  -- a small net that exercises what P cannot: mixer width != backbone width (LN+Linear maps),
      per-block qkv / ffn ratios, MoE in the backbone, 16 latent channels at mask 0.
"""
PARITY_CONFIGS = {
    "P": dict(ctor=dict(input_size=32, patch_size=2, in_channels=4, dim=128, depth=2, head_dim=32,
                        patch_mixer_depth=2, patch_mixer_dim=128, use_bias=False, expert_capacity=2.0),
              batch=4, mask_ratio=0.75, p_mean=-0.6, p_std=1.2),
    "S": dict(ctor=dict(input_size=16, patch_size=2, in_channels=4, dim=256, depth=4, head_dim=32,
                        multiple_of=64, qkv_multipliers=[0.5, 0.75, 1.0, 1.0], ffn_multipliers=[0.5, 1.5, 2.5, 4.0],
                        patch_mixer_depth=2, patch_mixer_dim=192, patch_mixer_qkv_ratio=1.0,
                        patch_mixer_mlp_ratio=2.0, use_bias=False, num_experts=8, expert_capacity=2.0,
                        pos_interp_scale=2.0),
              batch=3, mask_ratio=0.5, p_mean=0.0, p_std=0.6),
    "S16": dict(ctor=dict(input_size=16, patch_size=2, in_channels=16, dim=128, depth=2, head_dim=64,
                          multiple_of=128, qkv_multipliers=[1.0], ffn_multipliers=[2.0],
                          patch_mixer_depth=2, patch_mixer_dim=64, patch_mixer_mlp_ratio=4.0,
                          use_bias=False, num_experts=4, expert_capacity=1.0),
                batch=2, mask_ratio=0.0, p_mean=0.0, p_std=0.6),
}
