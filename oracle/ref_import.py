"""Import the unmodified reference (SonyResearch/micro_diffusion) on CPU with dependency stubs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's training modules need
timm / composer / easydict / diffusers / torchmetrics / open_clip, none of which is installed and none
of which is on the arithmetic path except timm's PatchEmbed (un-vendored, unpinned: setup.py:13), which
is restated here as Conv2d(k=s=p) + flatten(2).transpose(1,2) -- the behaviour the reference relies on
at dit.py:312-317,479,569.  Everything else is imported from /root/reference as is.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MICRODIT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "micro_diffusion", "models"))


def _stub(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Returns (dit_module, model_module, utils_module) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import transformers  # noqa: F401  (must come first: its lazy loader find_spec()s timm)
    import torch
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = _stub("timm")
        tm = _stub("timm.models")
        vt = _stub("timm.models.vision_transformer")

        class PatchEmbed(nn.Module):
            def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
                super().__init__()
                self.patch_size = (patch_size, patch_size)
                self.img_size = (img_size, img_size)
                self.grid_size = (img_size // patch_size, img_size // patch_size)
                self.num_patches = self.grid_size[0] * self.grid_size[1]
                self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

            def forward(self, x):
                return self.proj(x).flatten(2).transpose(1, 2)

        vt.PatchEmbed = PatchEmbed
        timm.models = tm
        tm.vision_transformer = vt
    if "composer" not in sys.modules:
        comp = _stub("composer")
        cm = _stub("composer.models")
        cm.ComposerModel = nn.Module
        comp.models = cm
    if "easydict" not in sys.modules:
        ed = _stub("easydict")

        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in {**(d or {}), **kw}.items():
                    self[k] = v

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        ed.EasyDict = EasyDict
    if "diffusers" not in sys.modules:
        df = _stub("diffusers")
        df.AutoencoderKL = object
    if "torchmetrics" not in sys.modules:
        tmx = _stub("torchmetrics")

        class Metric(nn.Module):
            def __init__(self, **kw):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                self.register_buffer(name, default.clone())

        tmx.Metric = Metric
    if "open_clip" not in sys.modules:
        _stub("open_clip")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # Our own drop-in shim package is also called `micro_diffusion`; make sure the reference wins here.
    for k in [k for k in sys.modules if k == "micro_diffusion" or k.startswith("micro_diffusion.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[k]
    saved = list(sys.path)
    try:
        sys.path = [REFERENCE_ROOT] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(
            os.path.join(os.path.dirname(__file__), ".."))]
        from micro_diffusion.models import dit as ref_dit
        from micro_diffusion.models import model as ref_model
        from micro_diffusion.models import utils as ref_utils
    finally:
        sys.path = saved
    # keep the reference modules reachable under private names and free the public name for the shim
    for k in [k for k in sys.modules if k == "micro_diffusion" or k.startswith("micro_diffusion.")]:
        sys.modules["_ref_" + k] = sys.modules.pop(k)
    _loaded = (ref_dit, ref_model, ref_utils)
    return _loaded


class FakeVAE:
    """Stands in for diffusers.AutoencoderKL: only `.config.scaling_factor` and `.requires_grad_` are
    touched when precomputed latents are used (model.py:92,98)."""

    class _Cfg:
        scaling_factor = 0.13025

    config = _Cfg()
    device = "cpu"

    def requires_grad_(self, flag):
        return self


class FakeTextEncoder:
    def requires_grad_(self, flag):
        return self


def build_reference_latent_diffusion(dit, p_mean=-0.6, p_std=1.2, train_mask_ratio=0.75, latent_res=32):
    """LatentDiffusion(model.py:22) around a reference DiT with fake frozen encoders."""
    _, ref_model, _ = load_reference()
    return ref_model.LatentDiffusion(
        dit=dit, vae=FakeVAE(), text_encoder=FakeTextEncoder(), tokenizer=None,
        precomputed_latents=True, dtype="bfloat16", latent_res=latent_res,
        p_mean=p_mean, p_std=p_std, train_mask_ratio=train_mask_ratio)
