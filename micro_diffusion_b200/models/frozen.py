"""Frozen text encoder / tokenizer loaders (reference utils.py:429-582).  Out of the hot-path scope
(unused when latents are precomputed, which train.py:25 asserts); kept as thin
delegations to the same third-party loaders so `create_latent_diffusion()` has the reference's behaviour
when those packages and weights are available."""
from __future__ import annotations

import torch

from .utils import DATA_TYPES, text_encoder_embedding_format


class UniversalTextEncoder(torch.nn.Module):
    def __init__(self, name: str, dtype: str, pretrained: bool = True):
        super().__init__()
        self.name = name
        if name.startswith("openclip:"):
            import open_clip
            self.clip = open_clip.create_model_and_transforms(name[len("openclip:"):])[0]
            self.cast = DATA_TYPES[dtype]
            self.encoder = None
        elif name == "DeepFloyd/t5-v1_1-xxl":
            from transformers import T5EncoderModel
            self.encoder = T5EncoderModel.from_pretrained(name, torch_dtype=DATA_TYPES[dtype])
        else:
            from transformers import CLIPTextModel
            self.encoder = CLIPTextModel.from_pretrained(name, subfolder="text_encoder", torch_dtype=DATA_TYPES[dtype])

    def encode(self, tokens, attention_mask=None):
        if self.name == "DeepFloyd/t5-v1_1-xxl":
            return self.encoder(tokens, attention_mask=attention_mask)["last_hidden_state"].unsqueeze(1), None
        if self.encoder is not None:
            return self.encoder(tokens)
        m = self.clip
        with torch.autocast(device_type="cuda", dtype=self.cast):
            x = m.token_embedding(tokens) + m.positional_embedding
            x = m.transformer(x.permute(1, 0, 2), attn_mask=m.attn_mask).permute(1, 0, 2)
            return m.ln_final(x).unsqueeze(1), None


class UniversalTokenizer:
    def __init__(self, name: str):
        self.name = name
        self.model_max_length, _ = text_encoder_embedding_format(name)
        if name.startswith("openclip:"):
            import open_clip
            self._tok = open_clip.get_tokenizer(name[len("openclip:"):])
        elif name == "DeepFloyd/t5-v1_1-xxl":
            from transformers import T5Tokenizer
            self._tok = T5Tokenizer.from_pretrained(name)
        else:
            from transformers import CLIPTokenizer
            self._tok = CLIPTokenizer.from_pretrained(name, subfolder="tokenizer")

    def tokenize(self, captions):
        if self.name.startswith("openclip:"):
            return {"input_ids": self._tok(captions, context_length=self.model_max_length)}
        out = self._tok(captions, padding="max_length", max_length=self.model_max_length, truncation=True,
                        return_attention_mask=self.name == "DeepFloyd/t5-v1_1-xxl", return_tensors="pt")
        res = {"input_ids": out["input_ids"]}
        if "attention_mask" in out:
            res["attention_mask"] = out["attention_mask"]
        return res
