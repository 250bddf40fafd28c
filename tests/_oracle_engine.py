"""TEST INFRASTRUCTURE: a CPU stand-in for `engine.VitEngine` built on the oracle (torchvision fp32), so that the
plugin's host logic -- node catalogue, residency shortcuts, batching, the reference-bound class -- can be driven without
a GPU.  It mimics the one property of the real engine that the plugin must respect: growing the workspace
(`reserve` / any call with a larger batch) RE-ALLOCATES it, i.e. everything device-resident is lost and
`workspace_generation()` changes.  Never imported by the product."""
from __future__ import annotations

import torch

from oracle import vit_oracle as O


class OracleEngine:
    def __init__(self, cfg, model, max_batch: int = 1):
        self.cfg, self.model = cfg, model
        self.cap, self.gen = max_batch, 1
        self.uploads = {"tokens": 0, "maps": 0, "images": 0}
        self._wipe()

    def _wipe(self):
        self.tokens = None
        self.images = None
        self.avg, self.cls, self.heads = {}, {}, {}

    # ---- what the plugin calls at construction
    def load_state_dict(self, sd):
        pass

    def set_deferred(self, on):
        pass

    def begin_request(self):
        pass

    def close(self):
        pass

    # ---- workspace
    def workspace_generation(self) -> int:
        return self.gen

    def reserve(self, batch: int, flags: int = 0) -> None:
        if batch > self.cap:
            self.cap = batch
            self.gen += 1
            self._wipe()      # like cudaFree + cudaMalloc: contents gone

    # ---- stages
    def stage_transform(self, imgs, resize):
        self.reserve(imgs.shape[0])
        self.images = O.preprocess(imgs, self.cfg.image_size, resize)
        return self.images.clone()

    def stage_embed(self, images):
        self.reserve(images.shape[0])
        self.uploads["images"] += 1
        self.tokens = O.embed(self.model, images)

    def stage_embed_resident(self, batch):
        assert self.images is not None, "embed_resident without resident images"
        self.tokens = O.embed(self.model, self.images[:batch])

    def _need_tokens(self, batch):
        assert self.tokens is not None and self.tokens.shape[0] == batch, "token stream is not resident"

    def stage_layer(self, i, batch, flags):
        self._need_tokens(batch)
        x, p = O.encoder_layer(self.model, i, self.tokens)
        self.tokens, self.avg[i], self.cls[i], self.heads[i] = x, p.mean(1), p[:, :, 0, :], p

    def stage_attn_block(self, i, batch, flags=3):
        self._need_tokens(batch)
        x, p = O.encoder_attn_half(self.model, i, self.tokens)
        self.tokens, self.avg[i], self.cls[i], self.heads[i] = x, p.mean(1), p[:, :, 0, :], p

    def stage_mlp_block(self, i, batch):
        self._need_tokens(batch)
        self.tokens = O.encoder_mlp_half(self.model, i, self.tokens)

    def stage_head(self, batch, shape=None):
        self._need_tokens(batch)
        out = O.head(self.model, self.tokens)
        return out.reshape(shape) if shape is not None else out

    def stage_rollout(self, batch, shape=None):
        L = self.cfg.num_layers
        for i in range(L):
            assert i in self.avg and self.avg[i].shape[0] == batch, f"map {i} is not resident"
        out = O.rollout_from_avg([self.avg[i] for i in range(L)])
        return out.reshape(shape) if shape is not None else out

    # ---- moving state across the boundary
    def set_tokens(self, t):
        self.reserve(t.shape[0])
        self.uploads["tokens"] += 1
        self.tokens = t.clone()

    def get_tokens(self, batch, shape=None):
        self._need_tokens(batch)
        return self.tokens.clone().reshape(shape) if shape is not None else self.tokens.clone()

    def set_avg_map(self, i, amap):
        self.reserve(amap.shape[0])
        self.uploads["maps"] += 1
        self.avg[i] = amap.clone()

    def get_avg_map(self, i, batch, shape=None):
        return self.avg[i].clone().reshape(shape)

    def get_cls_map(self, i, batch, shape=None):
        return self.cls[i].clone().reshape(shape)

    def get_cls_grid(self, i, batch, shape=None):
        return self.cls[i][:, :, 1:].clone().reshape(shape)

    def get_head_map(self, i, batch, shape=None):
        return self.heads[i].clone().reshape(shape)
