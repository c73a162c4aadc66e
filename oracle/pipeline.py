"""ORACLE — test infrastructure only (never imported by the product path).

fp32 CPU restatement of the reference's step loops on synthetic inputs:
  * ``unet_extension_forward``  — I2VGenXLUnetExtension.forward, pipelines/pipeline_i2vgen_xl.py:109-362
  * ``init_pnp``                — composite.py:38-69
  * ``composite_loop``          — sample_with_pnp_..._attn_injection, pipeline_i2vgen_xl.py:1552-1734
  * ``invert_loop``             — invert, pipeline_i2vgen_xl.py:1914-2003
VAE / CLIP / file IO are replaced by the tensors of ``mvoc_b200.synthetic``.
The loops in the reference tree are restated line by line; the scheduler and UNet internals come
from un-vendored diffusers (see oracle/scheduler.py, oracle/unet.py) — parity for those is unpinned.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import hooks
from .scheduler import DDIMInverseScheduler, DDIMScheduler
from .unet import I2VGenXLUNet, UNetConfig, unet_body


def build_unet(kind: str = "reduced", seed: int = 0) -> I2VGenXLUNet:
    """Random-init weights of the named architecture: torch.manual_seed(seed) + PyTorch default inits,
    conv4 of every TemporalConvLayer zero."""
    torch.manual_seed(seed)
    cfg = UNetConfig.full() if kind == "full" else UNetConfig.reduced()
    return I2VGenXLUNet(cfg).eval().requires_grad_(False)


@torch.no_grad()
def unet_extension_forward(model, sample, timestep, fps, image_latents_first, image_latents,
                           image_embeddings=None, encoder_hidden_states=None, multi_frame_guidance=False):
    """pipelines/pipeline_i2vgen_xl.py:148-362."""
    batch_size, channels, num_frames, height, width = sample.shape
    if not multi_frame_guidance:
        image_embeddings = image_embeddings[:, 0:1, :].repeat(1, num_frames, 1)  # :151
    up_factor = 2 ** model.num_upsamplers
    forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])  # :162
    timesteps = timestep
    if not torch.is_tensor(timesteps):
        timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
    elif timesteps.dim() == 0:
        timesteps = timesteps[None].to(sample.device)
    timesteps = timesteps.expand(sample.shape[0])
    t_emb = model.time_embedding(model.time_proj(timesteps).to(model.dtype), None)  # :182-188
    fps = fps.expand(fps.shape[0])
    fps_emb = model.fps_embedding(model.time_proj(fps).to(model.dtype))  # :193
    emb = (t_emb + fps_emb).repeat_interleave(repeats=num_frames, dim=0)  # :196-197

    context_emb = sample.new_zeros(batch_size, 0, model.config.cross_attention_dim)
    context_emb = torch.cat([context_emb, encoder_hidden_states], dim=1)  # :204-207
    context_list = []
    for i in range(image_latents.size(2)):  # :211-240 (T identical iterations when not multi-frame)
        il = image_latents[:, :, i if multi_frame_guidance else 0, :].unsqueeze(2)
        il = il.permute(0, 2, 1, 3, 4).reshape(il.shape[0] * il.shape[2], il.shape[1], il.shape[3], il.shape[4])
        il = model.image_latents_context_embedding(il)
        _b, _c, _h, _w = il.shape
        il = il.permute(0, 2, 3, 1).reshape(_b, _h * _w, _c)
        ctx = torch.cat([context_emb, il], dim=1)
        image_emb = model.context_embedding(image_embeddings[:, i, :].unsqueeze(1))
        image_emb = image_emb.view(-1, model.config.in_channels, model.config.cross_attention_dim)
        context_list.append(torch.cat([ctx, image_emb], dim=1).unsqueeze(1))
        if not multi_frame_guidance and i == 0:
            # the remaining T-1 iterations recompute the same tensor; reuse it (values identical)
            context_list = context_list * image_latents.size(2)
            break
    ctx_all = torch.cat(context_list, dim=1)  # :255
    context_emb = ctx_all.reshape(ctx_all.shape[0] * ctx_all.shape[1], ctx_all.shape[2], ctx_all.shape[3])

    il = image_latents_first.permute(0, 2, 1, 3, 4).reshape(  # :264-279
        image_latents.shape[0] * image_latents.shape[2], image_latents.shape[1], image_latents.shape[3],
        image_latents.shape[4])
    il = model.image_latents_proj_in(il)
    il = (il[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 3, 4, 1, 2)
          .reshape(batch_size * height * width, num_frames, channels))
    il = model.image_latents_temporal_encoder(il)
    il = il.reshape(batch_size, height, width, num_frames, channels).permute(0, 4, 3, 1, 2)

    sample = torch.cat([sample, il], dim=1)  # :282-290
    sample = sample.permute(0, 2, 1, 3, 4).reshape((sample.shape[0] * num_frames, -1) + sample.shape[3:])
    sample = model.conv_in(sample)
    sample = model.transformer_in(sample, num_frames=num_frames, return_dict=False)[0]
    return unet_body(model, sample, emb, context_emb, num_frames, forward_upsample_size)[0]  # :293-357


def init_pnp(pipe, timesteps_full: torch.Tensor, wl) -> Dict[str, torch.Tensor]:
    """composite.py:38-69 — fractions -> leading slices of the FULL n_steps grid; installs the hooks."""
    k_conv = int(wl.n_steps * wl.pnp_f_t)
    k_spa = int(wl.n_steps * wl.pnp_spatial_attn_t)
    k_tmp = int(wl.n_steps * wl.pnp_temp_attn_t)
    conv_t = timesteps_full[:k_conv] if k_conv >= 0 else []
    spa_t = timesteps_full[:k_spa] if k_spa >= 0 else []
    tmp_t = timesteps_full[:k_tmp] if k_tmp >= 0 else []
    hooks.register_temp_attention_pnp(pipe, tmp_t, wl.inject_background)
    hooks.register_spatial_attention_pnp(pipe, spa_t, wl.inject_background)
    hooks.register_temp_conv_injection(pipe, conv_t)
    hooks.register_out_conv_injection(pipe, conv_t)
    hooks.register_resnet_injection(pipe, conv_t)
    return {"conv": conv_t, "spatial": spa_t, "temporal": tmp_t}


@torch.no_grad()
def composite_loop(unet, wl, inputs: dict, max_steps: Optional[int] = None, record: Optional[list] = None):
    """pipelines/pipeline_i2vgen_xl.py:1552-1734 on synthetic inputs; returns the final latents."""
    pipe = SimpleNamespace(unet=unet)
    sched = DDIMScheduler()
    sched.set_timesteps(wl.n_steps)
    timesteps_full = sched.timesteps.clone()
    init_pnp(pipe, timesteps_full, wl)
    timesteps = timesteps_full[wl.ddim_init_latents_t_idx:]  # :1554
    n_obj = wl.n_obj
    obj_offsets = list(getattr(wl, "obj_ddim_latents_idx_offset", None) or [0] * n_obj)  # template.yaml:61
    fusion_steps = tuple(wl.fusion_step)
    obj_fusion_timesteps = [[int(timesteps_full[obj_offsets[i]:][j]) for j in range(*fusion_steps)]
                            for i in range(n_obj)]  # :1560-1566
    latents = inputs["init_latents"].clone() * sched.init_noise_sigma  # :1570-1580
    masks = inputs["masks"]
    mask_f = [m for m, _ in masks]
    src = inputs["source_latents"]
    fusion_counter = 0  # never incremented in the reference (:1634, :1649)
    for i, t in enumerate(timesteps):
        if max_steps is not None and i >= max_steps:
            break
        t = int(t)
        bg_t = src[0][t]  # :1637
        if fusion_steps[0] <= i < fusion_steps[1]:  # :1639-1665
            r = wl.random_noise_ratio
            latents = r * latents + (1.0 - r) * bg_t
            objs_t = []
            for j in range(n_obj):
                obj = src[j + 1][obj_fusion_timesteps[j][fusion_counter]]
                objs_t.append(obj)
                inv_obj = obj * mask_f[j]
                background = latents * (1.0 - mask_f[j])
                if wl.obj_random_noise_fusion:
                    fusion = (latents * mask_f[j]) * r + (1 - r) * inv_obj
                else:
                    fusion = inv_obj
                latents = background + fusion
        else:  # :1666-1673
            objs_t = [src[j + 1][t] for j in range(n_obj)]
        latent_model_input = torch.cat([bg_t, *objs_t, latents, latents])  # :1676
        hooks.register_time_all(pipe, t, masks)  # :1684-1685
        noise_pred = unet_extension_forward(  # :1688-1699
            unet, latent_model_input, t, inputs["fps"], inputs["image_latents_first"], inputs["image_latents"],
            inputs["image_embeddings"], inputs["prompt_embeds"])
        neg, edit = noise_pred.chunk(n_obj + 3)[-2], noise_pred.chunk(n_obj + 3)[-1]  # :1714-1717
        noise = neg + wl.cfg * (edit - neg)
        b, c, f, hh, ww = latents.shape  # :1723-1731
        lat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, hh, ww)
        noise = noise.permute(0, 2, 1, 3, 4).reshape(b * f, c, hh, ww)
        lat = sched.step(noise, t, lat)
        latents = lat[None, :].reshape(b, f, c, hh, ww).permute(0, 2, 1, 3, 4)
        if record is not None:
            record.append(latents.clone())
    return latents


@torch.no_grad()
def invert_loop(unet, wl, inv_inputs: dict, n_steps: Optional[int] = None, max_steps: Optional[int] = None):
    """pipelines/pipeline_i2vgen_xl.py:1914-2003: returns {t: latents at level t} (the saved files)."""
    sched = DDIMInverseScheduler()
    sched.set_timesteps(n_steps or wl.inversion_steps)
    latents = inv_inputs["latents"].clone()
    saved = {}
    for i, t in enumerate(sched.timesteps):
        if max_steps is not None and i >= max_steps:
            break
        t = int(t)
        noise_pred = unet(latents, t, inv_inputs["fps"], inv_inputs["image_latents"],
                          inv_inputs["image_embeddings"], inv_inputs["prompt_embeds"], return_dict=False)[0]
        b, c, f, hh, ww = latents.shape
        lat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, hh, ww)
        noise = noise_pred.permute(0, 2, 1, 3, 4).reshape(b * f, c, hh, ww)
        lat = sched.step(noise, t, lat)
        latents = lat[None, :].reshape(b, f, c, hh, ww).permute(0, 2, 1, 3, 4)
        saved[t] = latents.clone()
    return saved
