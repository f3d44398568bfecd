"""Oracle: the denoise loop around the transformer (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates FlowMatchEulerDiscreteScheduler (schedulers/scheduling_flow_match_euler_discrete.py: __init__ :91-131,
set_timesteps :249-349, step :373-470, default config, static `shift`) and the Wan VAP pipeline's loop body
(pipelines/wan/pipeline_wan_i2v_mot.py:801-877: latent_model_input = cat[latents, condition], reference stream at
timestep 1, two forwards for classifier-free guidance, scheduler.step in fp32), and CogVideoXDPMScheduler
(schedulers/scheduling_dpm_cogvideox.py: __init__ :181-232, rescale_zero_terminal_snr :96-124, set_timesteps :261-304,
get_variables / get_mult :306-328, step :330-440) with the CogVideoX VAP pipeline's loop body
(pipelines/cogvideo/pipeline_cogvideox_image2video_mot.py:964-1057: one B=2 forward per step for classifier-free guidance,
dynamic guidance scale, DPM-solver++ second-order update with its two noise draws)."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


def flow_match_schedule(num_inference_steps: int, shift: float = 1.0, num_train_timesteps: int = 1000):
    """(timesteps[n], sigmas[n+1]) of FlowMatchEulerDiscreteScheduler(shift=shift).set_timesteps(n)."""
    ts = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
    sig = torch.from_numpy(ts).to(torch.float32) / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)          # __init__ :119-123
    sigma_max, sigma_min = sig[0].item(), sig[-1].item()
    t = np.linspace(sigma_max * num_train_timesteps, sigma_min * num_train_timesteps, num_inference_steps)  # :302-304
    s = t / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)                 # :315 (the static shift is applied again)
    s = torch.from_numpy(s).to(dtype=torch.float32)
    timesteps = s * num_train_timesteps
    sigmas = torch.cat([s, torch.zeros(1)])
    return timesteps, sigmas


def flow_match_step(model_output: torch.Tensor, sample: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor) -> torch.Tensor:
    """step() :434-462 without stochastic sampling: fp32 Euler update, result cast to model_output.dtype."""
    sample = sample.to(torch.float32)
    return (sample + (sigma_next - sigma) * model_output).to(model_output.dtype)


def wan_denoise(forward: Callable[..., torch.Tensor], latents: torch.Tensor, condition: torch.Tensor, latents_ref: torch.Tensor,
                condition_ref: torch.Tensor, cond_kwargs: dict, uncond_kwargs: Optional[dict], num_steps: int, shift: float,
                guidance_scale: float, dtype=torch.bfloat16):
    """pipeline_wan_i2v_mot.py:801-877.  forward(hidden_states=, timestep=, hidden_states_mot_ref=, timestep_list_mot_ref=, **kw)
    -> noise prediction.  Returns the final latents and the list of per-step noise predictions."""
    timesteps, sigmas = flow_match_schedule(num_steps, shift)
    preds = []
    for i, t in enumerate(timesteps):
        x_in = torch.cat([latents, condition], dim=1).to(dtype)
        ts = t.expand(latents.shape[0])
        x_ref = torch.cat([latents_ref, condition_ref], dim=1).to(dtype)
        ts_ref = (timesteps[-1] * 0 + 1).unsqueeze(0).unsqueeze(0).repeat(1, 1)  # :812-813
        noise = forward(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, **cond_kwargs)
        if uncond_kwargs is not None:
            noise_u = forward(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, **uncond_kwargs)
            noise = noise_u + guidance_scale * (noise - noise_u)  # :874
        preds.append(noise)
        latents = flow_match_step(noise, latents, sigmas[i], sigmas[i + 1])
    return latents, preds


# ----------------------------------------------------------------------------------------------
# CogVideoX: CogVideoXDPMScheduler + pipeline loop
# ----------------------------------------------------------------------------------------------
def cog_dpm_tables(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.0120,
                   snr_shift_scale: float = 1.0, rescale_betas_zero_snr: bool = True):
    """(alphas_cumprod [1000] float64, timesteps [n] int64) of CogVideoXDPMScheduler(beta_schedule="scaled_linear",
    timestep_spacing="trailing", ...).set_timesteps(n) — the scheduler config convert_cogvideox_to_diffusers.py:312-326 writes
    (snr_shift_scale 1.0 for the 5B model, :259)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2   # :205
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)                                                     # :216
    if rescale_betas_zero_snr:                                                                                   # :96-124
        sq = ac.sqrt()
        s0, sT = sq[0].clone(), sq[-1].clone()
        sq = (sq - sT) * (s0 / (s0 - sT))
        ac = sq ** 2
    step_ratio = num_train_timesteps / num_inference_steps                                                       # "trailing" :294-299
    ts = np.round(np.arange(num_train_timesteps, 0, -step_ratio)).astype(np.int64) - 1
    return ac, torch.from_numpy(ts)


def cog_dpm_step(ac: torch.Tensor, num_inference_steps: int, model_output: torch.Tensor, old_pred: Optional[torch.Tensor], t: int,
                 t_back: Optional[int], sample: torch.Tensor, generator: torch.Generator, num_train_timesteps: int = 1000):
    """step() :330-440 with prediction_type "v_prediction", final_alpha_cumprod = 1 (set_alpha_to_one).  Noise is drawn on the
    CPU generator exactly like diffusers' randn_tensor does for a CPU generator (one draw, a second one on the 2nd-order branch)."""
    prev_t = t - num_train_timesteps // num_inference_steps
    a_t = ac[t]
    a_prev = ac[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    a_back = ac[t_back] if t_back is not None else None
    b_t = 1 - a_t
    pred = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output                                                   # v_prediction :405
    lamb = ((a_t / (1 - a_t)) ** 0.5).log()
    lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
    h = lamb_next - lamb
    mult1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
    mult2 = (-2 * h).expm1() * a_prev ** 0.5
    mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
    noise = torch.randn(sample.shape, generator=generator, dtype=sample.dtype).to(sample.device)
    prev_sample = mult1 * sample - mult2 * pred + mult_noise * noise
    if old_pred is None or prev_t < 0:
        return prev_sample, pred
    r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
    denoised = (1 + 1 / (2 * r)) * pred - (1 / (2 * r)) * old_pred
    noise = torch.randn(sample.shape, generator=generator, dtype=sample.dtype).to(sample.device)
    return mult1 * sample - mult2 * denoised + mult_noise * noise, pred


def cog_guidance_scale(guidance_scale: float, num_inference_steps: int, t: int, dynamic: bool) -> float:
    """use_dynamic_cfg :1036-1039."""
    import math
    if not dynamic:
        return guidance_scale
    return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)


def cog_denoise(forward: Callable[..., torch.Tensor], latents: torch.Tensor, image_latents: torch.Tensor, ref_latents: torch.Tensor,
                ref_image_latents: torch.Tensor, kwargs2: dict, num_steps: int, guidance_scale: float, dynamic_cfg: bool, noise_seed: int,
                dtype=torch.bfloat16, snr_shift_scale: float = 1.0):
    """pipeline_cogvideox_image2video_mot.py:964-1057 with classifier-free guidance.  latents / image_latents [1,F,16,h,w] (the
    transformer input is their channel concatenation, :975-976), ref_* the clean reference video (same timestep as the target,
    cogvideox_transformer_3d_mot.py:944-949); kwargs2 = the B=2 conditioning ([negative, positive] text, RoPE tables).
    Returns the final latents and the per-step guided noise predictions."""
    ac, timesteps = cog_dpm_tables(num_steps, snr_shift_scale=snr_shift_scale)
    gen = torch.Generator().manual_seed(noise_seed)
    latents = latents.to(dtype)  # the pipeline prepares the latents in the text-embedding dtype (:931-941)
    old_pred, preds = None, []
    for i, t in enumerate(timesteps.tolist()):
        x = torch.cat([torch.cat([latents] * 2), torch.cat([image_latents] * 2)], dim=2).to(dtype)
        xr = torch.cat([torch.cat([ref_latents] * 2), torch.cat([ref_image_latents] * 2)], dim=2).to(dtype)
        ts = torch.full((2,), t, dtype=torch.int64, device=latents.device)
        noise = forward(hidden_states=x, hidden_states_mot_ref=xr, timestep=ts, **kwargs2).float()
        g = cog_guidance_scale(guidance_scale, num_steps, t, dynamic_cfg)
        n_u, n_c = noise.chunk(2)
        noise = n_u + g * (n_c - n_u)
        preds.append(noise)
        latents, old_pred = cog_dpm_step(ac, num_steps, noise, old_pred, t, timesteps[i - 1].item() if i > 0 else None, latents, gen)
        latents = latents.to(dtype)
    return latents, preds
