"""Oracle: the denoise loop around the transformer (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates FlowMatchEulerDiscreteScheduler (schedulers/scheduling_flow_match_euler_discrete.py: __init__ :91-131,
set_timesteps :249-349, step :373-470, default config, static `shift`) and the Wan VAP pipeline's loop body
(pipelines/wan/pipeline_wan_i2v_mot.py:801-877: latent_model_input = cat[latents, condition], reference stream at
timestep 1, two forwards for classifier-free guidance, scheduler.step in fp32)."""
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
