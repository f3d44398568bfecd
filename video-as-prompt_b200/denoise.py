"""Host-side denoise loop around the transformer (the unit BASELINE.json's metric counts: one denoise step = one
transformer forward at B=1 + the scheduler update; with classifier-free guidance two forwards per step).

Mirrors the reference pipeline's loop body (diffusers/pipelines/wan/pipeline_wan_i2v_mot.py:801-877) and the default
FlowMatchEulerDiscreteScheduler (diffusers/schedulers/scheduling_flow_match_euler_discrete.py:91-131, 249-349, 373-470);
the scheduler update is O(latent) elementwise fp32 work and stays in torch."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def flow_match_schedule(num_inference_steps: int, shift: float = 1.0, num_train_timesteps: int = 1000, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """timesteps [n] and sigmas [n+1] (terminal 0) of FlowMatchEulerDiscreteScheduler(shift).set_timesteps(n)."""
    base = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1] / np.float32(num_train_timesteps)
    base = shift * base / (1 + (shift - 1) * base)
    t = np.linspace(float(base[0]) * num_train_timesteps, float(base[-1]) * num_train_timesteps, num_inference_steps)
    s = t / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    sig = torch.from_numpy(s).to(dtype=torch.float32, device=device)
    return sig * num_train_timesteps, torch.cat([sig, torch.zeros(1, device=device)])


def flow_match_step(model_output: torch.Tensor, sample: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor) -> torch.Tensor:
    return (sample.to(torch.float32) + (sigma_next - sigma) * model_output).to(model_output.dtype)


@torch.no_grad()
def wan_denoise(model, latents: torch.Tensor, condition: torch.Tensor, latents_ref: torch.Tensor, condition_ref: torch.Tensor, cond_kwargs: dict,
                uncond_kwargs: Optional[dict], num_steps: int, shift: float = 3.0, guidance_scale: float = 5.0, dtype=torch.bfloat16) -> torch.Tensor:
    """Run `num_steps` denoise steps of the Wan VAP pipeline loop on `model` (ours or the reference's after install())."""
    dev = latents.device
    timesteps, sigmas = flow_match_schedule(num_steps, shift, device=dev)
    x_ref = torch.cat([latents_ref, condition_ref], dim=1).to(dtype)
    ts_ref = torch.ones((1, latents.shape[0]), dtype=torch.float32, device=dev)  # reference video is clean: timestep 1 (:812-813)
    for i in range(num_steps):
        x_in = torch.cat([latents, condition], dim=1).to(dtype)
        ts = timesteps[i].expand(latents.shape[0])
        noise = model(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False, **cond_kwargs)[0]
        if uncond_kwargs is not None:
            noise_u = model(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False,
                            **uncond_kwargs)[0]
            noise = noise_u + guidance_scale * (noise - noise_u)
        latents = flow_match_step(noise, latents, sigmas[i], sigmas[i + 1])
    return latents
