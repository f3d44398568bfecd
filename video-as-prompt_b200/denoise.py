"""Host-side denoise loop around the transformer (the unit BASELINE.json's metric counts: one denoise step = one
transformer forward at B=1 + the scheduler update; with classifier-free guidance two forwards per step).

Mirrors the reference pipeline's loop body (diffusers/pipelines/wan/pipeline_wan_i2v_mot.py:801-877) and the default
FlowMatchEulerDiscreteScheduler (diffusers/schedulers/scheduling_flow_match_euler_discrete.py:91-131, 249-349, 373-470);
the scheduler update is O(latent) elementwise work: torch by default, or fused with the classifier-free-guidance combine into one
kernel (`wan_denoise(fused_step=True)` -> vap_cfg_flow_match_step).  The CogVideoX loop
(diffusers/pipelines/cogvideo/pipeline_cogvideox_image2video_mot.py:964-1057) runs one B=2 forward per step for classifier-free
guidance and CogVideoXDPMScheduler (diffusers/schedulers/scheduling_dpm_cogvideox.py:181-232, 261-304, 306-440)."""
from __future__ import annotations

import math
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
                uncond_kwargs: Optional[dict], num_steps: int, shift: float = 3.0, guidance_scale: float = 5.0, dtype=torch.bfloat16,
                fused_step: bool = True, cache_context: bool = True, batch_cfg: bool = False) -> torch.Tensor:
    """Run `num_steps` denoise steps of the Wan VAP pipeline loop on `model` (ours or the reference's after install()).
    cache_context: the text / CLIP context embeddings and every block's cross-attention K / V are the same at every step and in both
    guidance passes; compute them once per loop (wan.context_cache; identical results — the cached tensors are what would be recomputed).
    batch_cfg: run the conditional and the unconditional pass of a step as ONE B = 2 forward (what the CogVideoX pipeline does,
    pipeline_cogvideox_image2video_mot.py:972-1001) instead of the Wan pipeline's two sequential B = 1 forwards (:815-861): every weight
    is read once per step and the per-stream GEMMs see twice the rows.  Same arithmetic per sample.
    fused_step (default): classifier-free guidance + scheduler update in ONE kernel (ops.cfg_flow_match_step) instead of seven torch
    elementwise launches — same rounding points, bit-exact with the torch expressions on a B200 (tests/gpu_checks.py cfg_flow_match_step,
    wan_denoise_fused); fused_step=False keeps the torch expressions."""
    dev = latents.device
    timesteps, sigmas = flow_match_schedule(num_steps, shift, device=dev)
    if fused_step:
        sig_host = flow_match_schedule(num_steps, shift, device="cpu")[1]
        # dt = sigma_next - sigma is a 0-dim fp32 TENSOR in the scheduler, and torch casts a 0-dim tensor operand to the other operand's
        # dtype (bf16) before the multiply (TensorIterator's common dtype; only Python / CPU-scalar operands keep fp32): hand the kernel that value
        dts = [float((sig_host[i + 1] - sig_host[i]).to(torch.bfloat16)) for i in range(num_steps)]
    from . import wan
    x_ref = torch.cat([latents_ref, condition_ref], dim=1).to(dtype)
    ts_ref = torch.ones((1, latents.shape[0]), dtype=torch.float32, device=dev)  # reference video is clean: timestep 1 (:812-813)
    if batch_cfg and uncond_kwargs is not None:
        # the B = 2 conditioning is concatenated ONCE, so the context cache sees the same tensors at every step
        cond_kwargs = {k: (torch.cat([v, uncond_kwargs[k]], dim=0) if torch.is_tensor(v) else v) for k, v in cond_kwargs.items()}
        x_ref, ts_ref = torch.cat([x_ref, x_ref], dim=0), torch.cat([ts_ref, ts_ref], dim=1)
    else:
        batch_cfg = False
    try:
        with wan.context_cache(cache_context):
            return _wan_denoise_loop(model, latents, condition, x_ref, ts_ref, cond_kwargs, uncond_kwargs, num_steps, timesteps, sigmas, guidance_scale, dtype,
                                     dts if fused_step else None, batch_cfg)
    finally:
        if cache_context and isinstance(model, torch.nn.Module):
            wan.clear_context_cache(model)  # the entries keep the conditioning tensors alive: drop them with the loop


def _wan_denoise_loop(model, latents, condition, x_ref, ts_ref, cond_kwargs, uncond_kwargs, num_steps, timesteps, sigmas, guidance_scale, dtype, dts,
                      batch_cfg):
    fused_step = dts is not None
    if fused_step:
        from . import ops
    nb = latents.shape[0]
    for i in range(num_steps):
        x_in = torch.cat([latents, condition], dim=1).to(dtype)
        ts = timesteps[i].expand(nb)
        if batch_cfg:  # [conditional | unconditional] samples in one forward
            both = model(hidden_states=torch.cat([x_in, x_in], dim=0), timestep=timesteps[i].expand(2 * nb), hidden_states_mot_ref=x_ref,
                         timestep_list_mot_ref=ts_ref, return_dict=False, **cond_kwargs)[0]
            noise, noise_u = both[:nb].contiguous(), both[nb:].contiguous()
        else:
            noise = model(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False, **cond_kwargs)[0]
        if uncond_kwargs is not None:
            if not batch_cfg:
                noise_u = model(hidden_states=x_in, timestep=ts, hidden_states_mot_ref=x_ref, timestep_list_mot_ref=ts_ref, return_dict=False,
                                **uncond_kwargs)[0]
            if fused_step:
                latents = ops.cfg_flow_match_step(noise, noise_u, latents.contiguous(), guidance_scale=guidance_scale, dt=dts[i])
                continue
            noise = noise_u + guidance_scale * (noise - noise_u)
        if fused_step:
            latents = ops.cfg_flow_match_step(noise, None, latents.contiguous(), guidance_scale=1.0, dt=dts[i])
            continue
        latents = flow_match_step(noise, latents, sigmas[i], sigmas[i + 1])
    return latents


# ----------------------------------------------------------------------------------------------
# CogVideoX: CogVideoXDPMScheduler + pipeline loop
# ----------------------------------------------------------------------------------------------
def cog_dpm_schedule(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.0120,
                     snr_shift_scale: float = 1.0, rescale_betas_zero_snr: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """alphas_cumprod [T] (float64) and the "trailing" timesteps [n] of CogVideoXDPMScheduler in the released 5B configuration
    (scaled-linear betas, SNR shift, zero-terminal-SNR rescale; convert_cogvideox_to_diffusers.py:312-326)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    if rescale_betas_zero_snr:
        root = ac.sqrt()
        first, last = root[0].clone(), root[-1].clone()
        ac = ((root - last) * (first / (first - last))) ** 2
    ts = np.round(np.arange(num_train_timesteps, 0, -num_train_timesteps / num_inference_steps)).astype(np.int64) - 1
    return ac, torch.from_numpy(ts)


def cog_dpm_step(ac: torch.Tensor, n_steps: int, v_pred: torch.Tensor, old_x0: Optional[torch.Tensor], t: int, t_back: Optional[int],
                 sample: torch.Tensor, generator: torch.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
    """One DPM-solver++ (SDE) update with v-prediction; the coefficients are 0-dim float64 tensors like the reference's, the noise
    comes from a CPU generator (what diffusers' randn_tensor does for one) so that a GPU run reproduces the CPU reference."""
    t_prev = t - ac.numel() // n_steps
    a, a_prev = ac[t], (ac[t_prev] if t_prev >= 0 else torch.tensor(1.0))
    x0 = (a ** 0.5) * sample - ((1 - a) ** 0.5) * v_pred
    log_snr = lambda al: ((al / (1 - al)) ** 0.5).log()  # noqa: E731
    h = log_snr(a_prev) - log_snr(a)
    c_sample = ((1 - a_prev) / (1 - a)) ** 0.5 * (-h).exp()
    c_x0 = (-2 * h).expm1() * a_prev ** 0.5
    c_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
    draw = lambda: torch.randn(sample.shape, generator=generator, dtype=sample.dtype).to(sample.device)  # noqa: E731
    first_order = c_sample * sample - c_x0 * x0 + c_noise * draw()
    if old_x0 is None or t_prev < 0:
        return first_order, x0
    r = (log_snr(a) - log_snr(ac[t_back])) / h
    x0_2nd = (1 + 1 / (2 * r)) * x0 - (1 / (2 * r)) * old_x0
    return c_sample * sample - c_x0 * x0_2nd + c_noise * draw(), x0


@torch.no_grad()
def cog_denoise(model, latents: torch.Tensor, image_latents: torch.Tensor, ref_latents: torch.Tensor, ref_image_latents: torch.Tensor, kwargs2: dict,
                num_steps: int, guidance_scale: float = 6.0, dynamic_cfg: bool = True, noise_seed: int = 0, dtype=torch.bfloat16,
                snr_shift_scale: float = 1.0) -> torch.Tensor:
    """`num_steps` denoise steps of the CogVideoX VAP pipeline loop on `model` (ours or the reference's after install()) with
    classifier-free guidance: kwargs2 holds the B=2 conditioning ([negative, positive] text embeddings of both streams, the RoPE
    tables, num_mot_ref); latents / image_latents [1, F, 16, h, w] are concatenated on channels into the transformer input."""
    ac, timesteps = cog_dpm_schedule(num_steps, snr_shift_scale=snr_shift_scale)
    gen = torch.Generator().manual_seed(noise_seed)
    latents = latents.to(dtype)
    xr = torch.cat([torch.cat([ref_latents] * 2), torch.cat([ref_image_latents] * 2)], dim=2).to(dtype)
    img2 = torch.cat([image_latents] * 2)
    old_x0 = None
    ts_list = timesteps.tolist()
    for i, t in enumerate(ts_list):
        x = torch.cat([torch.cat([latents] * 2), img2], dim=2).to(dtype)
        ts = torch.full((2,), t, dtype=torch.int64, device=latents.device)
        v = model(hidden_states=x, hidden_states_mot_ref=xr, timestep=ts, return_dict=False, **kwargs2)[0].float()
        g = guidance_scale if not dynamic_cfg else 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_steps - t) / num_steps) ** 5.0)) / 2)
        v_u, v_c = v.chunk(2)
        latents, old_x0 = cog_dpm_step(ac, num_steps, v_u + g * (v_c - v_u), old_x0, t, ts_list[i - 1] if i > 0 else None, latents, gen)
        latents = latents.to(dtype)
    return latents
