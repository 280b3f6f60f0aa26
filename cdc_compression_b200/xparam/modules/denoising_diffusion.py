"""``GaussianDiffusion`` of the x-parameterised variant, decoding on the CUDA engine.

Interface reference: xparam/modules/denoising_diffusion.py:12-28 (ctor), :89-108
(set_sample_schedule), :152-174 (ddim), :176-205 (p_sample / p_sample_loop), :207-231 (compress).
"""
import torch

from cdc_compression_b200._shared.diffusion_impl import DiffusionBase


class GaussianDiffusion(DiffusionBase):
    variant = "x"

    def __init__(self, denoise_fn, context_fn, ae_fn=None, num_timesteps=1000, loss_type="l1", lagrangian=1e-3,
                 pred_mode="noise", var_schedule="linear", aux_loss_weight=0, aux_loss_type="l1",
                 use_loss_weight=False, loss_weight_min=5, use_aux_loss_weight_schedule=False):
        super().__init__()
        assert pred_mode in ["noise", "x", "v"]
        self.denoise_fn = denoise_fn
        self.context_fn = context_fn
        self.ae_fn = ae_fn
        self.otherlogs = {}
        self.loss_type = loss_type
        self.lagrangian_beta = lagrangian
        self.var_schedule = var_schedule
        self.sample_steps = None
        self.aux_loss_weight = aux_loss_weight
        self.aux_loss_type = aux_loss_type
        self.use_aux_loss_weight_schedule = use_aux_loss_weight_schedule
        self.pred_mode = pred_mode
        self.use_loss_weight = use_loss_weight
        self.loss_weight_min = float(loss_weight_min)
        self._init_lpips(aux_loss_weight)
        self._init_schedule_buffers(var_schedule, num_timesteps, with_snr=True)

    def parameters(self, skip_keywords=("loss_fn_vgg", "ae_fn"), recurse=True):
        return (p for n, p in self.named_parameters(recurse=recurse) if not any(k in n for k in skip_keywords))

    # ---- schedule details of this variant ----------------------------------------------------------
    def _sample_indices(self, sample_steps, device):
        if sample_steps == 1:
            return torch.tensor([self.num_timesteps - 1], device=device).long()
        return super()._sample_indices(sample_steps, device)

    def _set_sigma(self, idx):
        self.snr = self.train_snr[idx]
        self.index = torch.arange(self.num_timesteps, device=idx.device)[idx]
        self.sigma = (self.sqrt_one_minus_alphas_cumprod_prev / self.sqrt_one_minus_alphas_cumprod
                      * torch.sqrt(1.0 - self.alphas_cumprod / self.alphas_cumprod_prev))

    def _unet_time_table(self):
        # the U-Net sees the *training* index normalised by T  (denoising_diffusion.py:154)
        return self.index.float() / self.num_timesteps

    def _dir_coef(self, eta):
        return torch.sqrt((self.one_minus_alphas_cumprod_prev - (eta * self.sigma) ** 2).clamp(min=0))

    # ---- the reference interface -----------------------------------------------------------------------
    @torch.no_grad()
    def ddim(self, x, t, context, clip_denoised, eta=0):
        return self._single_step(x, t, context, eta, self.pred_mode, "full" if clip_denoised else "none")

    def p_sample(self, x, t, context, clip_denoised, eta=0):
        return self.ddim(x=x, t=t, context=context, clip_denoised=clip_denoised, eta=eta)

    @torch.no_grad()
    def p_sample_loop(self, shape, context, clip_denoised=False, init=None, eta=0):
        return self._run_loop(shape, context, init, eta, self.pred_mode, "full" if clip_denoised else "none")

    @torch.no_grad()
    def compress(self, images, sample_steps=None, bpp_return_mean=True, init=None, eta=0):
        if self.ae_fn is not None:
            raise NotImplementedError("latent-space decoding (ae_fn) is not part of the B200 hot path; the demo "
                                      "configuration uses ae_fn=None")
        steps = self.num_timesteps if sample_steps is None else sample_steps
        split = self._encode_for_decode(images)
        if split is not None:      # context_fn.decode runs on the engine (no NCHW fp32 context round trip)
            q_latent, bpp, src = split
            self.set_sample_schedule(steps, images.device)
            decoded = self._run_loop(images.shape, None, init, eta, self.pred_mode, "full", q_latent=q_latent, ctxdec=src)
            return decoded, (bpp.mean() if bpp_return_mean else bpp)
        ctx = self.context_fn(images)
        self.set_sample_schedule(steps, ctx["output"][0].device)
        bpp = ctx["bpp"].mean() if bpp_return_mean else ctx["bpp"]
        return self.p_sample_loop(images.shape, ctx["output"], clip_denoised=True, init=init, eta=eta), bpp
