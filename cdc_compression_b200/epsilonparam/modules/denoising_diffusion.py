"""``GaussianDiffusion`` of the epsilon-parameterised variant, decoding on the CUDA engine.

Interface reference: epsilonparam/modules/denoising_diffusion.py:12-27 (ctor), :81-97
(set_sample_schedule), :137-152 (ddim), :154-192 (p_sample / p_sample_loop), :194-215 (compress).
The U-Net forward and the eps -> x0 -> x_{t-1} update run fused in libcdc_b200.so; with eta == 0
the whole loop is a CUDA-graph replay per step with no PyTorch op in between.
"""
import torch

from cdc_compression_b200._shared.diffusion_impl import DiffusionBase


class GaussianDiffusion(DiffusionBase):
    variant = "eps"

    def __init__(self, denoise_fn, context_fn, channels=3, num_timesteps=1000, loss_type="l1", clip_noise="half",
                 vbr=False, lagrangian=1e-3, pred_mode="noise", var_schedule="linear", aux_loss_weight=0,
                 aux_loss_type="l1"):
        super().__init__()
        assert pred_mode in ["noise", "image", "renoise"]
        self.channels = channels
        self.denoise_fn = denoise_fn
        self.context_fn = context_fn
        self.clip_noise = clip_noise
        self.vbr = vbr
        self.otherlogs = {}
        self.loss_type = loss_type
        self.lagrangian_beta = lagrangian
        self.var_schedule = var_schedule
        self.sample_steps = None
        self.aux_loss_weight = aux_loss_weight
        self.aux_loss_type = aux_loss_type
        self.pred_mode = pred_mode
        self._init_lpips(aux_loss_weight)
        self._init_schedule_buffers(var_schedule, num_timesteps, with_snr=False)

    def parameters(self, recurse=True):
        return (p for n, p in self.named_parameters(recurse=recurse) if "loss_fn_vgg" not in n)

    # ---- schedule details of this variant ----------------------------------------------------------
    def _set_sigma(self, idx):
        a, ap = self.alphas_cumprod, self.alphas_cumprod_prev
        self.sigma = torch.sqrt((1 - ap) / (1 - a)) * torch.sqrt(1 - a / ap)

    def _unet_time_table(self):
        # the U-Net sees t / sample_steps  (denoising_diffusion.py:138)
        i = torch.arange(self.sample_steps, device=self.alphas_cumprod.device)
        return i.float() / self.sample_steps

    def _dir_coef(self, eta):
        return torch.sqrt(self.one_minus_alphas_cumprod_prev - (eta * self.sigma) ** 2)

    def _clip_mode(self, clip_denoised):
        if clip_denoised in ("full", "half"):
            return clip_denoised
        return "none"

    # ---- the reference interface -----------------------------------------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        from cdc_compression_b200._shared.diffusion_impl import extract
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    @torch.no_grad()
    def ddim(self, x, t, context, clip_denoised, eta=0):
        # like the reference (:137-152), ddim() treats the U-Net output as the noise whatever pred_mode says
        return self._single_step(x, t, context, eta, "noise", self._clip_mode(clip_denoised))

    @torch.no_grad()
    def p_sample(self, x, t, context, clip_denoised, sample_mode="ddpm", eta=0):
        if sample_mode == "ddim":
            return self.ddim(x=x, t=t, context=context, clip_denoised=clip_denoised, eta=eta)
        if sample_mode == "ddpm":
            # the reference's ddpm branch reads posterior_mean_coef1/2, which it never defines
            raise NotImplementedError("sample_mode='ddpm' is broken upstream (undefined posterior coefficients); "
                                      "use 'ddim'")
        raise NotImplementedError

    @torch.no_grad()
    def p_sample_loop(self, shape, context, sample_mode, init=None, eta=0):
        if sample_mode != "ddim":
            return self.p_sample(None, None, context, self.clip_noise, sample_mode, eta)
        return self._run_loop(shape, context, init, eta, "noise", self._clip_mode(self.clip_noise))

    @torch.no_grad()
    def compress(self, images, sample_steps=None, bitrate_scale=None, sample_mode="ddpm", bpp_return_mean=True,
                 init=None, eta=0):
        steps = self.num_timesteps if sample_steps is None else sample_steps
        split = self._encode_for_decode(images, bitrate_scale) if sample_mode == "ddim" else None
        if split is not None:      # context_fn.decode runs on the engine (no NCHW fp32 context round trip)
            q_latent, bpp, src = split
            self.set_sample_schedule(steps, images.device)
            decoded = self._run_loop(images.shape, None, init, eta, "noise", self._clip_mode(self.clip_noise),
                                     q_latent=q_latent, ctxdec=src)
            return decoded, (bpp.mean() if bpp_return_mean else bpp)
        ctx = self.context_fn(images, bitrate_scale)
        self.set_sample_schedule(steps, ctx["output"][0].device)
        decoded = self.p_sample_loop(images.shape, ctx["output"], sample_mode, init=init, eta=eta)
        return decoded, (ctx["bpp"].mean() if bpp_return_mean else ctx["bpp"])
