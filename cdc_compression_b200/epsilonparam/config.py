"""Decode-relevant constants of the reference's epsilonparam/config.py (the demo script imports the
module but reads nothing from it; training settings are out of scope)."""
pred_mode = "noise"
loss_type = "l1"
iteration_step = 20000
sample_steps = 500
embed_dim = 64
dim_mults = (1, 2, 3, 4, 5, 6)
hyper_dim_mults = (4, 4, 4)
context_channels = 3
context_dim_mults = (1, 2, 3, 4)
clip_noise = "none"
vbr = False
sample_mode = "ddim"
