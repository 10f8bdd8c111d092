"""Model / diffusion factories with the reference's keyword surface
(guided_diffusion/script_util.py:42-187,389-455).  Classifier and super-resolution factories are out
of scope (never used by the editor)."""
import argparse

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps
from .unet import UNetModel

NUM_CLASSES = 1000


def diffusion_defaults():
    return dict(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing="",
                use_kl=False, predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False)


def model_and_diffusion_defaults():
    res = dict(image_size=128, num_channels=256, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
               num_head_channels=-1, attention_resolutions="16,8", channel_mult="", dropout=0.0, class_cond=False,
               use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=False, use_fp16=False,
               use_new_attention_order=False, in_out_channels=3)
    res.update(diffusion_defaults())
    return res


def create_model_and_diffusion(image_size, class_cond, learn_sigma, num_channels, num_res_blocks, channel_mult,
                               num_heads, num_head_channels, num_heads_upsample, attention_resolutions, dropout,
                               diffusion_steps, noise_schedule, timestep_respacing, use_kl, predict_xstart,
                               rescale_timesteps, rescale_learned_sigmas, use_checkpoint, use_scale_shift_norm,
                               resblock_updown, use_fp16, use_new_attention_order, in_out_channels):
    model = create_model(image_size, num_channels, num_res_blocks, channel_mult=channel_mult, learn_sigma=learn_sigma,
                         class_cond=class_cond, use_checkpoint=use_checkpoint,
                         attention_resolutions=attention_resolutions, num_heads=num_heads,
                         num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                         use_scale_shift_norm=use_scale_shift_norm, dropout=dropout, resblock_updown=resblock_updown,
                         use_fp16=use_fp16, use_new_attention_order=use_new_attention_order,
                         in_out_channels=in_out_channels)
    diffusion = create_gaussian_diffusion(steps=diffusion_steps, learn_sigma=learn_sigma, noise_schedule=noise_schedule,
                                          use_kl=use_kl, predict_xstart=predict_xstart,
                                          rescale_timesteps=rescale_timesteps,
                                          rescale_learned_sigmas=rescale_learned_sigmas,
                                          timestep_respacing=timestep_respacing)
    return model, diffusion


_CHANNEL_MULT = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}


def create_model(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False, class_cond=False,
                 use_checkpoint=False, attention_resolutions="16", num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0, resblock_updown=False, use_fp16=False,
                 use_new_attention_order=False, in_out_channels=3):
    if channel_mult == "":
        if image_size not in _CHANNEL_MULT:
            raise ValueError(f"unsupported image size: {image_size}")
        channel_mult = _CHANNEL_MULT[image_size]
    else:
        channel_mult = tuple(int(m) for m in channel_mult.split(","))
    attention_ds = tuple(image_size // int(res) for res in attention_resolutions.split(","))
    return UNetModel(image_size=image_size, in_channels=in_out_channels, model_channels=num_channels,
                     out_channels=(in_out_channels if not learn_sigma else in_out_channels * 2),
                     num_res_blocks=num_res_blocks, attention_resolutions=attention_ds, dropout=dropout,
                     channel_mult=channel_mult, num_classes=(NUM_CLASSES if class_cond else None),
                     use_checkpoint=use_checkpoint, use_fp16=use_fp16, num_heads=num_heads,
                     num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                     use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                     use_new_attention_order=use_new_attention_order)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing=""):
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if not timestep_respacing:
        timestep_respacing = [steps]
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                           model_mean_type=(gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON),
                           model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps)


def add_dict_to_argparser(parser, default_dict):
    for k, v in default_dict.items():
        v_type = type(v)
        if v is None:
            v_type = str
        elif isinstance(v, bool):
            v_type = str2bool
        parser.add_argument(f"--{k}", default=v, type=v_type)


def args_to_dict(args, keys):
    return {k: getattr(args, k) for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")
