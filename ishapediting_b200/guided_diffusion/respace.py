"""Timestep respacing with the reference's semantics (guided_diffusion/respace.py:6-127).

`SpacedDiffusion` keeps a subset of the base process' steps and recomputes betas from the
cumulative-alpha ratios; `_WrappedModel` feeds the UNet the ORIGINAL timestep of each kept step.
The 200-of-1000 map used by the editor is [0, 5, 10, ..., 989, 994, 999].  The map lives on the
device once (the reference re-uploads it on every call, respace.py:123).
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):          # accumulated float stride + round(): order matters for parity
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class SpacedDiffusion(GaussianDiffusion):
    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas = 1.0, []
        for i, acp in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - acp / last)
                last = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        self._map_dev = {}

    def timestep_map_tensor(self, device):
        key = str(device)
        if key not in self._map_dev:
            self._map_dev[key] = th.tensor(self.timestep_map, device=device, dtype=th.int64)
        return self._map_dev[key]

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        """Reference respace.py:97-98: cond_fn sees ORIGINAL timesteps, like the model."""
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def condition_score(self, cond_fn, *args, **kwargs):
        """Reference respace.py:100-101."""
        return super().condition_score(self._wrap_model(cond_fn), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # scaling is done by the wrapped model


class _WrappedModel:
    def __init__(self, model, diffusion, rescale_timesteps, original_num_steps):
        self.model = model
        self.diffusion = diffusion
        self.timestep_map = diffusion.timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, ts, **kwargs):
        new_ts = self.diffusion.timestep_map_tensor(ts.device)[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
