"""Host side of DDPMSampler (reference sampler.mojo:5-124) and get_time_embedding
(helpers/utils.mojo:353-370): schedule scalars only - the elementwise update runs on the device
(tsd_sampler_step / the loop graph).  T defaults to 1000 training steps."""
from __future__ import annotations

import numpy as np


def get_time_embedding(timestep: float, as_written: bool = False) -> np.ndarray:
    """(320,) = [cos(t f) ; sin(t f)].  Intended f_i = 10000^(-i/160); `as_written` reproduces the
    reference's swapped base/exponent (utils.mojo:361)."""
    i = np.arange(160, dtype=np.float64)
    with np.errstate(over="ignore"):
        f = np.power(-i / 160.0, 10000.0) if as_written else np.power(10000.0, -i / 160.0)
    x = f * float(timestep)
    return np.concatenate([np.cos(x), np.sin(x)]).astype(np.float32)


class DDPMSampler:
    def __init__(self, seed_val: int = 0, num_training_steps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.0120):
        self.seed_val = seed_val
        self.num_training_steps = num_training_steps
        self.betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_training_steps) ** 2   # sampler.mojo:28-30
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = np.cumprod(self.alphas)                                            # :31-32
        self.num_inference_steps = 1
        self.start_step = 0
        self.timesteps = np.arange(num_training_steps - 1, -1, -1)

    def set_inference_timesteps(self, num_inference_steps: int = 1):                              # :35-44
        self.num_inference_steps = num_inference_steps
        ratio = self.num_training_steps // num_inference_steps
        self.timesteps = np.round(np.arange(num_inference_steps - 1, -1, -1) * ratio).astype(np.int64)

    def get_previous_timestep(self, t: int) -> int:                                              # :46-51
        return t - self.num_training_steps // self.num_inference_steps

    def set_strength(self, strength: float):                                                     # :67-73 (intended slice)
        start = self.num_inference_steps - int(self.num_inference_steps * strength)
        self.timesteps = self.timesteps[start:]
        self.start_step = start

    def coefficients(self, t: int) -> np.ndarray:
        """[sqrt(ab_t), sqrt(1-ab_t), c0, c1, sigma] of step() (:75-109; variance :53-65)."""
        prev = self.get_previous_timestep(t)
        ab = self.alphas_cumprod[t]
        ab_prev = self.alphas_cumprod[prev] if prev >= 0 else 1.0
        cur_alpha = ab / ab_prev
        cur_beta = 1.0 - cur_alpha
        c0 = (ab_prev ** 0.5 * cur_beta) / (1.0 - ab)
        c1 = cur_alpha ** 0.5 * (1.0 - ab_prev) / (1.0 - ab)
        sigma = 0.0
        if t > 0:
            sigma = max((1.0 - ab_prev) / (1.0 - ab) * cur_beta, 1e-20) ** 0.5
        return np.array([ab ** 0.5, (1.0 - ab) ** 0.5, c0, c1, sigma], np.float32)

    def coefficient_table(self) -> np.ndarray:
        return np.stack([self.coefficients(int(t)) for t in self.timesteps]).astype(np.float32)

    def add_noise_coefficients(self, t: int):                                                    # :111-124
        ab = self.alphas_cumprod[int(t)]
        return np.float32(ab ** 0.5), np.float32((1.0 - ab) ** 0.5)
