"""Host-side DDIM schedule for the fused sampling loop.

Restates diffusers' DDIMScheduler as the reference's pipelines drive it (scheduler.step call sites:
models/pipeline.py:1649,2725-2730; models/pipeline_new_d4p.py:1447-1448) with the SD-1.x scheduler config
(beta_start .00085, beta_end .012, "scaled_linear", steps_offset 1, set_alpha_to_one False, clip_sample False,
timestep_spacing "leading") at eta = 0.  At eta = 0 every prediction_type makes the update an affine map

    x_{t-1} = c_out(t) * model_output + c_x(t) * x_t

so the whole schedule collapses into a [steps, 2] fp32 coefficient table that the conv_out epilogue of each stream
indexes with a device-side step counter (csrc/gemm_sm100.cu EPI_AXPBY) -- no per-step host work, CUDA-graph friendly.
`scale_model_input` is the identity and `init_noise_sigma` is 1 for this scheduler.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

PREDICTION_TYPES = ("epsilon", "sample", "v_prediction")


@dataclass
class DDIMSchedule:
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    steps_offset: int = 1
    prediction_type: str = "epsilon"
    set_alpha_to_one: bool = False
    alphas_cumprod: List[float] = field(init=False, repr=False)

    def __post_init__(self):
        if self.prediction_type not in PREDICTION_TYPES:
            raise ValueError(f"prediction_type must be one of {PREDICTION_TYPES}, got {self.prediction_type!r}")
        n = self.num_train_timesteps
        s0, s1 = math.sqrt(self.beta_start), math.sqrt(self.beta_end)
        # betas = linspace(sqrt(b0), sqrt(b1), n)^2 evaluated like torch's fp32 linspace/cumprod would be too lossy to
        # reproduce bit-for-bit on the host; keep the table in float64 and round once when it is uploaded.
        acc, out = 1.0, []
        for i in range(n):
            b = (s0 + (s1 - s0) * i / (n - 1)) ** 2
            acc *= 1.0 - b
            out.append(acc)
        self.alphas_cumprod = out

    def timesteps(self, num_inference_steps: int) -> List[int]:
        """timestep_spacing="leading": (arange(n) * (T // n))[::-1] + steps_offset."""
        if not 0 < num_inference_steps <= self.num_train_timesteps:
            raise ValueError("num_inference_steps must be in [1, num_train_timesteps]")
        ratio = self.num_train_timesteps // num_inference_steps
        return [i * ratio + self.steps_offset for i in range(num_inference_steps)][::-1]

    def coefficients(self, t: int, num_inference_steps: int) -> Tuple[float, float]:
        """(c_out, c_x) of the eta=0 update at timestep t."""
        a_t = self.alphas_cumprod[t]
        prev = t - self.num_train_timesteps // num_inference_steps
        a_p = self.alphas_cumprod[prev] if prev >= 0 else (1.0 if self.set_alpha_to_one else self.alphas_cumprod[0])
        sa, s1a = math.sqrt(a_t), math.sqrt(1.0 - a_t)
        sp, s1p = math.sqrt(a_p), math.sqrt(1.0 - a_p)
        if self.prediction_type == "epsilon":       # x0 = (x - s1a*e)/sa ; x_prev = sp*x0 + s1p*e
            return s1p - sp * s1a / sa, sp / sa
        if self.prediction_type == "sample":        # e = (x - sa*x0)/s1a ; x_prev = sp*x0 + s1p*e
            return sp - s1p * sa / s1a, s1p / s1a
        return -sp * s1a + s1p * sa, sp * sa + s1p * s1a   # v_prediction

    def table(self, num_inference_steps: int) -> Tuple[List[int], List[Tuple[float, float]]]:
        ts = self.timesteps(num_inference_steps)
        return ts, [self.coefficients(t, num_inference_steps) for t in ts]
