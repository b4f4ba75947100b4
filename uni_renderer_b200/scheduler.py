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


@dataclass
class UniPCSchedule:
    """Host-side UniPCMultistepScheduler for the fused loops -- the scheduler the shipped eval attaches to every stream
    (eval/test_real.py:485-493, 20 steps; step call sites models/pipeline.py:1649,2725-2730): solver_order 2, "bh2",
    predict_x0, lower_order_final, timestep_spacing "linspace", SD-1.x betas.

    Every quantity of the solver is a scalar, so one step -- convert_model_output, corrector (from the second step on),
    history shift, predictor -- is a set of linear combinations of five tensors: the current sample S, the last
    corrected sample LS, the two newest converted outputs H0 / H1 and the network output.  `table()` returns the
    timesteps and a [steps][10] coefficient table (row layout in csrc/elementwise.cu unipc_step_kernel) that a single
    fused kernel per stream and step consumes with a device-side step counter: no host work inside the loop."""
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    prediction_type: str = "epsilon"
    alphas_cumprod: List[float] = field(init=False, repr=False)

    def __post_init__(self):
        if self.prediction_type not in PREDICTION_TYPES:
            raise ValueError(f"prediction_type must be one of {PREDICTION_TYPES}, got {self.prediction_type!r}")
        n = self.num_train_timesteps
        s0, s1 = math.sqrt(self.beta_start), math.sqrt(self.beta_end)
        acc, out = 1.0, []
        for i in range(n):
            acc *= 1.0 - (s0 + (s1 - s0) * i / (n - 1)) ** 2
            out.append(acc)
        self.alphas_cumprod = out

    def timesteps(self, num_inference_steps: int) -> List[int]:
        """timestep_spacing="linspace": round(linspace(0, T-1, n+1))[::-1][:-1] (round half to even, like numpy)."""
        if not 0 < num_inference_steps < self.num_train_timesteps:
            raise ValueError("num_inference_steps must be in [1, num_train_timesteps)")
        T, n = self.num_train_timesteps, num_inference_steps
        return [int(round((T - 1) * i / n)) for i in range(n + 1)][::-1][:-1]

    def _lambdas(self, ts: List[int]):
        """(alpha, sigma, lambda = log alpha - log sigma) at each timestep of the walk + the appended final sigma."""
        ac = self.alphas_cumprod
        sig = [math.sqrt((1.0 - ac[t]) / ac[t]) for t in ts] + [math.sqrt((1.0 - ac[0]) / ac[0])]
        out = []
        for s in sig:
            a = 1.0 / math.sqrt(s * s + 1.0)
            out.append((a, s * a, math.log(a) - math.log(s * a)))
        return out

    @staticmethod
    def _phis(h: float):
        """bh2 / predict_x0 scalars for step size h: (h_phi_1, B_h, b_1, b_2)."""
        hh = -h
        h_phi_1 = math.expm1(hh)
        B_h = h_phi_1
        phi_k = h_phi_1 / hh - 1.0
        b1 = phi_k / B_h
        phi_k2 = phi_k / hh - 0.5
        b2 = phi_k2 * 2.0 / B_h
        return h_phi_1, B_h, b1, b2

    def table(self, num_inference_steps: int) -> Tuple[List[int], List[List[float]]]:
        ts = self.timesteps(num_inference_steps)
        als = self._lambdas(ts)
        n = len(ts)
        rows: List[List[float]] = []
        lower_order_nums, prev_order = 0, 1
        for i in range(n):
            a_i, s_i, lam_i = als[i]
            if self.prediction_type == "epsilon":        # x0 = (S - sigma*out) / alpha
                ca, cb = -s_i / a_i, 1.0 / a_i
            elif self.prediction_type == "sample":
                ca, cb = 1.0, 0.0
            else:                                        # v_prediction: x0 = alpha*S - sigma*out
                ca, cb = -s_i, a_i
            # ---- corrector: from (i-1) to i with the order the previous predictor used
            k_ls = k_h0 = k_h1 = k_x0 = 0.0
            use_corr = 1.0 if i > 0 else 0.0
            if i > 0:
                a_s0, s_s0, lam_s0 = als[i - 1]
                h = lam_i - lam_s0
                h_phi_1, B_h, b1, b2 = self._phis(h)
                k_ls = s_i / s_s0
                k_h0 = -a_i * h_phi_1
                if prev_order == 1:
                    rho_t = 0.5
                    k_x0 = -a_i * B_h * rho_t
                    k_h0 += a_i * B_h * rho_t
                else:
                    rk = (als[i - 2][2] - lam_s0) / h                 # history point i-2 relative to s0 = i-1
                    # solve [[1, 1], [rk, 1]] [rho_0, rho_t]^T = [b1, b2]^T
                    rho_0 = (b1 - b2) / (1.0 - rk)
                    rho_t = b1 - rho_0
                    # - alpha*B_h * ( rho_0 * (H1 - H0)/rk + rho_t * (x0 - H0) )
                    k_h1 = -a_i * B_h * rho_0 / rk
                    k_x0 = -a_i * B_h * rho_t
                    k_h0 += a_i * B_h * (rho_0 / rk + rho_t)
            # ---- predictor: from i to i+1; after the shift the newest output is x0 (-> c8) and the previous one H0 (-> c9)
            order = min(2, n - i, lower_order_nums + 1)
            a_t, s_t, lam_t = als[i + 1]
            h = lam_t - lam_i
            h_phi_1, B_h, _, _ = self._phis(h)
            p_s = s_t / s_i
            p_new = -a_t * h_phi_1
            p_old = 0.0
            if order == 2:
                rk = (als[i - 1][2] - lam_i) / h
                p_old = -a_t * B_h * 0.5 / rk
                p_new += a_t * B_h * 0.5 / rk
            rows.append([ca, cb, k_ls, k_h0, k_h1, k_x0, use_corr, p_s, p_new, p_old])
            prev_order = order
            if lower_order_nums < 2:
                lower_order_nums += 1
        return ts, rows
