"""`Cloud` -- host mirror of the reference's particle container (src/particle.jl:31-63).

`particles` is n_parts x (n_para+5), stored column-major (Fortran order) so that its bytes are
exactly the Julia `Matrix{Float64}` and exactly the device struct-of-arrays: columns
0..n_para-1 parameters, then loglh, logprior, old_loglh, accept, weight (particle.jl:58-63).
"""
from dataclasses import dataclass, field

import numpy as np


def ind_para_end(ncols):
    return ncols - 5


def ind_loglh(ncols):
    return ncols - 5


def ind_logprior(ncols):
    return ncols - 4


def ind_old_loglh(ncols):
    return ncols - 3


def ind_accept(ncols):
    return ncols - 2


def ind_weight(ncols):
    return ncols - 1


@dataclass
class Cloud:
    particles: np.ndarray
    tempering_schedule: np.ndarray = field(default_factory=lambda: np.zeros(1))
    ESS: np.ndarray = field(default_factory=lambda: np.zeros(1))
    stage_index: int = 1
    n_Φ: int = 0
    resamples: int = 0
    c: float = 0.0
    accept: float = 0.25
    total_sampling_time: float = 0.0

    @classmethod
    def empty(cls, n_params, n_parts):
        """Cloud(n_params, n_parts), src/particle.jl:50-53."""
        return cls(np.zeros((n_parts, n_params + 5), order="F"))

    def __len__(self):
        return self.particles.shape[0]

    @property
    def n_para(self):
        return self.particles.shape[1] - 5


def _col(c, k):
    p = c.particles if isinstance(c, Cloud) else c
    return p[:, k]


def get_vals(c, transpose=True):
    """n_para x n_parts (transpose=True, the reference default) or n_parts x n_para."""
    p = c.particles if isinstance(c, Cloud) else c
    v = p[:, :p.shape[1] - 5]
    return np.array(v.T if transpose else v)


def get_loglh(c):
    return np.array(_col(c, -5))


def get_logprior(c):
    return np.array(_col(c, -4))


def get_logpost(c):
    return get_loglh(c) + get_logprior(c)


def get_old_loglh(c):
    return np.array(_col(c, -3))


def get_accept(c):
    return np.array(_col(c, -2))


def get_weights(c):
    return np.array(_col(c, -1))


def cloud_isempty(c):
    p = c.particles if isinstance(c, Cloud) else c
    return p.shape[0] == 0


def update_loglh(c, v):
    c.particles[:, -5] = v


def update_logprior(c, v):
    c.particles[:, -4] = v


def update_old_loglh(c, v):
    c.particles[:, -3] = v


def set_weights(c, v):
    c.particles[:, -1] = v


def reset_weights(c):
    """src/particle.jl:378-383"""
    c.particles[:, -1] = 1.0


def update_draws(c, draws):
    """update_draws! (src/particle.jl:229-241): accepts n_parts x n_para or its transpose."""
    n, d = len(c), c.n_para
    draws = np.asarray(draws)
    if draws.shape == (n, d):
        c.particles[:, :d] = draws
    elif draws.shape == (d, n):
        c.particles[:, :d] = draws.T
    else:
        raise ValueError("update_draws!(c::Cloud, draws::Matrix): Draws are incorrectly sized!")
