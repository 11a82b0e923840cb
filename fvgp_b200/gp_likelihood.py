"""Noise model (reference: fvgp/gp_likelihood.py:89-144).  V enters the hot path as the noise
diagonal fused into the K-fill; everything here is O(N) host bookkeeping."""
import inspect
import warnings

import numpy as np


class GPlikelihood:
    def __init__(self, data, trainer, noise_function=None, noise_function_grad=None):
        self.data, self.trainer = data, trainer
        self.noise_function = noise_function
        self.noise_function_grad = noise_function_grad
        if data.noise_variances is not None and callable(noise_function):
            raise Exception("Noise function and measurement noise provided. Decide which one to use.")
        if data.noise_variances is None and noise_function is None:
            warnings.warn("No noise function or measurement noise provided. "
                          "Noise variances will be set to (0.01 * mean(|y_data|))^2.")
        self.V = self.calculate_V(self.data.x_data, self.trainer.hyperparameters)

    @property
    def args(self):
        return self.data.args

    def calculate_V(self, x, hyperparameters):
        """gp_likelihood.py:89-110."""
        if self.noise_function is not None:
            if len(inspect.signature(self.noise_function).parameters) == 3:
                V = self.noise_function(x, hyperparameters, self.args)
            else:
                V = self.noise_function(x, hyperparameters)
            return V
        if self.data.noise_variances is not None:
            return self.data.noise_variances
        return np.full(len(x), (np.mean(np.abs(self.data.y_data)) / 100.0) ** 2)

    def calculate_V_grad(self, x, hyperparameters, direction=None):
        """gp_likelihood.py:112-144: zeros by default, user gradient, or central differences."""
        H = len(hyperparameters)
        if self.noise_function is None:
            return np.zeros((H, len(x))) if direction is None else np.zeros(len(x))
        if self.noise_function_grad is not None:
            if direction is None:
                return np.asarray(self.noise_function_grad(x, hyperparameters))
            return np.asarray(self.noise_function_grad(x, hyperparameters, direction))
        dirs = range(H) if direction is None else [direction]
        out = []
        for i in dirs:
            hp, hm = np.array(hyperparameters, dtype=float), np.array(hyperparameters, dtype=float)
            hp[i] += 1e-6
            hm[i] -= 1e-6
            out.append((np.asarray(self.calculate_V(x, hp)) - np.asarray(self.calculate_V(x, hm))) / 2e-6)
        return np.stack(out) if direction is None else out[0]

    def update_state(self):
        self.V = self.calculate_V(self.data.x_data, self.trainer.hyperparameters)
