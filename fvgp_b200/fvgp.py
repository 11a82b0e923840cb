"""`fvGP`: multi-task GP, drop-in for `fvgp.fvGP` (reference: fvgp/fvgp.py:5-671).

A multi-task GP over (x, task) is a single-task GP on the index set x (+) task: the rows
are stacked TASK-MAJOR and NaN observations dropped (fvgp.py:626-660), everything else is
inherited from `GP`.  The transform is vectorised numpy (the reference's O(V*No) Python
double loop is SURVEY 8(f) #4)."""
import numpy as np

from .gp import GP


class fvGP(GP):
    def __init__(self, x_data, y_data, init_hyperparameters=None, noise_variances=None, compute_device="cpu",
                 kernel_function=None, kernel_function_grad=None, noise_function=None, noise_function_grad=None,
                 prior_mean_function=None, prior_mean_function_grad=None, gp2Scale=False, dask_client=None,
                 gp2Scale_batch_size=10000, gp2Scale_distribution="blockwise", linalg_mode=None, ram_economy=False,
                 args=None):
        assert isinstance(y_data, np.ndarray) and np.ndim(y_data) == 2, "y_data has to be a 2-d array (points x tasks)"
        assert noise_variances is None or np.shape(noise_variances) == np.shape(y_data), \
            "noise_variances must have the shape of y_data"
        assert isinstance(x_data, np.ndarray) and np.ndim(x_data) == 2, "fvgp_b200.fvGP needs Euclidean inputs"
        self.fvgp_x_data, self.fvgp_y_data, self.fvgp_noise_variances = x_data, y_data, noise_variances
        self.input_space_dim = x_data.shape[1]
        self.output_num = y_data.shape[1]
        x_index, y_flat, v_flat = self._transform_index_set(x_data, y_data, noise_variances)
        if init_hyperparameters is None and kernel_function is None:
            init_hyperparameters = np.ones(x_index.shape[1] + 1)
        super().__init__(x_index, y_flat, init_hyperparameters=init_hyperparameters, noise_variances=v_flat,
                         compute_device=compute_device, kernel_function=kernel_function,
                         kernel_function_grad=kernel_function_grad, noise_function=noise_function,
                         noise_function_grad=noise_function_grad, prior_mean_function=prior_mean_function,
                         prior_mean_function_grad=prior_mean_function_grad, gp2Scale=gp2Scale,
                         dask_client=dask_client, gp2Scale_batch_size=gp2Scale_batch_size,
                         gp2Scale_distribution=gp2Scale_distribution, linalg_mode=linalg_mode,
                         ram_economy=ram_economy, args=args)
        self.posterior.x_out = np.arange(self.output_num, dtype=np.float64)

    @staticmethod
    def _transform_index_set(x, y, noise=None):
        """fvgp.py:626-660."""
        n, tasks = y.shape
        keep = ~np.isnan(y)                                     # (n, tasks)
        pts, tks = np.nonzero(keep.T)[1], np.nonzero(keep.T)[0]  # task-major order
        x_index = np.column_stack([x[pts], tks.astype(np.float64)])
        y_flat = y[pts, tks]
        v_flat = None if noise is None else np.ascontiguousarray(noise[pts, tks])
        return np.ascontiguousarray(x_index), np.ascontiguousarray(y_flat), v_flat

    @property
    def input_set_dim(self):
        return self.input_space_dim

    def update_gp_data(self, x_new, y_new, noise_variances_new=None, append=True, gp_rank_n_update=None):
        x_index, y_flat, v_flat = self._transform_index_set(x_new, y_new, noise_variances_new)
        if append:
            self.fvgp_x_data = np.vstack([self.fvgp_x_data, x_new])
            self.fvgp_y_data = np.vstack([self.fvgp_y_data, y_new])
        else:
            self.fvgp_x_data, self.fvgp_y_data = x_new, y_new
        super().update_gp_data(x_index, y_flat, noise_variances_new=v_flat, append=append)
