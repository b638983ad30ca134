"""The reference's declarative Monte-Carlo runner (experiments.py:90-458) over the GPU.

``Experiment`` keeps the reference's interface -- ``add_axis``, ``set_static``, ``add_stochastic_variable``,
``set_config_factory``, ``add_analysis``, ``get_params_for_point``, ``run``, ``save_results`` / ``load_results`` -- and
the structure of the result dictionary: ``results[analysis][column]['all_trials' | 'mean' | 'std' | 'min' | 'max' |
'worst']`` over the axes grid.  What changes is the execution.  The reference maps one process-pool job per trial:
simulate ('asd' mode), fit, ``results_df.mean()`` (experiments.py:15-88, 381).  Here the job list is built in the
same order (same trial numbering, same order of calls to the stochastic generators), every trial of a wave is
simulated by one launch of the device generator, fitted by one batched launch per analysis, and reduced to the grid
statistics by one kernel.  ``n_cores`` is accepted and ignored.  Analyses: 'nls' and 'ekf' (the W-DFMI fitters are
outside this package).
"""
from __future__ import annotations

import copy
import itertools
import logging
import pickle
from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np

from . import _lib
from . import fit as fit_tunables
from .factories import ExperimentFactory
from .simulation import WaveformTables, pack_asd_trial, simulate_asd_batch

RESULT_KEYS = ["amp", "m", "phi", "psi", "dc", "ssq", "fitok", "tau"]  # results_df columns + tau (core.py:506-507)


class Experiment:
    def __init__(self, description: str = "Unnamed Experiment", filename: Optional[str] = None):
        self.description = description
        self.axes: Dict[str, np.ndarray] = {}
        self.static_params: Dict[str, Any] = {}
        self.stochastic_vars: Dict[str, Dict[str, Any]] = {}
        self.config_factory = None
        self._expected_params_keys = set()
        self.analyses: List[Dict[str, Any]] = []
        self.n_trials: int = 1
        self.n_fit_buffers_per_trial: int = 10
        self.f_samp: int = 200000
        self.results: Optional[Dict[str, Any]] = None
        if filename is not None:
            self.load_results(filename)

    # ---- configuration (experiments.py:127-180) -----------------------------------------------------------
    def _validate_param_name(self, name: str):
        if not self._expected_params_keys:
            logging.warning("No config factory set yet. Parameter validation will be skipped until set_config_factory() is called.")
            return
        if name not in self._expected_params_keys:
            raise ValueError(
                f"Parameter '{name}' is not recognized by the current ExperimentFactory "
                f"({type(self.config_factory).__name__}).\nExpected parameters are: "
                f"{sorted(list(self._expected_params_keys))}.\nPlease update your ExperimentFactory to handle this "
                f"parameter or remove it from your experiment configuration.")

    def add_axis(self, name: str, values):
        self._validate_param_name(name)
        self.axes[name] = np.asarray(values)

    def set_static(self, params: Dict[str, Any]):
        for name in params.keys():
            self._validate_param_name(name)
        self.static_params.update(params)

    def add_stochastic_variable(self, name: str, generator_func: Callable, depends_on: Optional[str] = None):
        self._validate_param_name(name)
        if depends_on is not None:
            self._validate_param_name(depends_on)
        self.stochastic_vars[name] = {"generator": generator_func, "depends_on": depends_on}

    def set_config_factory(self, factory):
        ok = isinstance(factory, ExperimentFactory) or (callable(factory) and hasattr(factory, "_get_expected_params_keys"))
        if not ok:
            raise TypeError("factory must be an instance of a class that inherits from ExperimentFactory.")
        self.config_factory = factory
        self._expected_params_keys = set(factory._get_expected_params_keys())

    def add_analysis(self, name: str, fitter_method: str, result_cols: Optional[List[str]] = None,
                     fitter_kwargs: Optional[Dict[str, Any]] = None):
        self.analyses.append({"name": name, "fitter_method": fitter_method, "result_cols": result_cols,
                              "fitter_kwargs": fitter_kwargs or {}})

    def get_params_for_point(self, axis_idx: Union[int, tuple]) -> Dict[str, Any]:
        """Representative parameters of one grid point (experiments.py:182-270): static + axis values + one
        deterministic draw (seed 0) of every stochastic variable; the global numpy random state is preserved."""
        params = copy.deepcopy(self.static_params)
        axis_names = list(self.axes.keys())
        if isinstance(axis_idx, (int, np.integer)):
            axis_idx = (int(axis_idx),)
        if len(axis_idx) != len(axis_names):
            raise ValueError(f"Dimension of axis_idx ({len(axis_idx)}) does not match the number of defined axes "
                             f"({len(axis_names)}).")
        for i, axis_name in enumerate(axis_names):
            params[axis_name] = self.axes[axis_name][axis_idx[i]]
        state = np.random.get_state()
        np.random.seed(0)
        try:
            for var_name, var_info in self.stochastic_vars.items():
                dep = var_info.get("depends_on")
                if dep:
                    if dep not in params:
                        raise ValueError(f"Stochastic variable '{var_name}' depends on '{dep}', which is not a defined "
                                         f"axis or static parameter.")
                    params[var_name] = var_info["generator"](params[dep])
                else:
                    params[var_name] = var_info["generator"]()
        finally:
            np.random.set_state(state)
        return {k: v for k, v in params.items() if k in self._expected_params_keys or k.startswith("_exp_")}

    def save_results(self, filename: str):
        if self.results is None:
            raise RuntimeError("No results to save. Run the experiment first.")
        with open(filename, "wb") as f:
            pickle.dump(self.results, f)

    def load_results(self, filename: str):
        with open(filename, "rb") as f:
            self.results = pickle.load(f)

    # ---- execution --------------------------------------------------------------------------------------------
    def _jobs(self, first_trial_only=False):
        """The reference's job list (experiments.py:318-372): grid points in itertools.product order, n_trials jobs
        each, trial numbers counting up across the whole experiment, stochastic generators called per job in order.
        first_trial_only: trial 0 of every point (its counter still counts all trials)."""
        axis_names = list(self.axes.keys())
        combos = list(itertools.product(*[range(len(ax)) for ax in self.axes.values()]))
        counter = 0
        for point in combos:
            point_params = copy.deepcopy(self.static_params)
            for i, name in enumerate(axis_names):
                point_params[name] = self.axes[name][point[i]]
            for j in range(1 if first_trial_only else self.n_trials):
                tp = copy.deepcopy(point_params) if self.stochastic_vars else dict(point_params)
                tp["_exp_point_idx"] = point
                tp["_exp_trial_idx"] = j
                for var_name, var_info in self.stochastic_vars.items():
                    dep = var_info.get("depends_on")
                    tp[var_name] = var_info["generator"](tp[dep]) if dep else var_info["generator"]()
                yield point, j, counter, {k: v for k, v in tp.items()
                                          if k in self._expected_params_keys or k.startswith("_exp_")}
                counter += self.n_trials if first_trial_only else 1

    def _assemble_results(self, parts, grid_shape, ntr):
        """The result dictionary from the per-rank tables ``{analysis: (stats[points, cols, 6], values[points, trials,
        cols])}``, ranks in order, each holding a contiguous range of grid points (one part on one GPU)."""
        npoints_all = int(np.prod(grid_shape)) if grid_shape else 1
        results: Dict[str, Any] = {"axes": self.axes}
        for a in self.analyses:
            st = np.concatenate([p[a["name"]][0] for p in parts], axis=0)
            allv = np.concatenate([p[a["name"]][1] for p in parts], axis=0).reshape(grid_shape + (ntr, len(RESULT_KEYS)))
            if st.shape[0] != npoints_all:
                raise RuntimeError(f"gathered {st.shape[0]} grid points, expected {npoints_all}")
            cols = a.get("result_cols") or sorted(RESULT_KEYS)
            out = {}
            for col in cols:
                if col not in RESULT_KEYS:
                    grid = np.full(grid_shape + (ntr,), np.nan)
                    out[col] = {"all_trials": grid, **{k: np.full(grid_shape, np.nan) for k in ("mean", "std", "min", "max", "worst")}}
                    continue
                c = RESULT_KEYS.index(col)
                d = {"all_trials": np.ascontiguousarray(allv[..., c])}
                for k, name in enumerate(("mean", "std", "min", "max", "worst")):
                    d[name] = st[:, c, k].reshape(grid_shape) if grid_shape else st[0, c, k]
                out[col] = d
            results[a["name"]] = out
        return results

    def run(self, n_cores: Optional[int] = None, filename: Optional[str] = None, device: Optional[int] = None,
            max_resident_bytes: int = 8 << 30, group=None) -> Dict[str, Any]:
        import torch
        # group: a torch.distributed process group (one process per GPU).  Grid points are independent, so they are cut
        # into contiguous ranges, one per rank; nothing crosses the GPUs but the finished per-point tables, gathered as
        # Python objects at the end (every rank returns the full result).  Trial numbers -- the noise keys -- are those
        # of the one-GPU run, so the result does not depend on the number of ranks.
        rank, world = 0, 1
        if group is not None:
            rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device() if world > 1 else 0
        if self.config_factory is None:
            raise ValueError("A configuration factory must be set using set_config_factory().")
        if not self.axes and not self.n_trials > 0:
            raise ValueError("At least one parameter axis must be defined using add_axis(), or n_trials must be > 0.")
        for a in self.analyses:
            if a["fitter_method"] not in ("nls", "ekf"):
                raise NotImplementedError(f"analysis '{a['name']}': fitter '{a['fitter_method']}' is not a GPU fitter "
                                          "(nls, ekf)")
        axis_names = list(self.axes.keys())
        grid_shape = tuple(len(ax) for ax in self.axes.values())
        npoints = int(np.prod(grid_shape)) if grid_shape else 1
        ntr = int(self.n_trials)
        f_samp = float(self.f_samp)

        # ---- 1. every trial's physics, packed for the device generator -------------------------------------------
        records = np.zeros((npoints * ntr, _lib.ASD_TRIAL_DOUBLES))
        df_of = np.zeros(npoints * ntr)
        tables = None
        f_mod = None
        n_samples = None

        def pack(params, counter):
            nonlocal tables, f_mod, n_samples
            cfg = self.config_factory(params)
            # a 'witness_ifo_config' (the W-DFMI factory) is simulated by the reference beside the main channel and read by
            # its W-DFMI fitters only; the analyses run here ('nls', 'ekf': checked above) never look at it
            laser, ifo = cfg["laser_config"], cfg["main_ifo_config"]
            if f_mod is None:
                f_mod = laser.f_mod
                R = int(f_samp / f_mod)  # experiments.py:28
                n_samples = self.n_fit_buffers_per_trial * R or 1
                n_samples = int((n_samples / f_samp) * f_samp)  # int(n_seconds * f_samp), physics.py:425
                tables = WaveformTables(n_samples, f_samp)
            elif laser.f_mod != f_mod:
                raise NotImplementedError("all trials of an experiment must share the modulation frequency")
            return pack_asd_trial(laser, ifo, f_samp, counter, tables, dynamic=True), laser.df

        if not self.stochastic_vars:
            # every trial of a grid point has the point's physics; only the noise key (the trial number, counted across
            # the whole experiment as in experiments.py:318-372) differs: one factory call per point, the rest is array work
            for pi, (point, j, counter, params) in enumerate(self._jobs(first_trial_only=True)):
                rec, df = pack(params, counter)
                lo = pi * ntr
                records[lo:lo + ntr] = rec
                records[lo:lo + ntr, 14] = np.arange(lo, lo + ntr, dtype=np.float64)
                df_of[lo:lo + ntr] = df
        else:
            for point, j, counter, params in self._jobs():
                records[counter], df_of[counter] = pack(params, counter)

        # ---- 2. simulate and fit in waves ----------------------------------------------------------------------------
        from .sharding import slab_bounds
        p_lo, p_hi = slab_bounds(npoints, world, rank) if world > 1 else (0, npoints)
        records, df_of = records[p_lo * ntr:p_hi * ntr], df_of[p_lo * ntr:p_hi * ntr]
        npoints_all, npoints = npoints, p_hi - p_lo
        dev = torch.device("cuda", device)
        ctx = _lib.get_context(device)
        J = records.shape[0]
        per_wave = max(1, min(J, int(max_resident_bytes // (2 * 8 * max(n_samples, 1)))))
        n_cyc = self.n_fit_buffers_per_trial  # fitter_args['n'] = num_fit_buffers (experiments.py:73)
        R_fit = int(f_samp / f_mod * n_cyc)
        nbuf = int(n_samples / R_fit) if R_fit > 0 else 0
        w0 = 2.0 * np.pi * f_mod / f_samp
        values = {a["name"]: torch.full((J, len(RESULT_KEYS)), float("nan"), dtype=torch.float64, device=dev)
                  for a in self.analyses}
        with torch.cuda.device(dev):
            per_wave = max(per_wave, 1)
            y = torch.empty((per_wave, n_samples), dtype=torch.float64, device=dev)
            rows = torch.empty((per_wave, max(nbuf, 1), _lib.ROW_STRIDE), dtype=torch.float64, device=dev)
            df_dev = torch.from_numpy(df_of).to(dev)
            for lo in range(0, J, per_wave):
                hi = min(J, lo + per_wave)
                nw = hi - lo
                simulate_asd_batch(records[lo:hi], n_samples, f_samp, tables, device=device, out=y[:nw])
                for a in self.analyses:
                    if nbuf == 0:
                        continue  # "Check buffer size": the reference's fit returns None and the trial stays NaN
                    kw = a["fitter_kwargs"]
                    ctx.use_torch_stream()
                    try:
                        if a["fitter_method"] == "nls":
                            init = [float(kw.get("init_a", 1.6)), float(kw.get("init_m", 6.0)), 0.0, float(kw.get("init_psi", 0.0))]
                            # parallel=False inside a trial (experiments.py:69): buffer 0 cold, then one chain
                            ctx.nls_fit_batch_dev(y.data_ptr(), nw, nbuf, n_samples, R_fit, int(kw.get("ndata", 10)), w0, init,
                                                  None, 0, 1 if nbuf > 1 else False,
                                                  fit_tunables.current_lm_opts(kw.get("tunables_from")), rows.data_ptr())
                        else:
                            opts = _ekf_opts(kw)
                            ctx.ekf_dev(y.data_ptr(), n_samples, nw, 1, n_samples, R_fit, f_samp, float(f_mod), opts,
                                        rows.data_ptr())
                    finally:
                        ctx.use_default_stream()
                    r = rows[:nw, :nbuf, :7].mean(dim=1)  # results_df.mean() over the trial's rows (experiments.py:79)
                    v = values[a["name"]]
                    v[lo:hi, :7] = r
                    v[lo:hi, 7] = r[:, 1] / (2.0 * np.pi * df_dev[lo:hi])  # tau (core.py:506-507)
            # ---- 3. grid statistics on the device ------------------------------------------------------------------
            local = {}
            for a in self.analyses:
                v = values[a["name"]]
                stats = torch.empty((npoints, len(RESULT_KEYS), _lib.TRIAL_STATS_DOUBLES), dtype=torch.float64, device=dev)
                if npoints:
                    ctx.use_torch_stream()
                    try:
                        ctx.trial_stats_dev(v.data_ptr(), npoints, ntr, len(RESULT_KEYS), len(RESULT_KEYS), stats.data_ptr())
                    finally:
                        ctx.use_default_stream()
                local[a["name"]] = (stats.cpu().numpy(), v.cpu().numpy().reshape(npoints, ntr, len(RESULT_KEYS)))
        if world > 1:
            parts = [None] * world
            torch.distributed.all_gather_object(parts, local, group=group)
        else:
            parts = [local]
        results = self._assemble_results(parts, grid_shape, ntr)
        self.results = results
        if filename is not None:
            self.save_results(filename)
        return results


def _ekf_opts(kw):
    opts = _lib.default_ekf_opts()
    for i, key in enumerate(("init_a", "init_m", "init_phi", "init_psi")):
        if key in kw:
            opts.init[i] = float(kw[key])
    for i in range(5):
        if "P0_diag" in kw:
            opts.p0_diag[i] = float(kw["P0_diag"][i])
        if "Q_diag" in kw:
            opts.q_diag[i] = float(kw["Q_diag"][i])
    if kw.get("R_val") is not None:
        opts.r_val = float(kw["R_val"])
    return opts
