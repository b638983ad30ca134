"""torchrun worker: an Experiment whose grid points are split over the ranks must give the one-GPU result dictionary,
stochastic variables included (every rank walks the whole job list, so the generators are called in the same order).

    python -m torch.distributed.run --nproc-per-node N tests/multi/sharded_experiment.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import Experiment, factories, waveforms  # noqa: E402


class NoisyFactory(factories.StandardDFMIExperimentFactory):
    def __call__(self, params):
        cfg = super().__call__(params)
        cfg["laser_config"].amp_n = 1e-5
        return cfg


def phi_generator():
    return np.random.uniform(-1.0, 1.0)


def build(stochastic):
    exp = Experiment("sharded")
    exp.set_config_factory(NoisyFactory(waveforms.second_harmonic_distortion, opd_main=0.2))
    exp.add_axis("m_main", np.linspace(4.0, 12.0, 7))
    exp.add_axis("distortion_amp", np.array([0.0, 0.02, 0.04]))
    if stochastic:
        exp.add_stochastic_variable("phi", phi_generator)
    exp.n_trials = 5
    exp.n_fit_buffers_per_trial = 10
    exp.add_analysis("nls", "nls", fitter_kwargs={"ndata": 15})
    exp.add_analysis("ekf", "ekf", result_cols=["m", "phi"])
    return exp


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for stochastic in (False, True):
        np.random.seed(7)
        out = build(stochastic).run(group=dist.group.WORLD)
        if dist.get_rank() == 0:
            np.random.seed(7)
            ref = build(stochastic).run(device=local)
            for name in ("nls", "ekf"):
                for col, d in ref[name].items():
                    for key, val in d.items():
                        assert np.array_equal(out[name][col][key], val, equal_nan=True), (stochastic, name, col, key)
    if dist.get_rank() == 0:
        print(f"EXPERIMENT_OK world={dist.get_world_size()}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
