"""torchrun worker: the Monte-Carlo sweep split over the ranks must equal the one-GPU sweep (same seeds per realisation).

    python -m torch.distributed.run --nproc-per-node N tests/multi/sharded_sweep.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepfmkit_b200.montecarlo import nls_sweep, nls_sweep_sharded  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ms, n_trials = [4.0, 9.0, 15.0], 3001
    out = nls_sweep_sharded(ms, n_trials, device=local, seed=11, ndata=15, snr_db=30.0)
    if dist.get_rank() == 0:
        ref = nls_sweep(ms, n_trials, device=local, seed=11, ndata=15, snr_db=30.0)
        for key in ("m_mean", "m_std", "m_min", "m_max", "m_worst", "amp_mean", "phi_std", "fitok", "ssq_mean", "crlb_sigma_m"):
            assert np.allclose(out[key], ref[key], rtol=1e-9, atol=1e-12), (key, out[key], ref[key])
        assert out["n_trials"] == n_trials
        print(f"SWEEP_OK world={dist.get_world_size()}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
