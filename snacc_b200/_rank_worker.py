"""One rank of ``ncd_matrix(..., gpus=N)`` / ``snacc --gpus N``: started by ``torch.distributed.run`` (see
sharding.run_multi_gpu), joins the NCCL group, computes its share and -- on rank 0 -- writes the result."""
import sys
from pathlib import Path

import numpy as np


def main(job_path, out_path):
    from . import sharding
    job = np.load(job_path)
    files = [Path(str(f)) for f in job["files"]]
    sharding.init_distributed()
    try:
        _, C, S, D = sharding.all_pairs(files, str(job["algorithm"]), bool(job["reverse_complement"]),
                                        bool(job["fast_mode"]))
        if sharding.is_rank0():
            np.savez(out_path, C=C, S=S, D=D)
    finally:
        sharding.shutdown_distributed()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
