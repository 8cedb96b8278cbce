"""One rank of the missing-peer test (tests/test_comm_multi_gpu.py): `python timeout_worker.py <rank> <world> <rendezvous file> <route>`.
Both ranks run one fold over the sharded axis together (which also sets the exchange areas up); then rank 0 calls it ALONE: the kernel's
bounded spins must give up (~2 s) and mdim_fold_sharded_axis_status must report MDIM_ERR_NCCL — not a hung GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import multidimension_b200 as P
from multidimension_b200 import Add, _ffi as F
from multidimension_b200.runtime import Storage
from multidimension_b200.sharding import Comm


def main():
    rank, world, path, route = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    ctx = P.Context(rank)
    comm = Comm.from_env(ctx, path=path)
    rows, cols = 16, 4096
    x = np.full(rows * cols, np.float32(rank + 1))
    st = Storage.device(ctx, F.F32, rows * cols)
    ctx.upload(st.dptr, x)
    blocked = route == "blocked"
    got = comm.fold_sharded_axis(st, rows, cols, Add, np.float32(0), blocked=blocked).to_numpy()
    comm.fold_status()
    assert np.all(got == np.float32(rows * sum(range(1, world + 1)))), got[:4]
    comm.barrier()
    if rank == 0:  # alone: nobody sends this launch's lines
        t0 = time.time()
        comm.fold_sharded_axis(st, rows, cols, Add, np.float32(0), blocked=blocked)
        try:
            comm.fold_status()
        except F.MdimError as e:
            assert e.status == F.ERR_NCCL, e.status
            print(f"rank 0: gave up after {time.time() - t0:.1f} s: {e}", flush=True)
        else:
            raise AssertionError("a fold without its peers reported success")
        comm.fold_status()  # the error word was cleared
    comm.barrier()
    comm.close_peers()
    comm.close()
    ctx.close()
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
