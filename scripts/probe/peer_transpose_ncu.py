#!/usr/bin/env python
"""The peer-mapped transpose of a row-sharded 16384^2 f32 Array, for an ncu capture with the NVLink sections: ONE process
drives GPU 0 and reads GPU 1's block through an in-process peer mapping (cudaDeviceEnablePeerAccess) — the kernel, its
tensor maps and its NVLink traffic are those of the 2-rank run, but no NCCL rendezvous has to survive ncu's kernel replay
(a two-process attempt hung in the communicator's barrier under ncu).
  ncu -k regex:k_transpose_tma --set full --section Nvlink --section Nvlink_Tables --clock-control none -o ... python peer_transpose_ncu.py"""
import ctypes, glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import multidimension_b200 as P
from multidimension_b200 import usize, Array, _ffi as F
from multidimension_b200.runtime import Storage
from multidimension_b200.sharding import PeerStorage, shard_view

ctx0, ctx1 = P.Context(0), P.Context(1)
M = N = 16384
world = 2
tb = M * N // world
blocks = [Storage.device(ctx0, F.F32, tb), Storage.device(ctx1, F.F32, tb)]
ctx0.upload(blocks[0].dptr, (np.arange(0, tb, dtype=np.int64) % 8191).astype(np.float32))
ctx1.upload(blocks[1].dptr, (np.arange(tb, 2 * tb, dtype=np.int64) % 8191).astype(np.float32))
ctx0.sync(); ctx1.sync()
import nvidia.cuda_runtime  # a shared libcudart (the product library links the runtime statically and exports no peer-access call);
rt = ctypes.CDLL(os.path.join(os.path.dirname(nvidia.cuda_runtime.__file__), "lib", "libcudart.so.12"))  # primary contexts are shared
assert rt.cudaSetDevice(0) == 0
rc = rt.cudaDeviceEnablePeerAccess(1, 0)
assert rc in (0, 704), rc  # 704: already enabled
P.set_default_context(ctx0)
peers = PeerStorage(F.F32, M * N, [blocks[0].dptr, blocks[1].dptr], tb, keep=blocks, ctx=ctx0)
view = shard_view(Array((usize, usize), (M, N), peers, "f32").transpose((), usize, usize, ()), 0, world)
out = Storage.device(ctx0, F.F32, tb)
p = view.prepare(out=out, flags=F.COLLECT_ASYNC)
for _ in range(3):
    p.run()
ctx0.sync()
got = out.to_numpy().reshape(N // world, M)
assert got[5, 9000] == np.float32((9000 * N + 5) % 8191) and got[7, 100] == np.float32((100 * N + 7) % 8191), got[5, 9000]   # out[x][y] = in[y][x]; y = 9000 lives on GPU 1
print("kernel:", ctx0.last_kernel(), flush=True)
