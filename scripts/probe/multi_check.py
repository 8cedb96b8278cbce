import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = P.Context(local); P.set_default_context(ctx)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
I, J, K = 1024 // world, 1024, 256
t = torch.empty(I * J * K, device="cuda", dtype=torch.float32).uniform_(0, 1)
a = Array.from_device((usize, usize, usize), (I, J, K), t.data_ptr(), "f32", ctx=ctx, keep=t)
v = fold_rows(a.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))
tp = torch.empty(J * K, device="cuda", dtype=torch.float32)
o = Storage.wrap_device(ctx, F.F32, J * K, tp.data_ptr(), keep=tp)
v.collect(out=o)
torch.cuda.synchronize()
ref = t.view(I, J * K).double().sum(dim=0)
print(rank, "local rel", ((tp.double() - ref).abs() / ref.abs()).max().item(), flush=True)
v.collect(out=o, flags=F.COLLECT_ASYNC)
dist.all_reduce(tp)
torch.cuda.synchronize()
dist.all_reduce(ref)
torch.cuda.synchronize()
print(rank, "reduced rel", ((tp.double() - ref).abs() / ref.abs()).max().item(), tp[:4].tolist(), ref[:4].tolist(), flush=True)
dist.destroy_process_group()
