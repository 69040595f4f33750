"""Cross-rank primitives alone (torchrun, one rank per GPU): the flat-gradient all-reduce of csrc/peer.cu (multimem and
peer load/store paths, several grid sizes) against NCCL, and the BatchNorm statistics exchange against an NCCL all-reduce of
the same size.  Device time per call from CUDA events over back-to-back calls, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from video_rep_learning_b200 import parallel, _lib as L

lib = L.lib()


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


out = {"world": world}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_800_000
flat = parallel.flat_grad_buffer(n, dev)
obj = [o for o in parallel.PeerFlatGrads._cache.values() if o.owns(flat)][0]
st = torch.cuda.current_stream().cuda_stream
mc = obj.mc_ptr
for name, mcp in (("multimem", mc), ("peer_ldst", 0)):
    if name == "multimem" and not mc:
        continue
    for ctas in (8, 16, 32, 64):
        call = lambda: L.check(lib.mvf_peer_allreduce_f32(mcp, obj.ptrs_dev, 0, obj.flag_off, obj.elems, obj.rank, obj.world,
                                                          obj.counters.data_ptr(), ctas, st))
        out[f"allreduce_{name}_ctas{ctas}_us"] = timeit(call)
x = torch.zeros(n, device=dev)
out["allreduce_nccl_us"] = timeit(lambda: dist.all_reduce(x))
stats = torch.zeros(1024, dtype=torch.float64, device=dev)
out["bn_exchange_peer_us"] = timeit(lambda: parallel.sync_stats_(stats), reps=100)
os.environ["MVF_PEER_BN"] = "0"
out["bn_exchange_nccl_us"] = timeit(lambda: dist.all_reduce(stats), reps=100)
out["floats"] = n
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
