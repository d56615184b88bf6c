"""torchrun check (N GPUs): mvg_allgather_poses with the process group's own ncclComm_t ==
sharding.gather_results.   torchrun --nproc-per-node 2 tools/check_allgather_cabi.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from mvgformer_b200 import _lib, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
B, Q, J, L = 2, 250, 15, 4
q0, q1 = sharding.shard_bounds(Q, rank, world)
rng = np.random.default_rng(rank)
poses = torch.from_numpy(rng.standard_normal((B, (q1 - q0) * J, 3)).astype(np.float32)).to(dev)
prob = torch.from_numpy(rng.uniform(size=(B, q1 - q0, 2)).astype(np.float32)).to(dev)
counts = torch.tensor([rank + 1, 0, 5, 2 * rank], dtype=torch.int32, device=dev)
want = sharding.gather_results(poses, prob, counts, Q, J, world)
dist.barrier()                                             # creates the communicator
comm = dist.distributed_c10d._get_default_group()._get_backend(dev)._comm_ptr()
o_pose = torch.zeros((B, Q * J, 3), device=dev); o_prob = torch.zeros((B, Q, 2), device=dev); o_cnt = torch.zeros(L, device=dev)
ws = torch.empty((int(lib.mvg_allgather_poses_workspace_bytes(B, Q, J, L, world)),), dtype=torch.uint8, device=dev)
_lib.check(lib.mvg_allgather_poses(comm, rank, world, poses.data_ptr(), prob.data_ptr(), counts.data_ptr(), B, Q, J, L,
                                   o_pose.data_ptr(), o_prob.data_ptr(), o_cnt.data_ptr(), ws.data_ptr(),
                                   _lib.stream_ptr(dev)), "mvg_allgather_poses")
torch.cuda.synchronize()
ok = torch.equal(o_pose, want[0]) and torch.equal(o_prob, want[1]) and torch.equal(o_cnt, want[2])
print(f"rank {rank}/{world}: mvg_allgather_poses == gather_results: {ok}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
