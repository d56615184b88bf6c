"""torchrun check: query-sharded decoder over WORLD_SIZE GPUs == single-GPU decoder."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from mvgformer_b200 import sharding, synthetic as syn
from helpers import scene_to
from test_gpu_parity import make_decoder

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, V, Q, L, J = 2, 5, 250, 3, 15         # uneven split for world=4/8
sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=3, levels=((32, 60), (16, 30), (8, 15)))
sd = syn.make_decoder_state_dict(L, np.random.default_rng(5))
dec = make_decoder(sc, sd, L)
THR = float(os.environ.get('THR', '0.1'))
scd = scene_to(sc, "cuda")
t_full = {k: scd[k] for k in ("tgt", "query_pos", "reference_points")}
with torch.no_grad():
    hs, refs, r2d, p2d, cls = dec(t_full["tgt"], t_full["reference_points"], scd["src_views"], scd["meta"],
                                  scd["spatial_shapes"], scd["level_start_index"], None,
                                  query_pos=t_full["query_pos"], threshold=THR)
    full_pose, full_prob = refs[-1], cls[-1]
    t = {k: sharding.shard_points(v, Q, J, rank, world) for k, v in t_full.items()}
    pose, prob = sharding.sharded_decoder_forward(dec, t["tgt"], t["reference_points"], scd["src_views"],
                                                  scd["meta"], scd["spatial_shapes"], scd["level_start_index"],
                                                  t["query_pos"], threshold=THR, num_queries=Q, joints=J,
                                                  rank=rank, world=world)
ok = torch.equal(prob, full_prob) and torch.equal(pose, full_pose)
print(f"rank {rank}/{world}: sharded == single-GPU: {ok}  (max |dpose| {float((pose - full_pose).abs().max()):.3e})", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
