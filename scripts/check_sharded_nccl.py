"""Two (or N) ranks over NCCL: every rank reconstructs its `shard_clips` slice on its own GPU, `gather_by_clip`
concatenates codes and waveforms, rank 0 compares with the unsharded single-GPU result (must be bit-identical).
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_sharded_nccl.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from distilcodec_nabeel_b200 import Engine, Pipeline, gather_by_clip, random_init, shard_clips
from tests.golden.inputs import make_mel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_clips, T = 13, 300                                       # ragged split (13 = 7 + 6 at world 2)
mel = make_mel(n_clips, T, seed=99)
eng = Engine(random_init.make_state_dict("W1"), local, "bf16")
pipe = Pipeline(eng)
mine = shard_clips(n_clips, world, rank)
c, w = pipe.reconstruct_device(mel[mine.start:mine.stop].to(eng.device))
codes = gather_by_clip(c, n_clips)
wav = gather_by_clip(w, n_clips)
if rank == 0:
    c0, w0 = pipe.reconstruct_device(mel.to(eng.device))
    print(json.dumps({"world_size": world, "clips": n_clips, "shards": [len(shard_clips(n_clips, world, r)) for r in range(world)],
                      "codes_identical": bool(torch.equal(codes, c0)), "waveform_identical": bool(torch.equal(wav, w0))}))
dist.barrier()
dist.destroy_process_group()
