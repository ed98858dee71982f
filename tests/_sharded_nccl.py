"""torchrun helper of tests/test_gpu_multi.py: one process per GPU, NCCL only for the result gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_multi import _batches  # noqa: E402
from zk_state_proofs_b200.sharding import verify_sharded  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b, _ = _batches()
st, voff, vlen = verify_sharded(b)
if dist.get_rank() == 0:
    np.savez(sys.argv[1], st=st, voff=voff, vlen=vlen)
dist.barrier()
dist.destroy_process_group()
