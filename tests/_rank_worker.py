"""worker for test_process_per_gpu_ranks_reduce_with_the_librarys_nccl: rank r of n on GPU r, no torch involved -- the NCCL id
travels through a file, the reduce is the library's own (trn_comm_init_rank / trn_render_rank)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from turner_b200 import api, scenes  # noqa: E402

rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
idfile = os.path.join(tmp, "nccl_id.bin")
if rank == 0:
    uid = api.Comm.unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid.tobytes())
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id after 120 s")
        time.sleep(0.05)
    uid = np.frombuffer(open(idfile, "rb").read(), np.uint8).copy()
comm = api.Comm(uid, world, rank, rank)
sc = scenes.fixture("cornell_box")
p = api.Scene.from_dict(sc)
cam, cfg = api.make_config(sc, 96, max_depth=3, mc_samples=2, pixel_samples=6, seed=12)
out = np.zeros((cfg.height, cfg.width, 4), np.float32) if rank == 0 else None
st = p.render_rank(comm, cam, cfg, out=out)
with open(os.path.join(tmp, "rays%d.txt" % rank), "w") as f:
    f.write(str(st.rays))
if rank == 0:
    assert st.ms_reduce > 0
    np.save(os.path.join(tmp, "image.npy"), out)
comm.close()
print("RANK_OK", rank, st.rays, "reduce ms %.3f d2h ms %.3f" % (st.ms_reduce, st.ms_d2h))
