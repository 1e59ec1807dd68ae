"""One rank of the multi-process GPV_GATHER test (tests/test_gpu_multiproc.py): one PROCESS per rank, the gathering rank's buffers
mapped through CUDA IPC handles (gpv_gather_create / gpv_gather_attach) -- the deployment shape of bench.py --gpus N, without
torchrun.  Rendezvous, the descriptor broadcast and the barriers go over torch.distributed (gloo, 127.0.0.1).

  python tests/mp_gather_worker.py RANK WORLD PORT DEVICE MESH L1 L2 OUT.json
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, device = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    path, l1, l2, out = sys.argv[5], int(sys.argv[6]), int(sys.argv[7]), sys.argv[8]
    import torch
    import torch.distributed as dist
    import gpview_b200 as gpv
    from gpview_b200 import binding as B
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mesh = gpv.load_mesh(path)
    ctx = gpv.Context(device)
    B._check(B.lib().gpv_gather_set_timeout(ctx.h, 30.0))   # ranks of this test may share one time-sliced GPU
    d_tris = ctx.upload(mesh)
    whole = ctx.voxelize_device(d_tris, mesh, gpv.Params(l1, l2, 0))    # grows the pools; rank 0 keeps the streams to compare with
    cells, nb, n23 = whole.cells, whole.nb, whole.n23
    want = (whole.level1_inout(), whole.prefix(), whole.level2_inout(), list(whole.counts)) if rank == 0 else None
    desc = B.CGatherDesc()
    if rank == 0:
        desc = ctx.gather_create(cells, nb * n23)
    t = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    desc = B.CGatherDesc.from_buffer_copy(t.numpy().tobytes())
    report = {"rank": rank, "epochs": []}
    for session in range(2):            # detach + attach again on the same buffers: a new session must not see the old one's flags
        ctx.gather_attach(desc, rank, world)
        dist.barrier()
        for epoch in range(2):
            if rank == 0:               # poison the buffers: every byte must be rewritten by this epoch's stores
                p1, p2, p3, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
                B._check(B.lib().gpv_gather_result(ctx.h, C.byref(p1), C.byref(p2), C.byref(p3), C.byref(n)))
                junk = np.full(max(cells * 4, nb * n23), 0x5a, np.uint8)
                for ptr, nbytes in ((p1, cells), (p2, cells * 4), (p3, nb * n23)):
                    B._check(B.lib().gpv_memcpy_h2d(ptr, junk.ctypes.data, nbytes, None))
                B._check(B.lib().gpv_stream_sync(None))
            dist.barrier()
            res = ctx.voxelize_device(d_tris, mesh, gpv.Params(l1, l2, gpv.GPV_GATHER), ctx.stream())
            entry = {"n_refined": res.n_refined, "nb": res.nb, "z": [res.z0, res.z1], "counts": list(res.counts)}
            if rank == 0:
                g1, gp, g2, gnb = ctx.gather_result(cells, n23)
                entry.update(ok_l1=bool(np.array_equal(g1, want[0])), ok_prefix=bool(np.array_equal(gp, want[1])), ok_l2=bool(np.array_equal(g2, want[2])),
                             ok_nb=gnb == nb, ok_counts=list(res.counts) == want[3])
            report["epochs"].append(entry)
            dist.barrier()
        ctx.gather_detach()
        dist.barrier()
    with open(out, "w") as f:
        json.dump(report, f)
    ctx.free_device(d_tris)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
