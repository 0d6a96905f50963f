"""N > 1 path on CPU: the sample split across ranks + sum-reduce + resolve with world_size 2 over gloo.
Each rank runs the ORACLE for its share of the samples (the host logic under test is the split/seed/reduce/resolve
definition that bench.py drives with NCCL on the GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from gknextengine_b200 import RT, GkcuFrameParams, host

W, H, SPP = 64, 36, 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = host.Scene("CornellBox.proc", samples=SPP, bounces=4, taa=False)
    scene.update_nodes()
    sc, fr = O.OracleScene(scene), O.OracleFrame(W, H)
    ubo = scene.ubo(W, H, 7)
    p = GkcuFrameParams(temporal_frames=16, spp_world=world, spp_rank=rank, skip_post=1)
    fr.render(sc, ubo, p, threads=2, run_post=False)
    t = torch.from_numpy(fr.radiance())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        fr.resolve(ubo)
        np.save(out_path, fr.image(RT["SINGLE_DIFFUSE"]).copy())
    dist.barrier()
    dist.destroy_process_group()


def test_spp_split_reduce_resolve_world2(tmp_path):
    world = 2
    out = str(tmp_path / "single_diffuse.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out).astype(np.float32)
    # in-process simulation of both ranks by the oracle (rank order summation)
    scene = host.Scene("CornellBox.proc", samples=SPP, bounces=4, taa=False)
    scene.update_nodes()
    sc, fr = O.OracleScene(scene), O.OracleFrame(W, H)
    fr.render(sc, scene.ubo(W, H, 7), GkcuFrameParams(temporal_frames=16, spp_world=world), run_post=False)
    expect = fr.image(RT["SINGLE_DIFFUSE"]).astype(np.float32)
    assert np.array_equal(got, expect)
    # and the split changes the random streams but not the estimator: same image statistics as the 1-rank frame
    fr1 = O.OracleFrame(W, H)
    fr1.render(sc, scene.ubo(W, H, 7), GkcuFrameParams(temporal_frames=16, spp_world=1), run_post=False)
    one = fr1.image(RT["SINGLE_DIFFUSE"]).astype(np.float32)
    assert not np.array_equal(one, expect)
    lit = fr.image(RT["OBJEDCTID_0"]) != 65535
    assert abs(np.minimum(one[lit], 50).mean() - np.minimum(expect[lit], 50).mean()) < 0.35 * max(1e-3, np.minimum(one[lit], 50).mean())
