"""CPU, world_size 2, gloo: the sharding helpers of ppea_depth_b200.dist and the property the N>1
bench relies on -- a batch-sharded run with per-rank normalisation equals evaluating the shards
independently, and the optional global statistics reproduce the single-process masked mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vsl_oracle as O
from ppea_depth_b200 import dist as D
from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise

B, H, W, S = 4, 24, 48, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sums_vector(maps, batch):
    """The layout of ppea_vsl_forward's `sums` (include/ppea_vsl.h), filled from oracle maps."""
    out = torch.zeros(S, 8 + 4 * batch)
    for s in range(S):
        out[s, 0] = float((maps[s]["r"] * maps[s]["mask"]).sum())
        out[s, 1] = float(maps[s]["mask"].sum())
    return out.reshape(-1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=9)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    ins, outs = D.shard_batch(inputs, B, rank, world), D.shard_batch(outputs, B, rank, world)
    nz = [D.shard_batch({"z": z}, B, rank, world)["z"] for z in noise]
    per = B // world
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=per)
    losses, grads, maps = O.run_fwd_bwd(ins, outs, opt, False, nz, want_maps=True)
    stats = D.global_loss_stats(_sums_vector(maps, per), per, S)
    g = [grads[("disp", 0)].mean().reshape(1).clone()]
    D.allreduce_mean_(g)
    q.put((rank, float(losses["loss"]), D.global_reproj_loss(stats).tolist(), float(g[0]),
           float(grads[("disp", 0)].mean()), tuple(ins[("color", 0, 0)].shape)))
    dist.destroy_process_group()


def test_sharded_equals_independent_shards_and_global_stats():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=9)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    # (1) each rank's loss == the same sub-batch evaluated alone (rank-local normalisation, no collective)
    for rank, loss, _, _, _, shape in res:
        assert shape == (B // world, 3, H, W)
        lo, hi = D.shard_range(B, rank, world)
        ins = {k: v[lo:hi] for k, v in inputs.items()}
        outs = {k: v[lo:hi] for k, v in outputs.items()}
        opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B // world)
        ref, _, _ = O.run_fwd_bwd(ins, outs, opt, False, [z[lo:hi] for z in noise])
        assert abs(loss - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    # (2) the all-reduced statistics give the single-process (global-batch) masked mean
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    full, _, _ = O.run_fwd_bwd(inputs, outputs, opt, False, noise)
    for s in range(S):
        assert abs(res[0][2][s] - float(full["reproj_loss/%d" % s])) <= 2e-6 * float(full["reproj_loss/%d" % s])
        assert res[0][2][s] == res[1][2][s]
    # (3) allreduce_mean_ averages like DDP
    assert abs(res[0][3] - 0.5 * (res[0][4] + res[1][4])) < 1e-9 and res[0][3] == res[1][3]


def test_shard_range_rejects_ragged_batches():
    import pytest
    with pytest.raises(ValueError):
        D.shard_range(10, 0, 4)
    assert D.shard_range(96, 3, 8) == (36, 48)
