"""CPU-only tests of the host-side PPO plumbing: network layout, mask packing, env sharding,
and the multi-rank logic (flat gradient bucket, scalar all-reduce) on a world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_network_layout_matches_reference_spec():
    from bbgpu.network import BlockBlastNetwork
    net = BlockBlastNetwork()
    assert sum(p.numel() for p in net.parameters()) == 5_290_113
    keys = list(net.state_dict().keys())
    assert keys[0] == "conv_encoder.0.weight" and "conv_encoder.6.conv1.weight" in keys
    assert "fc_encoder.0.weight" in keys and "policy_head.2.bias" in keys and "value_head.2.weight" in keys
    assert net.state_dict()["fc_encoder.0.weight"].shape == (512, 8192)
    assert net.state_dict()["policy_head.2.weight"].shape == (192, 256)
    logits, value = net.forward(torch.zeros(2, 8, 8), torch.zeros(2, 3, 8, 8), torch.ones(2, 192))
    assert logits.shape == (2, 192) and value.shape == (2,)
    m = torch.ones(2, 192)
    m[:, 5] = 0
    logits, _ = net.forward(torch.zeros(2, 8, 8), torch.zeros(2, 3, 8, 8), m)
    assert torch.isinf(logits[:, 5]).all() and torch.isfinite(logits[:, 6]).all()


@pytest.mark.reference
def test_network_is_state_dict_compatible_with_reference():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_gym_stub"))
    sys.path.insert(0, "/root/reference/src")
    sys.dont_write_bytecode = True
    from models.network import BlockBlastNetwork as Ref
    from bbgpu.network import BlockBlastNetwork
    ref, net = Ref(), BlockBlastNetwork()
    assert list(ref.state_dict().keys()) == list(net.state_dict().keys())
    net.load_state_dict(ref.state_dict())
    ref.eval(); net.eval()
    b, p = torch.rand(6, 8, 8).round(), torch.rand(6, 3, 8, 8).round()
    m = (torch.rand(6, 192) < 0.4).float()
    m[:, 0] = 1
    with torch.no_grad():
        lr, vr = ref.forward(b, p, m)
        ln, vn = net.forward(b, p, m)
        assert torch.allclose(vr, vn, atol=1e-5)
        assert torch.equal(torch.isinf(lr), torch.isinf(ln))
        assert torch.allclose(lr[torch.isfinite(lr)], ln[torch.isfinite(ln)], atol=1e-5)
        # differentiable evaluation path == the reference's get_action_and_value for given actions
        a = torch.multinomial(m, 1).squeeze(1)
        _, lp_r, en_r, _ = ref.get_action_and_value(b, p, m, action=a)
        _, lp_n, en_n, _ = net.evaluate_actions(torch.cat([b.unsqueeze(1), p], 1), m, a)
        assert torch.allclose(lp_r, lp_n, atol=1e-6) and torch.allclose(en_r, en_n, atol=1e-6)


def test_mask_plane_packing():
    from bbgpu.network import _pack_mask_planes
    rs = np.random.RandomState(0)
    m = rs.rand(7, 192) < 0.4
    m[0] = True
    pl = _pack_mask_planes(torch.from_numpy(m)).numpy().view(np.uint64)
    want = np.array([[sum(1 << k for k in range(64) if m[i, p * 64 + k]) for i in range(7)] for p in range(3)], dtype=np.uint64)
    assert np.array_equal(pl, want)


def test_env_sharding_covers_all_envs_once():
    from bbgpu import dist
    for total, world in ((1_048_576, 8), (64, 2), (10, 4), (7, 8)):
        got = [dist.shard(total, r, world) for r in range(world)]
        assert sum(n for _, n in got) == total
        pos = 0
        for off, n in got:
            assert off == pos
            pos += n
    assert dist.shard(1_048_576, 3, 8) == (393216, 131072)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from bbgpu import dist
    r, w, _ = dist.init("gloo")
    assert (r, w) == (rank, world) and dist.world_size() == world
    torch.manual_seed(100 + rank)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    dist.broadcast_module(lin)                                # everyone now has rank 0's weights
    w0 = torch.cat([p.detach().flatten() for p in lin.parameters()])
    bucket = dist.FlatGradBucket(lin.parameters())
    x = torch.full((4, 5), float(rank + 1))
    bucket.zero()
    lin(x).sum().backward()
    local = bucket.flat.clone()
    bucket.all_reduce_mean()                                  # C1: one flat all-reduce, averaged
    tot = dist.all_reduce_scalars([rank + 1.0, 10.0 * (rank + 1)], device="cpu")
    mx = dist.all_reduce_scalars([rank + 1.0], op="max", device="cpu")
    out.put((rank, w0.tolist(), local.tolist(), bucket.flat.tolist(), tot, mx, dist.shard(9)))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_gradient_bucket_and_scalars():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, l0, a0, t0, m0, s0), (r1, w1, l1, a1, t1, m1, s1) = res
    assert w0 == w1                                            # broadcast worked
    assert np.allclose(a0, a1) and np.allclose(a0, (np.array(l0) + np.array(l1)) / 2)
    assert l0 != l1
    assert t0 == t1 == [3.0, 30.0] and m0 == m1 == [2.0]
    assert s0 == (0, 5) and s1 == (5, 4)
