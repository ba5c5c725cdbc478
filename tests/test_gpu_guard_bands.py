"""Out-of-bounds write check of every kernel without compute-sanitizer (closed on this pool: gpurun
refuses it).  Every output and every input of a launch is carved out of ONE arena with a 4 KB band of
a known byte pattern on both sides; after the launch all bands must be untouched and (for inputs) the
payload unchanged.  Sizes straddle warp / block / vector-width boundaries, where an unguarded tail
would write past the end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BAND = 4096
SIZES = [1, 2, 31, 33, 127, 129, 1000, 4097]


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


class Arena:
    def __init__(self, torch, nbytes=1 << 26):
        self.t = torch
        self.buf = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device="cuda")
        self.off = BAND
        self.spans = []

    def alloc(self, shape, dtype, fill=None):
        t = self.t
        n = int(np.prod(shape)) * t.empty((), dtype=dtype).element_size()
        start = (self.off + 255) // 256 * 256
        self.off = start + n + BAND
        assert self.off < self.buf.numel()
        view = self.buf[start:start + n].view(dtype).view(*shape)
        if fill is not None:
            view.copy_(fill) if isinstance(fill, t.Tensor) else view.fill_(fill)
        self.spans.append((start, n))
        return view

    def check(self):
        self.t.cuda.synchronize()
        host = self.buf[:self.off].cpu().numpy()
        keep = np.ones(self.off, bool)
        for s, n in self.spans:
            keep[s:s + n] = False
        bad = np.flatnonzero(keep & (host != 0xA5))
        assert bad.size == 0, "guard band overwritten at arena offsets %s" % bad[:8]


@pytest.mark.parametrize("n", SIZES)
def test_env_kernels_stay_inside_their_buffers(torch, n):
    from bbgpu import capi
    A = Arena(torch)
    h = capi.EnvHandle(n, 5)
    act = A.alloc((n,), torch.int32)
    rew, term = A.alloc((n,), torch.float32), A.alloc((n,), torch.uint8)
    mask, board, pieces = A.alloc((3, n), torch.int64), A.alloc((n,), torch.int64), A.alloc((n,), torch.int32)
    eps, epl, info = A.alloc((n,), torch.int32), A.alloc((n,), torch.int32), A.alloc((n,), torch.int32)
    stats = A.alloc((8,), torch.int64, 0)
    ends = A.alloc((n * 32,), torch.uint8)
    h.set_episode_end_buffer(ends)
    h.observe(board, pieces, mask)
    for it in range(30):
        h.sample_valid_actions(it, act, None)
        h.step(act, rew, term, mask, eps, epl, info, board, pieces, stats)
    h.step_random(7, act, rew, term, mask, stats, mask)
    T = 5
    RA, RR, RT, RM = A.alloc((T, n), torch.int32), A.alloc((T, n), torch.float32), A.alloc((T, n), torch.uint8), A.alloc((T, 3, n), torch.int64)
    h.rollout_random(T, RA, RR, RT, RM, stats)
    rm = A.alloc((n,), torch.uint8, 1)
    h.reset(rm, mask)
    A.check()
    assert int(stats[0]) == n * (30 + 7 + T)
    h.close()


@pytest.mark.parametrize("n", SIZES)
def test_policy_and_gae_kernels_stay_inside_their_buffers(torch, n):
    from bbgpu import capi
    A = Arena(torch)
    g = torch.Generator(device="cuda").manual_seed(n)
    mask = A.alloc((3, n), torch.int64, torch.randint(-2 ** 62, 2 ** 62, (3, n), device="cuda", generator=g) | 1)
    board = A.alloc((n,), torch.int64, torch.randint(-2 ** 62, 2 ** 62, (n,), device="cuda", generator=g))
    pieces = A.alloc((n,), torch.int32, torch.randint(0, 37, (n,), device="cuda", generator=g).int() * 0x010101)
    for dt in (torch.float32, torch.bfloat16):
        logits = A.alloc((n, 192), dt, torch.randn(n, 192, device="cuda", generator=g).to(dt))
        before = logits.clone()
        act, lp, en = A.alloc((n,), torch.int32), A.alloc((n,), torch.float32), A.alloc((n,), torch.float32)
        for mode in (0, 1, 2):
            capi.masked_sample(logits, mask, n, 3, 1, mode, act, lp, en)
            capi.masked_sample(logits, mask, n, 3, 1, mode, act, lp, None)
        gl = A.alloc((n, 192), dt)
        g1, g2 = A.alloc((n,), torch.float32, 1.0), A.alloc((n,), torch.float32, 0.5)
        capi.masked_head_backward(logits, mask, n, act, g1, g2, gl)
        vals, gv = A.alloc((n,), torch.float32, 0.25), A.alloc((n,), torch.float32)
        sums = A.alloc((5,), torch.float64, 0)
        capi.ppo_loss(logits, mask, n, act, lp, g1, g2, vals, 0.2, 0.5, 0.01, gl, gv, sums)
        obs = A.alloc((n, 4, 8, 8), dt)
        dense = A.alloc((n, 192), torch.uint8)
        capi.unpack_obs(board, pieces, mask, n, obs=obs, mask_dense=dense, n=n)
        torch.cuda.synchronize()
        assert torch.equal(before, logits) and torch.isfinite(gl.float()).all() and torch.isfinite(lp).all()
    bf, pf, am = A.alloc((n, 8, 8), torch.float32), A.alloc((n, 3, 8, 8), torch.float32), A.alloc((n, 192), torch.int8)
    capi.unpack_obs_reference_layout(board, pieces, mask, n, bf, pf, am, n=n)
    # GAE: T x n with n not a multiple of the 4-env vector width
    T = 7
    r, v, d = (A.alloc((T, n), torch.float32, torch.rand(T, n, device="cuda", generator=g)) for _ in range(3))
    lv = A.alloc((n,), torch.float32, 0.5)
    adv, ret = A.alloc((T, n), torch.float32), A.alloc((T, n), torch.float32)
    mom = A.alloc((2,), torch.float64, 0)
    d.round_()
    capi.gae(r, v, d, lv, 0.99, 0.95, adv, ret, mom)
    # minibatch gather through an index
    B = max(1, n // 2)
    idx = A.alloc((B,), torch.int64, torch.randint(0, T * n, (B,), device="cuda", generator=g))
    gb = A.alloc((T, n), torch.int64, 7)
    gp = A.alloc((T, n), torch.int32, 0x020100)
    gm = A.alloc((T, 3, n), torch.int64, 3)
    ga = A.alloc((T, n), torch.int32, 1)
    ms = A.alloc((2,), torch.float32, torch.tensor([0.0, 1.0], device="cuda"))
    o_obs, o_m = A.alloc((B, 4, 8, 8), torch.float32), A.alloc((3, B), torch.int64)
    o_a, o_l, o_ad, o_r = A.alloc((B,), torch.int32), A.alloc((B,), torch.float32), A.alloc((B,), torch.float32), A.alloc((B,), torch.float32)
    capi.gather_minibatch(idx, n, gb, gp, gm, ga, r, adv, ret, ms, o_obs, o_m, o_a, o_l, o_ad, o_r)
    A.check()


@pytest.mark.parametrize("rows,ch", [(1, 8), (3, 64), (65, 128), (1000, 128), (4097, 64)])
def test_batchnorm_kernels_stay_inside_their_buffers(torch, rows, ch):
    from bbgpu import capi
    A = Arena(torch)
    mk = lambda: A.alloc((rows, ch), torch.bfloat16, torch.randn(rows, ch, device="cuda").to(torch.bfloat16))
    x, skip, dy = mk(), mk(), mk()
    y, dx, dsk = A.alloc((rows, ch), torch.bfloat16), A.alloc((rows, ch), torch.bfloat16), A.alloc((rows, ch), torch.bfloat16)
    gamma, beta = A.alloc((ch,), torch.float32, 1.0), A.alloc((ch,), torch.float32, 0.0)
    rm, rv = A.alloc((ch,), torch.float32, 0.0), A.alloc((ch,), torch.float32, 1.0)
    sm, sr, dg, db = (A.alloc((ch,), torch.float32) for _ in range(4))
    ws = A.alloc((capi.bn_workspace_size(ch),), torch.float32)
    capi.bn_relu_forward(x, skip, gamma, beta, None, rm, rv, 0.1, 1e-5, True, y, sm, sr, ws, rows, ch)
    capi.bn_relu_backward(x, y, dy, gamma, sm, sr, dx, dsk, dg, db, ws, rows, ch)
    dx2, dg2, db2 = A.alloc((rows, ch), torch.bfloat16), A.alloc((ch,), torch.float32), A.alloc((ch,), torch.float32)
    capi.bn_relu_forward(x, None, gamma, beta, None, rm, rv, 0.1, 1e-5, True, y, sm, sr, ws, rows, ch)
    capi.bn_relu_backward_no_skip(x, dy, gamma, beta, sm, sr, dx2, dg2, db2, ws, rows, ch)
    # same result as the variant that reads the ReLU mask from y
    capi.bn_relu_backward(x, y, dy, gamma, sm, sr, dx, None, dg, db, ws, rows, ch)
    A.check()
    assert torch.equal(dx, dx2) and torch.equal(dg, dg2) and torch.equal(db, db2)
    assert torch.isfinite(y.float()).all() and torch.isfinite(dx.float()).all()
