"""GPU: out-of-bounds write canaries.  compute-sanitizer is closed on this pool, so every output and
the workspace are carved out of a larger buffer pre-filled with a sentinel, and the guard zones on both
sides must be untouched after the launch (ragged N: tail tiles, TMA and plain-load paths)."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc

pytestmark = pytest.mark.gpu
GUARD = 4096          # floats on each side
SENT = -777.25


class Arena:
    def __init__(self, torch):
        self.t, self.bufs = torch, []

    def out(self, *shape):
        n = int(np.prod(shape))
        pad = (-n) % 4                                   # keep the payload 16-byte aligned
        big = self.t.full((GUARD + n + pad + GUARD,), SENT, device="cuda")
        self.bufs.append((big, n))
        return big[GUARD:GUARD + n].view(*shape)

    def check(self):
        for big, n in self.bufs:
            assert bool((big[:GUARD] == SENT).all()), "write before an output buffer"
            assert bool((big[GUARD + n + ((-n) % 4):] == SENT).all()), "write past an output buffer"


@pytest.mark.parametrize("N,Tn", [(132, 41), (131, 23), (260, 17), (1, 9)])
def test_no_out_of_bounds_writes(N, Tn):
    import torch
    from dynax_b200 import _lib
    L = _lib.lib()
    c = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    th, x0, u = c(orc.synth_theta(N, 151)), c(orc.synth_x0(N, 151)), c(orc.synth_u(Tn, N, 151))
    tg = c(22 + np.zeros((Tn, N)))
    ar = Arena(torch)
    xs, ys, xf = ar.out(Tn, N, 3), ar.out(Tn, N, 1), ar.out(N, 3)
    _lib.check(L.dynax_rc_forward_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                                      xf.data_ptr(), N, Tn, 3600.0, 0, None))
    loss, grad, gx0 = ar.out(N), ar.out(N, 7), ar.out(N, 3)
    nws = L.dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, N, Tn, 3, 5, 1, 0)
    ws = ar.out(max(4, (nws + 3) // 4))
    _lib.check(L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), tg.data_ptr(), None, None,
                                        loss.data_ptr(), grad.data_ptr(), gx0.data_ptr(), N, Tn, 3600.0, 0,
                                        ws.data_ptr(), ws.numel() * 4, None))
    gth, gx0b, gu = ar.out(N, 7), ar.out(N, 3), ar.out(Tn, N, 5)
    nwv = L.dynax_workspace_bytes(_lib.OP_RC_VJP, N, Tn, 3, 5, 1, 0)       # the adjoint's checkpoints
    ws2 = ar.out((nwv + 3) // 4)
    ones3, ones1 = torch.ones((Tn, N, 3), device="cuda"), torch.ones((Tn, N, 1), device="cuda")
    _lib.check(L.dynax_rc_vjp_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), ones3.data_ptr(), ones1.data_ptr(),
                                  gth.data_ptr(), gx0b.data_ptr(), gu.data_ptr(), N, Tn, 3600.0, 0, ws2.data_ptr(),
                                  ws2.numel() * 4, None))
    nwc = L.dynax_rc_loss_grad_chunked_workspace(N, Tn, 8)
    ws3 = ar.out((nwc + 3) // 4); loss3, grad3 = ar.out(N), ar.out(N, 7)
    _lib.check(L.dynax_rc_loss_grad_chunked_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), tg.data_ptr(), None, None,
                                                loss3.data_ptr(), grad3.data_ptr(), None, N, Tn, 3600.0, 0, 8,
                                                ws3.data_ptr(), ws3.numel() * 4, None))
    # loss only (g_theta = NULL) through the time-chunked entry point: same loss, no store through the NULL pointer
    loss4 = ar.out(N)
    _lib.check(L.dynax_rc_loss_grad_chunked_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), tg.data_ptr(), None, None,
                                                loss4.data_ptr(), None, None, N, Tn, 3600.0, 0, 8,
                                                ws3.data_ptr(), ws3.numel() * 4, None))
    torch.cuda.synchronize()
    assert torch.equal(loss4, loss3)
    W = [c(w) for w in orc.synth_mlp(152)]
    xn, yn = ar.out(Tn, N, 3), ar.out(Tn, N, 1)
    _lib.check(L.dynax_node_rk4_forward_f32(*[w.data_ptr() for w in W], x0.data_ptr(), u.data_ptr(), xn.data_ptr(),
                                            yn.data_ptr(), None, N, Tn, 3, 5, 1, 64, 0.1, 0, None))
    torch.cuda.synchronize()
    ar.check()
    assert bool(torch.isfinite(xs).all()) and bool(torch.isfinite(grad).all()) and bool(torch.isfinite(gu).all())
    assert bool(torch.isfinite(grad3).all()) and bool(torch.isfinite(xn).all())


def test_nonfinite_rows_flags_unstable_systems():
    """A parameter vector from the unstable corner of the box (Cai = 3e3, Ri = 0.5: dt*|lambda| = 2.6 > 2, SURVEY.md
    H3) blows up; dynax_nonfinite_rows_f32 names the systems, the others stay finite."""
    import torch
    from dynax_b200 import ops
    from oracle import dynax_oracle as orc
    N, Tn = 300, 600
    th, x0, u = orc.synth_theta(N, 91), orc.synth_x0(N, 91), orc.synth_u(Tn, N, 91)
    bad = [3, 77, 299]
    th[bad, 0] = 3.0e3; th[bad, 4] = 0.5
    c = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    _, _, xf = ops.rc_forward(c(th), c(x0), c(u), 3600.0, emit_x=False, emit_y=False, emit_final=True, time_chunks=0)
    flags, count = ops.nonfinite_rows(xf)
    assert count.item() == len(bad) and flags.nonzero().flatten().tolist() == bad
    tg = (22 + np.random.default_rng(5).normal(size=(Tn, N))).astype(np.float32)
    loss, grad, _ = ops.rc_loss_grad(c(th), c(x0), c(u), c(tg), 3600.0, time_chunks=0)
    assert ops.nonfinite_rows(loss)[0].nonzero().flatten().tolist() == bad
    assert ops.nonfinite_rows(grad)[1].item() == len(bad)
