"""Host-side argument checks of the Python layer: everything whose ADDRESS crosses the C boundary as an output is validated first
(the library writes prod(shape) elements of the dtype it was told, whatever the array really is).  These run without a GPU: the
checks fire before any library call."""
import numpy as np
import pytest

from aircraftoptimalcontrol_b200 import _lib as L
from aircraftoptimalcontrol_b200.batch import BatchedNewton, PipelinedNewton


def _handle(n, TT):
    bn = BatchedNewton.__new__(BatchedNewton)   # no context: every call below must raise before it would need one
    bn.N, bn.TT, bn._h = n, TT, None
    return bn


def test_out_array_rules():
    a = np.zeros((3, 6, 10))
    assert L.out_array(a, (3, 6, 10), np.float64) is a
    assert L.out_array(a.astype(np.float32), (3, 6, 10), (np.float32, np.float64)).dtype == np.float32
    for bad, why in ((a[:, :, ::2], "shape"), (np.zeros((3, 6, 10), dtype=np.float32), "dtype"), (np.zeros((3, 10, 6)).transpose(0, 2, 1), "contiguous"),
                     ([[0.0]], "numpy"), (np.zeros((3, 6, 11)), "shape")):
        with pytest.raises(ValueError, match=why):
            L.out_array(bad, (3, 6, 10), np.float64)
    ro = np.zeros((3, 6, 10))
    ro.setflags(write=False)
    with pytest.raises(ValueError, match="writeable"):
        L.out_array(ro, (3, 6, 10), np.float64)
    with pytest.raises(ValueError, match="contiguous"):
        L.ptr(np.zeros((4, 4)).T)
    assert L.ptr(None) is None


def test_result_outputs_are_checked_before_the_call():
    n, TT = 4, 16
    bn = _handle(n, TT)
    xs64, xs32, us = np.zeros((n, 6, TT)), np.zeros((n, 6, TT), dtype=np.float32), np.zeros((n, 2, TT))
    with pytest.raises(ValueError, match="dtype"):
        bn.result(out=(xs32, us))              # acoc_get_result writes float64 states
    with pytest.raises(ValueError, match="dtype"):
        bn.result_f32(out=(xs64, us))          # acoc_get_result_f32 writes float32 states
    with pytest.raises(ValueError, match="shape"):
        bn.result(out=(xs64, np.zeros((n, 2, TT + 1))))
    with pytest.raises(ValueError, match="shape"):
        bn.solve_deliver((np.zeros((n + 1, 6, TT)), us))
    with pytest.raises(ValueError, match="dtype"):
        bn.solve_deliver((xs64, us.astype(np.float32)))
    with pytest.raises(ValueError, match="x0"):
        bn.solve_deliver((xs64, us), x0=np.zeros((n, 5)))
    with pytest.raises(ValueError, match="contiguous"):
        bn.solve_deliver((np.zeros((n, TT, 6)).transpose(0, 2, 1), us))


def test_pipelined_solve_arguments_are_checked_before_any_upload():
    n, TT = 6, 12
    pn = PipelinedNewton.__new__(PipelinedNewton)
    pn.N, pn.TT, pn.parts, pn.bounds = n, TT, [], [0, n]
    xr, ur = np.zeros((n, 6, TT)), np.zeros((n, 2, TT))
    with pytest.raises(ValueError, match="either"):
        pn.solve()
    with pytest.raises(ValueError, match="either"):
        pn.solve(xr, ur, refs=("step", np.ones(n), np.ones(n)))
    with pytest.raises(ValueError, match="xx_ref has shape"):
        pn.solve(xr[:-1], ur)
    with pytest.raises(ValueError, match="pairs"):
        pn.solve(xr, None)
    with pytest.raises(ValueError, match="dx0"):
        pn.solve(xr, ur, dx0=np.zeros((n, 5)))
    with pytest.raises(ValueError, match="x_dtype"):
        pn.solve(xr, ur, x_dtype=np.float16)
    with pytest.raises(ValueError, match="dtype"):
        pn.solve(xr, ur, x_dtype=np.float32, out=(np.zeros((n, 6, TT)), np.zeros((n, 2, TT))))
    # with nothing to complain about and no sub-batches the call returns the (untouched) outputs
    xs, us, st = pn.solve(xr, ur)
    assert xs.shape == (n, 6, TT) and us.shape == (n, 2, TT) and st["iters"].shape == (n,)


class _FakePart:
    """Stands in for a BatchedNewton sub-batch: records what the pipeline asks of it, fails on request."""

    def __init__(self, n, TT, fail_in=None, log=None):
        self.N, self.TT, self.fail_in, self.log = n, TT, fail_in, log if log is not None else []

    def _step(self, what):
        self.log.append((what, self.N))
        if self.fail_in == what:
            raise RuntimeError("sub-batch of %d failed in %s" % (self.N, what))

    def set_refs(self, xr, ur):
        assert xr.shape == (self.N, 6, self.TT) and ur.shape == (self.N, 2, self.TT)
        self._step("set_refs")

    def init_guess(self, dx0=None):
        assert dx0 is None or dx0.shape == (self.N, 6)
        self._step("init_guess")

    def solve_deliver(self, out):
        self._step("solve_deliver")
        out[0][...] = self.N
        out[1][...] = -self.N
        return 7 * self.N, np.full((self.N, 6), 0.5)

    def stats(self):
        z = np.zeros(self.N)
        return dict(iters=np.full(self.N, 7, dtype=np.int32), status=np.ones(self.N, dtype=np.int32), J=z + self.N, descent=z, n_reg=z.astype(np.int32))


def _fake_pipeline(sizes, TT, fail=None):
    pn = PipelinedNewton.__new__(PipelinedNewton)
    pn.N, pn.TT = sum(sizes), TT
    pn.bounds = [0] + list(np.cumsum(sizes))
    log = []
    pn.parts = [_FakePart(n, TT, fail_in=(fail[1] if fail and fail[0] == k else None), log=log) for k, n in enumerate(sizes)]
    return pn, log


def test_pipeline_hands_every_sub_batch_its_slice_in_upload_order():
    sizes, TT = [3, 5, 4], 8
    pn, log = _fake_pipeline(sizes, TT)
    n = sum(sizes)
    xs, us, st = pn.solve(np.zeros((n, 6, TT)), np.zeros((n, 2, TT)), dx0=np.zeros((n, 6)), x_dtype=np.float32)
    assert [e for e in log if e[0] == "set_refs"] == [("set_refs", k) for k in sizes]       # uploads are serialised in sub-batch order
    lo = 0
    for k in sizes:   # every sub-batch wrote its own slice of the outputs and of the statistics
        assert (xs[lo:lo + k] == k).all() and (us[lo:lo + k] == -k).all() and (st["J"][lo:lo + k] == k).all()
        lo += k
    assert xs.dtype == np.float32 and st["x0"].shape == (n, 6) and (st["x0"] == 0.5).all() and (st["iters"] == 7).all()


@pytest.mark.parametrize("where", ["set_refs", "init_guess", "solve_deliver"])
def test_pipeline_surfaces_a_failing_sub_batch_and_does_not_hang(where):
    """A sub-batch that fails must not leave the others waiting for their upload turn, and later sub-batches do not start once an
    error is recorded during the upload phase (round-1 advisory finding)."""
    sizes, TT = [4, 4, 4, 4], 6
    pn, log = _fake_pipeline(sizes, TT, fail=(1, where))
    n = sum(sizes)
    with pytest.raises(RuntimeError, match="failed in " + where):
        pn.solve(np.zeros((n, 6, TT)), np.zeros((n, 2, TT)))
    if where == "set_refs":   # the failure happened while the others still waited for their turn: they never uploaded
        assert [e[0] for e in log].count("set_refs") == 2
