"""world_size-2 gloo tests of the multi-GPU plumbing (point sharding, event sharding) on the CPU.
The per-rank evaluation is the oracle here (tests may use it); on GPUs it is ll.batch / ll.batch_parts."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from blueice_b200 import distributed as bdist


def test_shard_bounds():
    assert bdist.shard_bounds(10, 2) == [(0, 5), (5, 10)]
    assert bdist.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert bdist.shard_bounds(1, 2) == [(0, 1), (1, 1)]
    assert bdist.shard_bounds(0, 3) == [(0, 0)] * 3
    b = bdist.shard_bounds(100000, 8, align=512)
    assert b[0][0] == 0 and b[-1][1] == 100000
    assert all(lo % 512 == 0 for lo, _ in b) and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) < 2 * 512   # one block + the ragged tail


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle.pipeline import UnbinnedOracle
    from oracle import unbinned as ounb
    import bench_workloads as wl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        axes, mus, ps, x = wl.c1_arrays(seed=3, n_expected=2000.)
        zs, mult = wl.scan_points(37, 1, 1, seed=5, mult_range=(0.5, 2.0))
        zs[3, 0] = 9.0                                              # out of range -> -inf
        names = ['mu', 's0_rate_multiplier']
        params = np.column_stack([zs, mult])

        class OracleLL(object):                                     # stands in for ll on one rank
            def __init__(self, ps_local):
                self.orc = UnbinnedOracle(axes, mus).set_ps(ps_local)

            def batch(self, p, names, livetime_days=None):
                return self.orc.batch(p[:, :1], p[:, 1:])

            def batch_parts(self, p, names, livetime_days=None):
                logsum, musum, status = [], [], []
                for row in p:
                    r = self.orc(row[:1], row[1:], full_output=True)
                    if r == -np.inf:
                        logsum.append(0.); musum.append(0.); status.append(1)
                        continue
                    _, m, pp = r
                    with np.errstate(all='ignore'):
                        dens = np.nansum(m[:, None] * pp, axis=0)
                        dens[True ^ (dens > 0)] = 1e-12
                    logsum.append(np.sum(np.log(dens))); musum.append(m.sum()); status.append(0)
                return np.array(logsum), np.array(musum), np.array(status), np.zeros(len(p))

        full = OracleLL(ps).batch(params, names)
        # point sharding: every rank ends up with the full, identical result
        got = bdist.PointShardedLikelihood(OracleLL(ps), None).batch(params, names)
        assert np.array_equal(got, full)
        # event sharding: superblock-aligned slices, fixed rank-order sum
        lo, hi = bdist.shard_bounds(ps.shape[-1], world, align=512)[rank]
        sharded = bdist.EventShardedLikelihood(OracleLL(ps[:, :, lo:hi]), None)
        got_ev = sharded.batch(params, names)
        assert np.isneginf(got_ev[3]) and np.isneginf(full[3])
        fin = np.isfinite(full)
        np.testing.assert_allclose(got_ev[fin], full[fin], rtol=1e-13)
        one = sharded(mu=float(zs[0, 0]), s0_rate_multiplier=float(mult[0, 0]))
        assert one == got_ev[0]
        # toy sharding: rank r owns a contiguous slice of the toys; the stand-in ll evaluates toy t = events
        # [offsets[t], offsets[t+1]) at params[t] with the oracle, on whatever slice it was given
        from oracle.pipeline import toy_loglikelihoods
        t_axes, t_edges, t_templates, t_mus = wl.c2_arrays(2, 1, (-1., 0., 1.), (12, 10))
        sizes = np.array([30, 0, 55, 41, 17])
        offsets = np.concatenate([[0], np.cumsum(sizes)])
        tx, ty = wl.c2_events(t_templates, t_mus, t_edges, int(offsets[-1]) + 5, seed=2)
        tzs, tmult = wl.scan_points(5, 1, 2, seed=9, z_range=(-1., 1.))

        class ToyLL(object):
            def set_toy_data(self, datasets, offsets=None):
                self.coords, self.offsets = datasets, offsets

            def batch_toys(self, p, names, livetime_days=None):
                return toy_loglikelihoods(t_axes, t_mus, t_templates, t_edges, self.coords, self.offsets, p[:, :1], p[:, 1:])

        tparams = np.column_stack([tzs, tmult])
        whole = ToyLL()
        whole.set_toy_data([tx[:offsets[-1]], ty[:offsets[-1]]], offsets)
        toys_full = whole.batch_toys(tparams, None)
        sharded_toys = bdist.ToyShardedLikelihood(ToyLL(), None)
        lo_t, hi_t = bdist.shard_bounds(5, world)[rank]
        sl = slice(offsets[lo_t], offsets[hi_t])
        sharded_toys.set_local_toys(5, [tx[sl], ty[sl]], offsets[lo_t:hi_t + 1] - offsets[lo_t])
        assert np.array_equal(sharded_toys.batch_toys(tparams), toys_full)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), got_ev)
        assert bdist.rank_ordered_sum(np.array([float(rank + 1)])).tolist() == [3.0]
    finally:
        dist.destroy_process_group()


def test_point_and_event_sharding_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)                                      # bit-identical on every rank
