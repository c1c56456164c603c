"""Several GPUs behind one C handle (hmx_multi_*): a host batch is split into contiguous blocks, one per
device slot, solved concurrently; log-likelihoods are bitwise independent of the split, the covariance agrees to rounding.
On a one-GPU box the slots are two contexts on the same device (the sharding logic is identical); with more GPUs visible
the first two are used."""
import numpy as np
import pytest

from tmcmc202601_b200 import workloads

pytestmark = pytest.mark.gpu


def test_multi_handle_shards_a_host_batch(built):
    import torch

    from tmcmc202601_b200.engine import HamiltonEngine, HmxError, MultiEngine

    wl = workloads.four_condition_workload(37, seed=21, interleave=True)  # 148 evaluations: ragged blocks for 3 slots
    single = HamiltonEngine(wl.config, device=0)
    ndev = torch.cuda.device_count()
    slots = [[0], [0, 0], [0, 1 % ndev, 0]]
    try:
        ref, st_ref, it_ref = single.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        rng = np.random.default_rng(0)
        w = rng.random(wl.theta.shape[0])
        w /= w.sum()
        mean_ref, cov_ref = single.weighted_cov(wl.theta, w)
        for devs in slots:
            m = MultiEngine(wl.config, devs)
            try:
                ll, st, it = m.loglik_batch(wl.theta, wl.cond_id, return_status=True)
                assert np.array_equal(ll, ref) and np.array_equal(st, st_ref) and np.array_equal(it, it_ref), devs
                mean, cov = m.weighted_cov(wl.theta, w)
                assert np.allclose(mean, mean_ref, rtol=1e-13, atol=0) and np.allclose(cov, cov_ref, rtol=1e-10, atol=1e-18), devs
                # the replicated stage arithmetic runs on the first slot's context with the gathered population
                db, _ = m.stage_engine.beta_step(ll, 0.0, 0.5, 0.02, 0.5)
                assert db == single.beta_step(ref, 0.0, 0.5, 0.02, 0.5)[0]
                # N smaller than the number of slots, and N = 0
                assert np.array_equal(m.loglik_batch(wl.theta[:1], wl.cond_id[:1]), ref[:1])
                assert m.loglik_batch(wl.theta[:0], wl.cond_id[:0]).size == 0
                bad = wl.cond_id.copy()
                bad[-1] = 99
                with pytest.raises(HmxError):
                    m.loglik_batch(wl.theta, bad)
            finally:
                m.close()
        with pytest.raises(HmxError):
            MultiEngine(wl.config, [0, 4096])
    finally:
        single.close()


def test_contexts_on_their_own_streams_run_side_by_side(built):
    """Several contexts on ONE device, driven from host threads (run_batch_TMCMC(merge_launches=False)): results are bitwise
    those of a single context, also while the contexts grow and release their scratch buffers next to each other's running
    kernels (the buffers come from the stream-ordered allocator, so nothing synchronises the device)."""
    import threading

    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(64, seed=5, interleave=True)
    single = HamiltonEngine(wl.config, device=0)
    ref = single.loglik_batch(wl.theta, wl.cond_id)
    single.close()
    out, errs = {}, []

    def work(k):
        try:
            eng = HamiltonEngine(wl.config, device=0)
            for n in (16, 256, 40, 256):  # growing and re-used scratch
                out[(k, n)] = eng.loglik_batch(wl.theta[:n], wl.cond_id[:n])
            w = np.full(256, 1.0 / 256)
            out[(k, "idx")] = eng.resample(w, 0.25)
            eng.close()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for k in range(4):
        for n in (16, 256, 40):
            assert np.array_equal(out[(k, n)], ref[:n])
        assert np.array_equal(out[(k, "idx")], np.arange(256))


def test_stream_switch_keeps_scratch_valid(built):
    """hmx_set_stream after buffers were allocated on the context's own stream: the old stream is drained, later growth and
    release happen in the new stream's order; values unchanged."""
    import torch

    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(48, seed=9)
    eng = HamiltonEngine(wl.config, device=0)
    try:
        a = eng.loglik_batch(wl.theta[:32], wl.cond_id[:32])
        eng.use_torch_stream()
        th = torch.from_numpy(wl.theta).cuda()
        cid = torch.from_numpy(wl.cond_id).cuda()
        ll = torch.empty(th.shape[0], dtype=torch.float64, device="cuda")
        eng.loglik_batch_device(th, cid, ll)  # larger batch: every scratch buffer grows on the new stream
        torch.cuda.synchronize()
        b = ll.cpu().numpy()
        assert np.array_equal(b[:32], a)
        assert np.array_equal(eng.loglik_batch(wl.theta, wl.cond_id), b)
    finally:
        eng.close()
