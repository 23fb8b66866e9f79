"""Column sharding behind the C ABI, on ONE GPU (so the single-GPU test box exercises it):

  * `pharmsol_cuda_log_likelihood_matrix_peers` / `_push` with two "ranks" whose full matrices live on the same
    device and two column shards (first_col 0 and n/2): `col_base`, `ll_peers`, the copy-engine pushes and the
    global-pair keying of the SDE random streams;
  * `pharmsol_cuda_ctx_create_multi` with the device list [0, 0]: the in-library partition, the per-device replicas of
    the population, per-device copies into the caller's matrix, first-error reduction, replicated psi.

Every result must equal the single-launch matrix bit for bit (pairs are independent; matrix.rs:52-106)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("c1", 24, 300, {}), ("c2", 20, 2304, dict(solver="Dopri5", tol=1e-6)), ("c3", 12, 333, {}), ("c5", 5, 70, dict(particles=64))]


class _DevArr:
    """A raw device pointer as a CUDA array (column-major psi == C-order (nspp, nsub))."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _objects(ps, name, nsub, nspp, kw, device=0):
    from benches import harness as H, workloads as W
    w = W.make(name, nsub=nsub, nspp=nspp, **({"particles": kw["particles"]} if "particles" in kw else {}))
    eq, data, ems = H.product_objects(w, device=device)
    if "solver" in kw:
        eq.with_solver(getattr(ps.OdeSolver, kw["solver"])).with_tolerances(kw["tol"], kw["tol"])
    if "particles" in kw:
        eq.with_particles(kw["particles"]).with_mode(ps.SdeMode.ParticleFilter).with_seed(1234)
    return w, eq, data, ems


@pytest.mark.parametrize("name,nsub,nspp,kw", CASES)
@pytest.mark.parametrize("mode", ["peers", "push"])
def test_two_column_shards_on_one_gpu_equal_the_single_launch(ps, name, nsub, nspp, kw, mode):
    import torch
    from pharmsol_b200 import _lib
    w, eq, data, ems = _objects(ps, name, nsub, nspp, kw)
    ref = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    ctx = eq._ctx()
    pop = eq.population(data, ems)
    dev = torch.device("cuda", 0)
    full = [torch.full((nspp, nsub), float("nan"), dtype=torch.float64, device=dev) for _ in range(2)]     # column-major psi of "rank" 0 and 1
    peers = [t.data_ptr() for t in full]
    half = nspp // 2
    stream = torch.cuda.current_stream(dev).cuda_stream or 1
    for r, (lo, hi) in enumerate(((0, half), (half, nspp))):
        n = hi - lo
        soa = torch.empty((eq.nparams(), n), dtype=torch.float64, device=dev)
        _lib.upload_support_points(ctx, w["support_points"][lo:hi], soa.data_ptr(), n, stream)
        if mode == "peers":
            _lib.log_likelihood_matrix_peers(ctx, eq._model, pop, soa.data_ptr(), n, n, peers, nsub, lo, stream)
        else:
            _lib.log_likelihood_matrix_push(ctx, eq._model, pop, soa.data_ptr(), n, n, peers, r, nsub, lo, stream)
        torch.cuda.synchronize(dev)
        ctx.collect_errors()
    for t in full:
        assert np.array_equal(t.t().cpu().numpy(), ref, equal_nan=True)


@pytest.mark.parametrize("name,nsub,nspp,kw", CASES)
def test_multi_device_context_equals_single_device(ps, name, nsub, nspp, kw):
    import torch
    from pharmsol_b200 import _lib
    w, eq, data, ems = _objects(ps, name, nsub, nspp, kw)
    ref = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    pred_ref, offs = eq.predictions_matrix(data, w["support_points"])
    w2, eq2, data2, ems2 = _objects(ps, name, nsub, nspp, kw, device=[0, 0])
    ctx = eq2._ctx()
    assert ctx.num_devices == 2
    got = ps.log_likelihood_matrix(eq2, data2, w["support_points"], ems2)
    assert np.array_equal(got, ref, equal_nan=True)
    pred, offs2 = eq2.predictions_matrix(data2, w["support_points"])
    assert np.array_equal(offs, offs2) and np.array_equal(pred, pred_ref, equal_nan=True)
    # psi (exp on the device) through the sharded host call
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.array_equal(ps.psi(eq2, data2, w["support_points"], ems2), ps.psi(eq, data, w["support_points"], ems), equal_nan=True)
    # replicated psi: both gather modes, every device's matrix complete
    pop = eq2.population(data2, ems2)
    for gather in (_lib.GATHER_COPY_ENGINE, _lib.GATHER_PEER_STORES):
        ptrs = _lib.log_likelihood_matrix_replicated(ctx, eq2._model, pop, w["support_points"], gather)
        assert len(ptrs) == 2 and all(ptrs)
        for q in ptrs:
            torch.cuda.synchronize()
            host = torch.as_tensor(_DevArr(q, (nspp, nsub)), device="cuda:0").t().cpu().numpy()
            assert np.array_equal(host, ref, equal_nan=True)


def test_multi_device_first_error_is_the_lowest_global_pair(ps):
    """matrix.rs:96-104 over shards: imaginary roots (status 12) in a column of the SECOND shard and a later column of
    the first shard must report the first shard's pair."""
    eq = ps.Equation.from_dsl("name = twocpt\nkind = analytical\nparams = ke, kcp, kpc, v\nstates = central, peripheral\noutputs = cp\nbolus(iv) -> central\n"
                              "structure = two_compartments\nout(cp) = central / v ~ continuous()\n", device=[0, 0])
    ops = [("bolus", 0.0, 100.0, "iv"), ("observation", 1.0, 50.0, "cp")]
    data = ps.Data([ps.Subject(f"s{i}", ops) for i in range(5)])
    ems = ps.AssayErrorModels().add("cp", ps.AssayErrorModel.additive(ps.ErrorPoly(0.1, 0.1, 0, 0), 0.0))
    for bad, want in (([333], 333), ([333, 90], 90), ([150, 399], 150)):
        spp = np.tile(np.array([[0.1, 3.0, 1.0, 1.0]]), (400, 1))
        for b in bad:
            spp[b] = [1.0, -3.0, 1.5, 1.0]
        with pytest.raises(ps.PharmsolError) as e:
            ps.log_likelihood_matrix(eq, data, spp, ems)
        assert e.value.code == 12 and e.value.pair == want * 5


def test_wide_model_support_point_transpose(ps, oracle):
    """a 32-parameter model exceeds the 48 KB tile of the row-chunk transpose; the tiled kernel takes over."""
    names = [f"p{k}" for k in range(31)]
    src = ("name = wide\nkind = analytical\nparams = ke, " + ", ".join(names) + "\nstates = central\noutputs = cp\nbolus(iv) -> central\n"
           "structure = one_compartment\nout(cp) = central * (1 + 0 * (" + " + ".join(names) + ")) / p30 ~ continuous()\n")
    eq = ps.Equation.from_dsl(src)
    ops = [("bolus", 0.0, 100.0, "iv"), ("observation", 1.0, 50.0, "cp"), ("observation", 3.0, 20.0, "cp")]
    data = ps.Data([ps.Subject("s", ops)])
    rng = np.random.default_rng(3)
    spp = rng.uniform(0.5, 2.0, size=(700, 32))
    pred, _ = eq.predictions_matrix(data, spp)
    want = np.stack([100.0 * np.exp(-spp[:, 0] * t) / spp[:, 31] for t in (1.0, 3.0)])
    assert np.max(np.abs(pred - want) / np.abs(want)) <= 1e-13


def test_predictions_are_chunked_through_two_device_buffers(ps, monkeypatch):
    """estimate_predictions for a large grid is produced in column chunks through two device buffers (a C3-sized
    shard would otherwise need 5 GB at once).  Force tiny chunks (128 columns) and compare with the one-piece result,
    on one device and on the two-shard context."""
    from benches import harness as H, workloads as W
    w = W.make("c3", nsub=20, nspp=1100)
    eq, data, ems = H.product_objects(w)
    want, offs = eq.predictions_matrix(data, w["support_points"])
    monkeypatch.setenv("PHARMSOL_B200_PRED_CHUNK_KB", "64")          # 200 rows x 8 B -> 40 columns -> rounded up to 128
    got, _ = eq.predictions_matrix(data, w["support_points"])
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)
    eq2, data2, _ = H.product_objects(w, device=[0, 0])
    got2, _ = eq2.predictions_matrix(data2, w["support_points"])
    assert np.array_equal(got2, want, equal_nan=True)


def test_few_columns_fall_back_to_the_first_device_and_uneven_shards(ps):
    """Fewer than 32 columns per device: the call runs on device_ids[0] alone; 3 shards of an odd column count split 34/34/32."""
    from benches import harness as H, workloads as W
    w = W.make("c1", nsub=9, nspp=100)
    eq, data, ems = H.product_objects(w)
    ref = ps.log_likelihood_matrix(eq, data, w["support_points"], ems)
    eq3, data3, ems3 = H.product_objects(w, device=[0, 0, 0])
    assert np.array_equal(ps.log_likelihood_matrix(eq3, data3, w["support_points"], ems3), ref)
    assert np.array_equal(ps.log_likelihood_matrix(eq3, data3, w["support_points"][:40], ems3), ref[:, :40])      # 40 < 3 * 32
    assert np.array_equal(ps.log_likelihood_matrix(eq3, data3, w["support_points"][:1], ems3), ref[:, :1])        # latency path


def test_long_timelines_bypass_the_shared_memory_staging(ps, oracle):
    """A subject whose timeline program exceeds the staged window (96 records) is executed from global memory; the CTA
    mixes both kinds of subject.  150 observations + 6 infusions vs the oracle, next to a short subject."""
    import fixtures as FX
    kernel = "two_compartments"
    long_ops = [("infusion", float(6 * k), 100.0, "0", 1.5) for k in range(6)] + [("observation", 0.25 * k + 0.1, 1.0 + 0.01 * k, "0") for k in range(150)]
    short_ops = [("bolus", 0.0, 100.0, "0"), ("observation", 1.0, 2.0, "0"), ("observation", 5.0, 1.0, "0")]
    subjects = [("long", long_ops), ("short", short_ops), ("long2", long_ops[::-1])]
    rng = np.random.default_rng(2)
    spp = np.column_stack([rng.uniform(0.05, 1.0, 300), rng.uniform(0.05, 1.0, 300), rng.uniform(0.05, 1.0, 300), rng.uniform(5, 80, 300)])
    eq = ps.Equation.from_dsl(FX.kernel_dsl(kernel))
    data = ps.Data([ps.Subject(i, o) for i, o in subjects])
    em = ("additive", 0.05, (0.1, 0.15, 0.0, 0.0))
    ems = ps.AssayErrorModels().add("outeq_0", ps.AssayErrorModel.additive(ps.ErrorPoly(*em[2]), em[1]))
    psi = ps.log_likelihood_matrix(eq, data, spp, ems)
    ref = oracle.Model(kernel).log_likelihood_matrix(oracle.Data([oracle.Subject(o, i) for i, o in subjects]), spp, oracle.ErrorModels([em]))
    assert np.all(np.isfinite(psi)) and np.max(np.abs(psi - ref) / (np.abs(ref) + 150)) <= 1e-12
