"""CPU-side checks: the C-ABI library builds, loads and exports every declared symbol; the host-side mirror
of the reference's plugin surface (flags, registries, parameter trees, dataset dict) behaves; the N>1 gradient
exchange works over gloo with world_size 2.  No CUDA compute is invoked here."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import nemar_oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    from nemar_b200.engine import lib
    names = lib.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(built_lib, n)]
    assert not missing, "symbols declared in include/nemar_b200.h but not exported: %s" % missing
    assert built_lib.nemar_version() == 100


def test_bad_arguments_are_rejected_without_a_gpu(built_lib):
    from nemar_b200.engine import lib
    rc = built_lib.nemar_affine_grid_fwd(None, None, None, 0, 0, 0, None, None)
    assert rc == -1 and b"affine_grid_fwd" in built_lib.nemar_last_error()
    with pytest.raises(lib.EngineError):
        lib.check(rc, "nemar_affine_grid_fwd")


def test_no_cpu_fallback_in_product():
    """The product must not import the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nemar_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), "%s references the oracle" % f
    # the oracle is test infrastructure: besides tests/ only bench.py (cpu_baseline / --impl reference legs) and
    # __graft_entry__.py (smoke's checker, build of the C oracle) may touch it — not train.py, not the developer tools
    for rel in ["train.py"] + [os.path.join("scripts", f) for f in os.listdir(os.path.join(ROOT, "scripts")) if f.endswith(".py")]:
        src = open(os.path.join(ROOT, rel)).read()
        assert "import oracle" not in src and "from oracle" not in src, "%s imports the oracle" % rel


@pytest.mark.parametrize("case", ["c1_affine64", "c4_multires256"])
def test_parameter_trees_match_reference_keys(case):
    kw, batch, extra = H.CASE_FLAGS[case]
    cfg = O.OracleConfig(**kw)
    opt = H.engine_opt(cfg, batch, extra, gpu_ids="-1")
    from nemar_b200.models import networks, stn
    T = networks.define_G(3, 3, opt.ngf, opt.netG, opt.norm, False, opt.init_type, opt.init_gain, [])
    D = networks.define_D(6, opt.ndf, opt.netD, 3, opt.norm, opt.init_type, opt.init_gain, [])
    R = stn.define_stn(opt, opt.stn_type)
    for net, shapes in ((T, O.resnet_generator_shapes(ngf=cfg.ngf, n_blocks=cfg.n_blocks)),
                        (D, O.discriminator_shapes(ndf=cfg.ndf)),
                        (R, O.affine_stn_shapes(height=cfg.height, width=cfg.width) if cfg.stn_type == "affine"
                         else O.unet_stn_shapes())):
        sd = net.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)


def test_init_statistics_follow_reference_quirks():
    """down/1x1/refine convs ~ N(0,0.02) ('normal'); up convs fall back to kaiming (std ~0.047);
    identity-init output conv ~ N(0,1e-5)."""
    cfg = O.OracleConfig(stn_type="unet", height=256, width=256)
    opt = H.engine_opt(cfg, 1, [], gpu_ids="-1")
    from nemar_b200.models import stn
    torch.manual_seed(0)
    R = stn.define_stn(opt, "unet").state_dict()
    assert abs(float(R["offset_map.down_2.conv_0.conv2d.weight"].std()) - 0.02) < 2e-3
    assert abs(float(R["offset_map.up_3.conv2d.weight"].std()) - (2 / (1 + 0.04) / (128 * 9)) ** 0.5) < 3e-3
    assert float(R["offset_map.output.conv2d.weight"].std()) < 2e-5


def test_flags_and_registries():
    from nemar_b200 import data, models
    from nemar_b200.options.train_options import TrainOptions
    opt = TrainOptions().parse(["--dataroot", "x", "--dataset_mode", "synthetic", "--gpu_ids", "-1", "--checkpoints_dir",
                                "/tmp/nemar_b200_ckpt"], quiet=True)
    assert (opt.model, opt.stn_type, opt.stn_cfg, opt.netG, opt.gan_mode, opt.lr, opt.beta1) == \
        ("nemar", "affine", "A", "resnet_9blocks", "vanilla", 0.0002, 0.5)
    assert (opt.img_height, opt.img_width, opt.lambda_recon, opt.lambda_GAN, opt.lambda_smooth, opt.multi_resolution) == \
        (288, 384, 100.0, 1.0, 0.0, 1)
    assert models.find_model_using_name("nemar").__name__ == "NEMARModel"
    with pytest.raises(ModuleNotFoundError):
        models.find_model_using_name("does_not_exist")
    ds = data.create_dataset(opt)
    item = ds.dataset[3]
    assert set(item) == {"A", "B", "A_paths", "B_paths"} and item["A"].shape == (3, 288, 384)
    assert float(item["A"].min()) >= -1 and float(item["A"].max()) <= 1
    assert torch.equal(ds.dataset[3]["B"], item["B"])


def test_gradient_bucket_allreduce_gloo_world2(tmp_path):
    """N>1 path on CPU: two ranks, different shards, one all-reduce per bucket => identical averaged grads."""
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from nemar_b200.engine import parallel
world, rank, _ = parallel.init_process_group_from_env("gloo")
assert world == 2
full = torch.arange(8.).view(4, 2)
shard = parallel.shard_batch(full, rank, world)
assert shard.shape[0] == 2 and float(shard[0, 0]) == 4.0 * rank
bucket = torch.full((10,), float(rank + 1))
hook = parallel.BucketAllReduce()
scale = hook(bucket)
assert scale == 0.5 and torch.allclose(bucket * scale, torch.full((10,), 1.5)) and hook.calls == 1
dist.barrier()
print("rank", rank, "ok")
''' % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_cuda_graph_mode_state_machine(monkeypatch):
    """Host logic of --cuda_graph 1 (NEMARModel._optimize_parameters_graphed) with torch.cuda's graph API mocked: three
    eager steps, one capture, replays; a batch of another shape runs eagerly and does not leave its outputs bound; a
    learning-rate change forces a re-capture."""
    import types
    from nemar_b200.models import nemar_model as NM

    class FakeGraph:
        made = []

        def __init__(self):
            self.replays = 0
            FakeGraph.made.append(self)

        def replay(self):
            self.replays += 1

    class Ctx:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class FakeStream:
        def wait_stream(self, s):
            pass

    for name, val in (("CUDAGraph", FakeGraph), ("graph", Ctx), ("Stream", FakeStream), ("stream", Ctx),
                      ("current_stream", lambda: FakeStream()), ("synchronize", lambda: None)):
        monkeypatch.setattr(torch.cuda, name, val)
    m = object.__new__(NM.NEMARModel)
    m.opt = types.SimpleNamespace(cuda_graph=1, direction="AtoB", no_dropout=True)
    m.device = torch.device("cpu")
    groups = [{"lr": 2e-4}]
    m.optimizer_TR = types.SimpleNamespace(param_groups=groups)
    m.optimizer_D = types.SimpleNamespace(param_groups=groups)
    calls = []

    def eager():
        calls.append(float(m.real_A.sum()))
        m.loss_D = torch.tensor(float(len(calls)))     # a fresh tensor per eager step, like the real losses
    m._optimize_parameters_eager = eager
    batch = lambda v, n=2: {"A": torch.full((n, 3, 4, 4), float(v)), "B": torch.zeros(n, 3, 4, 4), "A_paths": "", "B_paths": ""}
    for k in range(3):                                  # eager warm-up steps
        m.set_input(batch(k))
        m._optimize_parameters_graphed()
    assert len(calls) == 3 and not FakeGraph.made
    # the step that would capture sees a ragged batch (short last batch of a small dataset): set_input bound temporaries,
    # so it must run eagerly and leave the capture to a later full-size step (else the graph reads stale buffers forever)
    m.set_input(batch(9, n=1))
    m._optimize_parameters_graphed()
    assert len(calls) == 4 and calls[-1] == 9.0 * 48 and not FakeGraph.made
    calls.pop()
    m.set_input(batch(3))
    m._optimize_parameters_graphed()                    # capture (runs the step once under the mocked capture) + replay
    assert len(calls) == 4 and len(FakeGraph.made) == 1 and FakeGraph.made[0].replays == 1
    captured_loss = m.loss_D
    m.set_input(batch(4))
    m._optimize_parameters_graphed()                    # replay: no eager call, the static input buffer holds the new batch
    assert len(calls) == 4 and FakeGraph.made[0].replays == 2 and float(m.real_A[0, 0, 0, 0]) == 4.0
    m.set_input(batch(5, n=1))                          # ragged last batch: cannot be copied into the captured buffers
    m._optimize_parameters_graphed()
    assert len(calls) == 5 and calls[-1] == 5.0 * 48 and FakeGraph.made[0].replays == 2
    assert m.loss_D is not captured_loss
    m.set_input(batch(6))
    m._optimize_parameters_graphed()                    # back to replays, and the captured outputs are bound again
    assert FakeGraph.made[0].replays == 3 and m.loss_D is captured_loss
    groups[0]["lr"] = 1e-4                              # update_learning_rate()
    m.set_input(batch(7))
    m._optimize_parameters_graphed()
    assert len(FakeGraph.made) == 2 and FakeGraph.made[1].replays == 1 and len(calls) == 6
    # dropout does not prevent the capture: its masks are keyed by a device-resident step counter
    m2 = object.__new__(NM.NEMARModel)
    m2.opt = types.SimpleNamespace(cuda_graph=1, direction="AtoB", no_dropout=False)
    m2.device = torch.device("cpu")
    m2.optimizer_TR = types.SimpleNamespace(param_groups=groups)
    m2.optimizer_D = types.SimpleNamespace(param_groups=groups)
    ran = []
    m2._optimize_parameters_eager = lambda: ran.append(1)
    for k in range(6):
        m2.set_input(batch(k))
        m2._optimize_parameters_graphed()
    assert len(ran) == 4 and len(FakeGraph.made) == 3 and FakeGraph.made[2].replays == 3


def test_reference_written_checkpoint_loads():
    """tests/golden/refckpt/*.pth were written by the REFERENCE's own BaseModel.save_networks
    (tests/golden/make_ref_checkpoint.py); the engine's modules must take them key for key, value for value."""
    from nemar_b200.models import networks
    d = os.path.join(ROOT, "tests", "golden", "refckpt")
    netT = networks.define_G(3, 3, 8, "resnet_6blocks", "instance", False, "normal", 0.02, ())
    netD = networks.define_D(6, 8, "basic", 3, "instance", "normal", 0.02, ())
    for net, name in ((netT, "T"), (netD, "D")):
        sd = torch.load(os.path.join(d, "latest_net_%s.pth" % name), map_location="cpu")
        assert list(sd.keys()) == list(net.state_dict().keys())
        net.load_state_dict(sd)
        assert all(torch.equal(v, sd[k]) for k, v in net.state_dict().items())


def test_device_prefetcher_passthrough_on_cpu():
    from nemar_b200.data.prefetch import DevicePrefetcher
    batches = [{"A": torch.full((2, 3, 4, 4), float(i)), "B": torch.zeros(2, 3, 4, 4), "A_paths": "a%d" % i} for i in range(3)]
    got = list(DevicePrefetcher(batches, "cpu"))
    assert [float(b["A"][0, 0, 0, 0]) for b in got] == [0.0, 1.0, 2.0] and got[1]["A_paths"] == "a1" and len(DevicePrefetcher(batches, "cpu")) == 3


def test_roofline_traffic_is_tied_to_the_conv_engine_sources(tmp_path, monkeypatch):
    """bench.py reports `roofline.traffic` from the committed ncu capture only while the capture's conv-engine digest equals
    the digest the library was built from; any other digest must give null (never a number of another build)."""
    import json
    import bench
    from nemar_b200 import build as B
    root = tmp_path / "repo"
    (root / "profiles").mkdir(parents=True)
    (root / "nemar_b200" / "build").mkdir(parents=True)
    dig = B.conv_tc_digest()
    assert dig == B.conv_tc_digest() and len(dig) == 64          # a pure function of the sources and flags
    prof = {"conv_tc_digest": dig, "file": "profiles/dominant_kernel_ncu.json",
            "per_launch": {"dram__bytes_read_MB": 36.5, "dram__bytes_write_MB": 0.5}}
    (root / "profiles" / "dominant_kernel_ncu.json").write_text(json.dumps(prof))
    (root / "nemar_b200" / "build" / "stamp_conv_tc").write_text(dig)
    monkeypatch.setattr(bench, "ROOT", str(root))
    traffic, note = bench.dominant_kernel_traffic()
    assert traffic == 37000000 and "256->256" in note
    (root / "nemar_b200" / "build" / "stamp_conv_tc").write_text("0" * 64)
    traffic, note = bench.dominant_kernel_traffic()
    assert traffic is None and "no ncu capture" in note


def test_launch_share_aggregation(tmp_path):
    """scripts/launch_shares.py: per-kernel shares of an ncu launch list (units normalised, commas in kernel names kept apart)."""
    import subprocess
    import sys
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"',
            '"0","1","python","h","k_a(int, float)","1","7","(256, 1, 1)","(8, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","30.0"',
            '"1","1","python","h","k_b()","1","7","(256, 1, 1)","(8, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","10,000"',
            '"2","1","python","h","k_a(int, float)","1","7","(256, 1, 1)","(8, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","30.0"',
            '"3","1","python","h","k_b()","1","7","(256, 1, 1)","(8, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","10,000"']
    src = tmp_path / "l.csv"
    src.write_text("==PROF== noise\n" + "\n".join(rows) + "\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "launch_shares.py"), str(src), "2", "note"],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "# note" and "2 launches and 0.04 ms" in out[1]
    assert out[3] == "75.00,0.030,1.0,30.0,k_a(int; float)" and out[4] == "25.00,0.010,1.0,10.0,k_b()"


def test_zero_arena_bookkeeping():
    """engine/functional.py::_ZeroArena (host logic, device-agnostic): slices handed out within a step are disjoint and
    zero, dirty slices are zero again after the next begin(), and outside a step zeros() falls back to torch.zeros."""
    from nemar_b200.engine.functional import _ZeroArena
    ar = _ZeroArena()
    free = ar.zeros((3, 5), "cpu")                       # not inside a step: an ordinary tensor
    assert free.shape == (3, 5) and float(free.abs().sum()) == 0.0 and ar.buf is None
    ar.begin("cpu")
    a, b = ar.zeros((2, 3, 2), "cpu"), ar.zeros((7,), "cpu")
    assert a.data_ptr() != b.data_ptr() and (b.data_ptr() - a.data_ptr()) % 16 == 0 and b.data_ptr() - a.data_ptr() >= a.numel() * 4
    a.fill_(3.0)
    b.fill_(5.0)
    assert float(a.sum()) == 36.0 and float(b.sum()) == 35.0      # disjoint: neither fill touched the other slice
    big = ar.zeros((ar.buf.numel(),), "cpu")             # does not fit behind a and b: falls back, the arena stays intact
    assert big.data_ptr() != ar.buf.data_ptr() and float(a.sum()) == 36.0
    ar.end()
    assert ar.zeros((4,), "cpu").data_ptr() != a.data_ptr()
    ar.begin("cpu")
    a2 = ar.zeros((2, 3, 2), "cpu")
    assert a2.data_ptr() == a.data_ptr() and float(a2.abs().sum()) == 0.0 and float(ar.buf[:64].abs().sum()) == 0.0


def test_pack_refresh_selects_and_stamps_the_right_entries(monkeypatch):
    """engine/functional.py::refresh_packs (host logic; the C-ABI calls are recorded, not executed): after an optimizer
    step ONE multi-tensor launch re-packs exactly the cache entries whose parameter lives in that optimizer's flat
    buffer — forward pack, data-gradient pack and padded bias each — and marks them current; entries of other
    optimizers stay untouched, entries derived from a parameter by a host-side transform are invalidated instead."""
    from nemar_b200.engine import functional as F
    from nemar_b200.engine import lib as L
    calls = []
    monkeypatch.setattr(F, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(F, "stream", lambda: None)
    flat = torch.zeros(4096)
    other = torch.zeros(4096)
    w1 = flat[0:6 * 3 * 3 * 3].view(6, 3, 3, 3)           # cout 6 (padded to 16), cin 3 (padded to 16), k3
    b1 = flat[200:206]
    w2 = flat[1024:1024 + 16 * 16].view(16, 16, 1, 1)
    w3 = other[0:16 * 16].view(16, 16, 1, 1)
    cfg1 = F.ConvCfg(3, 6, 3, pad=1, cout_p=16)
    cfg2 = F.ConvCfg(16, 16, 1)
    p1, p2, p3, pd = F.PackedWeights(), F.PackedWeights(), F.PackedWeights(), F.PackedWeights()
    p1.get(w1, b1, cfg1, torch.float32, 16)
    p2.get(w2, None, cfg2, torch.float32, 16)
    p3.get(w3, None, cfg2, torch.float32, 16)
    pd.get(w2.clone(), None, cfg2, torch.float32, 16, owner=w2)        # e.g. a tap-transformed copy of w2
    assert [c[0] for c in calls] == ["nemar_pack_weights"] * 4
    e1, e2, e3, ed = (list(p.cache.values())[0] for p in (p1, p2, p3, pd))
    assert e1["bp"] is not None and e2["bp"] is None and ed["derived"]
    calls.clear()
    p1.get(w1, b1, cfg1, torch.float32, 16)               # current: no re-pack
    assert calls == []
    stamp3 = e3["stamp"]
    e1["stamp"] = e2["stamp"] = "stale"
    F.refresh_packs(flat)
    assert [c[0] for c in calls] == ["nemar_pack_weights_multi"]
    nblocks = calls[0][1][2]
    jobs = 3 + 2                                          # e1: forward, backward, bias; e2: forward, backward
    per = lambda o, k, i: (o * k * i + 2047) // 2048
    assert nblocks == 2 * per(16, 9, 16) + per(16, 1, 1) + 2 * per(16, 1, 16)
    plan = F._PACK_PLANS[flat.data_ptr()]
    assert plan[1].numel() == jobs * L.C.sizeof(F._PackJob) and plan[2].numel() == 2 * nblocks
    assert e1["stamp"] == F.PackedWeights._stamp(w1, b1, None) and e2["stamp"] == F.PackedWeights._stamp(w2, None, None)
    assert e3["stamp"] == stamp3 and ed["stamp"] is None  # other optimizer untouched; derived entry invalidated
    calls.clear()
    pd.get(w2.clone(), None, cfg2, torch.float32, 16, owner=w2)
    assert [c[0] for c in calls] == ["nemar_pack_weights"]           # ... and re-packed by its next use
    F._PACK_PLANS.pop(flat.data_ptr(), None)
