"""bench.py — paired 256x256 samples/sec of NeMAR's training step (NEMARModel.optimize_parameters) on B200.

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1]): 256x256 synthetic A/B ~ U(-1,1), --stn_type unet (cfg A), resnet_9blocks
generator, PatchGAN, LSGAN, --no_dropout, bf16 storage / fp32 accumulate, 16 paired samples per GPU (weak
scaling).  One "step" = forward + D step + T/R step, including both gradient all-reduces and both Adam updates.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HW = 256
PER_GPU_BATCH = 16
CONV_GFLOP_PER_SAMPLE = 724.72     # algorithmic conv FLOPs (2*MAC, fwd+bwd) per paired sample
WORKLOAD = "C2: 256x256 synthetic A/B, unet STN cfg A + resnet_9blocks + PatchGAN, lsgan, no_dropout, batch 16/GPU"
# BASELINE.json configs: name -> (size, per-GPU batch, extra flags, conv GFLOP per paired sample, label)
WORKLOADS = {
    "C2": dict(size=256, batch=16, multi_resolution=1, lambda_smooth=0.0, alpha=0.0, multires_reg=1, gflop=724.72, label=WORKLOAD),
    "C4": dict(size=512, batch=8, multi_resolution=3, lambda_smooth=200.0, alpha=1.0, multires_reg=3, gflop=3008.06,
               label="C4: 512x512 synthetic A/B, unet STN + resnet_9blocks, 3-scale PatchGAN (--multi_resolution 3), bilateral "
                     "smoothness (--lambda_smooth 200 --stn_bilateral_alpha 1.0 --stn_multires_reg 3), lsgan, no_dropout, batch 8/GPU"),
    "C5": dict(size=1024, batch=4, multi_resolution=1, lambda_smooth=0.0, alpha=0.0, multires_reg=1, gflop=11636.88,
               label="C5: 1024x1024 synthetic A/B, unet STN (dense deformation field) + resnet_9blocks + PatchGAN, lsgan, "
                     "no_dropout, batch 4/GPU"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_engine_model(per_gpu_batch, size=HW, precision="bf16", conv_engine="auto", netG="resnet_9blocks",
                       stn_type="unet", extra=()):
    from nemar_b200.engine import functional as F
    from nemar_b200.models import create_model
    from nemar_b200.options.train_options import TrainOptions
    argv = ["--dataroot", "none", "--name", "bench", "--checkpoints_dir", "/tmp/nemar_b200_bench", "--gpu_ids",
            str(int(os.environ.get("LOCAL_RANK", "0"))), "--gan_mode", "lsgan", "--no_dropout", "--stn_type", stn_type,
            "--netG", netG, "--img_height", str(size), "--img_width", str(size), "--batch_size", str(per_gpu_batch),
            "--dataset_mode", "synthetic", "--precision", precision, "--conv_engine", conv_engine] + list(extra)
    opt = TrainOptions().parse(argv, quiet=True)
    torch.manual_seed(0)          # identical initial replicas on every rank
    model = create_model(opt)
    F.bump_weights_epoch()
    return model, opt


def run_engine(args):
    from nemar_b200.engine import lib as L
    from nemar_b200.engine import parallel
    import torch.distributed as dist
    world, rank, local = parallel.init_process_group_from_env()
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    extra = []
    if args.multi_resolution > 1:
        extra += ["--multi_resolution", str(args.multi_resolution)]
    if args.lambda_smooth > 0:
        extra += ["--lambda_smooth", str(args.lambda_smooth), "--stn_bilateral_alpha", str(args.alpha), "--stn_multires_reg",
                  str(args.multires_reg)]
    if args.cuda_graph:
        extra += ["--cuda_graph", "1"]
    if args.batch_d >= 0:
        extra += ["--batch_d", str(args.batch_d)]
    if args.stream_overlap >= 0:
        extra += ["--stream_overlap", str(args.stream_overlap)]
    model, opt = build_engine_model(args.batch, args.size, args.precision, args.conv_engine, extra=extra)
    # the global batch is drawn once (seed 1) and sliced by rank so that 1-GPU and N-GPU runs see the same data
    g = torch.Generator().manual_seed(1)
    A_all = torch.rand((args.batch * world, 3, args.size, args.size), generator=g) * 2 - 1
    B_all = torch.rand((args.batch * world, 3, args.size, args.size), generator=g) * 2 - 1
    A_host = parallel.shard_batch(A_all, rank, world).contiguous().pin_memory()
    B_host = parallel.shard_batch(B_all, rank, world).contiguous().pin_memory()
    A_dev, B_dev = A_host.to(dev), B_host.to(dev)
    dev_batch = {"A": A_dev, "B": B_dev, "A_paths": "", "B_paths": ""}
    host_batch = {"A": A_host, "B": B_host, "A_paths": "", "B_paths": ""}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from nemar_b200.data.prefetch import DevicePrefetcher

    def timed(batch, steps, read_loss):
        # e2e (read_loss): host batches go through the product's own input pipeline — every step's inputs are copied
        # from pinned host memory inside the timed region (side stream, overlapped with the previous step)
        feed = [batch] * steps
        if read_loss:
            feed = DevicePrefetcher([batch] * 2, dev)
            for data in feed:                                    # untimed: first-use allocations of the input pipeline
                model.set_input(data)
                model.optimize_parameters()
            feed.loader = [batch] * steps                        # same staging buffers for the timed pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host_loss = torch.zeros(steps, dtype=torch.float32).pin_memory() if read_loss else None
        for i, data in enumerate(feed):
            model.set_input(data)
            model.optimize_parameters()
            if read_loss:       # device -> host read of every step's result: async D2H into pinned memory, consumed after the loop
                host_loss[i:i + 1].copy_(model.loss_D.detach().reshape(1), non_blocking=True)
        e1.record()
        barrier()
        if read_loss:
            assert bool(torch.isfinite(host_loss).all()), "a step's loss did not reach the host"
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step():
        model.set_input(dev_batch)
        model.optimize_parameters()

    graph_wanted = bool(args.cuda_graph) and not args.profile
    if not graph_wanted:
        opt.cuda_graph = 0
    # graph mode: three eager steps on the capture stream, the capture (+ first replay), then replays
    n_warm = args.warmup if args.profile else max(args.warmup, 3) + (4 if graph_wanted else 0)
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.COUNTERS["launches"] = 0
    graphed = graph_wanted and getattr(model, "_graph_state", {}).get("graph") is not None
    L.TIMER.enable(0 if graphed else args.kernel_timing)
    ms = timed(dev_batch, args.steps, read_loss=False)
    launches = L.COUNTERS["launches"]
    kstats = L.TIMER.collect()
    L.TIMER.enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = ms if args.profile else timed(host_batch, args.steps, read_loss=True)      # (profiler runs: no second pass)
    roofline_pass = "the timed region itself (CUDA events around every conv launch on the launching stream)"
    if graphed and args.kernel_timing and world == 1:
        # a graph replay has no place for per-kernel events: the roofline figures come from the SAME K steps launched
        # eagerly right after the timed (replayed) region — same kernels, same shapes, same stream.  (Only at N = 1: the
        # per-kernel figures do not depend on N, and a multi-rank run stays "eager warm-up, capture, replays only".)
        from nemar_b200.engine.config import CONFIG as ENGINE_CONFIG
        saved = (ENGINE_CONFIG.wgrad_stream, getattr(opt, "stream_overlap", 1))
        try:
            opt.cuda_graph = 0
            # per-kernel events must bracket kernels that run ALONE: no second / third stream in this pass
            ENGINE_CONFIG.wgrad_stream, opt.stream_overlap = False, 0
            step()
            L.TIMER.enable(args.kernel_timing)
            timed(dev_batch, args.steps, read_loss=False)
            kstats = L.TIMER.collect()
            roofline_pass = ("a separate eager, single-stream pass of the same %d steps right after the timed region (the timed "
                             "region replays a CUDA graph with three streams, which has no place for per-kernel events)" % args.steps)
        except Exception as e:      # noqa: BLE001 - the headline numbers above are already measured; never lose them
            sys.stderr.write("roofline pass failed: %r\n" % (e,))
            kstats = {}
        finally:
            L.TIMER.enable(False)
            opt.cuda_graph = 1
            ENGINE_CONFIG.wgrad_stream, opt.stream_overlap = saved

    global_batch = args.batch * world
    value = global_batch * args.steps / (ms / 1e3)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            if graphed:
                dist.barrier()          # rank 0 prints its line before any rank leaves (no peer dies under a live rank)
                torch.cuda.synchronize()
            leave_process_group(graphed)
        return
    # ---- roofline of the dominant kernel (largest share of timed kernel time)
    roof = None
    ncu_traffic, traffic_note = dominant_kernel_traffic()
    if kstats:
        name, st = max(kstats.items(), key=lambda kv: kv[1]["ms"])
        tf = st["flops"] / (st["ms"] / 1e3) / 1e12 if st["ms"] > 0 else 0.0
        tot_ms = sum(v["ms"] for v in kstats.values())
        roof = {"bound": "tensor", "kernel": name, "achieved": round(tf, 2), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(tf / peaks["tf_sustained"], 4), "traffic": ncu_traffic,
                "traffic_note": traffic_note, "peak_source": peaks["source"] + ", sustained",
                "timed_in": roofline_pass, "launches": st["n"], "avg_launch_ms": round(st["ms"] / max(st["n"], 1), 4),
                "share_of_timed_kernels": round(st["ms"] / max(tot_ms, 1e-9), 4),
                "kernel_note": "records are keyed on the kernel instance the library launched (nemar_last_conv_kernel); "
                               "`top` lists the layer geometries it served, prefixed by the pass",
                "by_kernel": {k: {"ms": round(v["ms"], 3), "n": v["n"],
                                  "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                                  "passes": {op: {"n": o["n"], "avg_us": round(1e3 * o["ms"] / max(o["n"], 1), 1),
                                                  "tflops": round(o["flops"] / max(o["ms"], 1e-9) / 1e9, 1)} for op, o in v.get("ops", {}).items()},
                                  "top": dict(sorted(((a, round(b, 3)) for a, b in v["top"].items()), key=lambda t: -t[1])[:args.top])}
                              for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])}}
    if roof is None and graphed and world > 1:
        roof = {"bound": "tensor", "achieved": None, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": None, "traffic": ncu_traffic,
                "note": "per-kernel figures are timed at N = 1 (eager pass after the replayed region); a multi-rank run replays "
                        "the captured step only — see conv_roofline_frac_whole_step for the whole-step figure at this N"}
    gflop = args.gflop
    conv_frac = value * gflop * 1e9 / (world * peaks["tf_sustained"] * 1e12) if gflop else None
    out = {"metric": "paired %dx%d samples/sec" % (args.size, args.size), "value": round(value, 3), "unit": "samples/s", "n_gpus": world,
           "steps": args.steps, "warmup": n_warm, "ms_per_step": round(ms / args.steps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
           "config": {"workload": args.label if args.label else
                      "%dx%d unet STN + resnet_9blocks, batch %d/GPU, multi_resolution %d, lambda_smooth %g alpha %g" % (
                          args.size, args.size, args.batch, args.multi_resolution, args.lambda_smooth, args.alpha),
                      "warmup_note": "%d warm-up steps%s" % (n_warm, " (3 eager + the capture step + replays)" if graph_wanted else ""),
                      "conv_gflop_per_sample": gflop,
                      "global_batch": global_batch, "parallelism": "dp%d" % world, "conv_engine": args.conv_engine,
                      "l2": "inputs larger than L2: a step streams several GB of activations, no flush needed",
                      "cuda_graph": graphed, "batch_d": int(getattr(opt, "batch_d", 1)),
                      "allreduce_per_step": 2},
           "clocks": clocks,
           "e2e": {"value": round(e2e_value, 3), "unit": "samples/s", "h2d_bytes_per_step": int(A_host.numel() * 4 * 2),
                   "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
           "gpu_launches": int(launches),
           "conv_roofline_frac_whole_step": round(conv_frac, 4) if conv_frac is not None else None,
           "roofline": roof}
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_baseline(args, budget_s=15.0)
        except Exception as e:      # noqa: BLE001
            out["cpu_baseline"] = {"error": repr(e)}
    if world == 1 and args.torch_gpu_reference:
        try:
            out["torch_gpu_reference"] = torch_gpu_reference(args, dev)
        except Exception as e:      # noqa: BLE001
            out["torch_gpu_reference"] = {"error": repr(e)[:300]}
    if args.grid_sample_bench:
        try:
            out["grid_sample"] = grid_sample_bench(dev, peaks)
        except Exception as e:      # noqa: BLE001
            out["grid_sample"] = {"error": repr(e)}
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1 and graphed:
        dist.barrier()
        torch.cuda.synchronize()
        leave_process_group(graphed)


def dominant_kernel_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the dominant conv instance (256->256 k3) from the committed
    `ncu --set full` summary — only when that capture was taken on the conv-engine sources this library was built from
    (digest of conv_tc.cu and its headers, written at build time by nemar_b200/build.py); otherwise null."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_ncu.json")))
        built = open(os.path.join(ROOT, "nemar_b200", "build", "stamp_conv_tc")).read().strip()
        if prof.get("conv_tc_digest") != built:
            return None, "no ncu capture of this build's conv engine (profiles/dominant_kernel_ncu.json was taken on digest %s...)" % str(prof.get("conv_tc_digest"))[:12]
        pl = prof["per_launch"]
        return (round((pl["dram__bytes_read_MB"] + pl["dram__bytes_write_MB"]) * 1e6),
                "dram bytes/launch of the 256->256 k3 fprop instance from %s (algorithmic 70.3 MB: the bf16 output stays in "
                "the 126 MB L2 and is consumed from there)" % prof.get("file", "profiles/dominant_kernel_ncu.json"))
    except Exception:
        return None, "no ncu capture of this build committed"


def leave_process_group(graphed):
    """End of a multi-rank run.  After a captured step (NCCL all-reduces inside the CUDA graph) destroy_process_group()
    does not return (measured at 2 ranks: the communicator teardown waits forever while the graph is alive), so a rank
    that replayed a graph leaves with os._exit once every rank has passed the final barrier; nothing is pending then."""
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    if graphed:
        os._exit(0)
    dist.destroy_process_group()


def pick_cpu_threads():
    """Host threads for the CPU arm — ONE deterministic rule: min(cores this process may use, 32).  The oracle's ATen /
    oneDNN convolutions at batch 2 stop scaling there on the GPU boxes (16 threads 2.2 samples/s, 32 threads 3.6,
    128 threads collapse); round 1's timing-based calibration flipped between 16 and 32 from run to run."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    n = max(1, min(avail, 32))
    torch.set_num_threads(n)
    return n


def oracle_cfg(args):
    from oracle import nemar_oracle as O
    return O.OracleConfig(stn_type="unet", n_blocks=9, height=args.size, width=args.size, lambda_smooth=args.lambda_smooth,
                          alpha=args.alpha, multires_reg=args.multires_reg, multi_resolution=args.multi_resolution)


def cpu_baseline(args, budget_s, batch=2, n_blocks=9):
    """The oracle port of the reference's CPU path, timed on this box's host cores on a bounded sample."""
    from oracle import nemar_oracle as O
    pick_cpu_threads()
    size = args.size
    batch = 1 if size >= 1024 else batch
    cfg = oracle_cfg(args)
    T, R, Ds = O.make_states(cfg, seed=0, live_head=False)
    A, B = O.synthetic_batch(batch, size, size, seed=1)
    st = O.OracleStep(cfg, T, R, Ds)
    st.step(A, B)                                   # warm-up
    t0, n = time.time(), 0
    while n < 1 or (time.time() - t0 < budget_s and n < 50):
        st.step(A, B)
        n += 1
    dt = time.time() - t0
    return {"value": round(batch * n / dt, 4), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d optimize_parameters steps of the %dx%d unet/resnet_%d workload at batch %d, fp32, torch CPU "
                      "(the reference's own ATen path restated in oracle/nemar_oracle.py)" % (n, size, size, n_blocks, batch)}


def torch_gpu_reference(args, dev, steps=5):
    """Informational: the reference's algorithm as eager PyTorch on the SAME B200 — fp32
    storage, ATen / cuDNN sm_100 kernels, what a NeMAR user gets from `--gpu_ids 0` today.  /root/reference does not
    exist on the GPU box, so this runs the golden-pinned restatement (oracle/nemar_oracle.py: the same torch ops on
    state dicts) with every tensor on the device; same workload and batch as the engine line."""
    from collections import OrderedDict
    from oracle import nemar_oracle as O
    cfg = O.OracleConfig(stn_type="unet", n_blocks=9, height=args.size, width=args.size, lambda_smooth=args.lambda_smooth,
                         alpha=args.alpha, multires_reg=args.multires_reg, multi_resolution=args.multi_resolution)
    T, R, Ds = O.make_states(cfg, seed=0, live_head=False)
    mv = lambda sd: OrderedDict((k, v.to(dev)) for k, v in sd.items())
    res = {}
    A, B = O.synthetic_batch(args.batch, args.size, args.size, seed=1)
    A, B = A.to(dev), B.to(dev)
    old_default = torch.get_default_device() if hasattr(torch, "get_default_device") else None
    torch.set_default_device(dev)      # the oracle builds its identity grids / targets with the default device
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True           # the reference sets it (models/base_model.py:39)
        st = O.OracleStep(cfg, mv(T), mv(R), [mv(d) for d in Ds])
        for _ in range(3):
            st.step(A, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            st.step(A, B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[name] = {"value": round(args.batch / (ms / 1e3), 2), "unit": "samples/s", "ms_per_step": round(ms, 2)}
        del st
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    torch.set_default_device(old_default if old_default is not None else "cpu")
    res["kind"] = "port"
    res["what"] = ("the reference's optimize_parameters restated in torch (oracle/nemar_oracle.py, golden-pinned to the reference) "
                   "run eagerly on cuda:0, batch %d: ATen/cuDNN kernels, fp32 storage; 'tf32' = same with TF32 tensor-core math "
                   "allowed.  Each step ends with the loss read-back the reference's logging does" % args.batch)
    return res


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), all host threads.
    Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import nemar_oracle as O
    pick_cpu_threads()
    batch = 1 if args.size >= 1024 else 2
    cfg = oracle_cfg(args)
    T, R, Ds = O.make_states(cfg, seed=0, live_head=False)
    A, B = O.synthetic_batch(batch, args.size, args.size, seed=1)
    st = O.OracleStep(cfg, T, R, Ds)
    for _ in range(max(1, min(args.warmup, 2))):
        st.step(A, B)
    steps = max(1, min(args.steps, 10 if args.size <= 256 else 3))
    t0 = time.time()
    for _ in range(steps):
        st.step(A, B)
    dt = time.time() - t0
    value = batch * steps / dt
    sample = "each step = one optimize_parameters at batch %d (bounded sample of the batch-%d workload; CPU samples/s is " \
             "batch-insensitive), fp32, %d threads" % (batch, args.batch, torch.get_num_threads())
    print(json.dumps({"impl": "reference", "metric": "paired %dx%d samples/sec" % (args.size, args.size), "value": round(value, 4), "unit": "samples/s",
                      "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)),
                      "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                      "config": {"workload": args.label or "%dx%d" % (args.size, args.size), "global_batch": batch, "parallelism": "cpu"},
                      "cpu_baseline": {"value": round(value, 4), "unit": "samples/s", "cores": torch.get_num_threads(),
                                       "kind": "port", "sample": sample},
                      "e2e": {"value": round(value, 4), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def grid_sample_bench(dev, peaks, size=1024, n=4, iters=20):
    """grid_sample HBM GB/s (second half of BASELINE.json's metric): 1024x1024, C=3, fp32, near-identity grid;
    algorithmic bytes per output pixel from SURVEY.md section 8d (fwd 32 B; bwd with grad_input 52 B)."""
    from nemar_b200.engine import functional as F
    g = torch.Generator().manual_seed(0)
    img = (torch.rand((n, 3, size, size), generator=g) * 2 - 1).to(dev)
    xs = torch.linspace(-1, 1, size)
    ident = torch.stack([xs.view(1, size).expand(size, size), xs.view(size, 1).expand(size, size)], -1)
    grid = (ident.unsqueeze(0) + torch.randn((n, size, size, 2), generator=g) * (4.0 / size)).to(dev).contiguous()
    dout = torch.randn((n, 3, size, size), generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    px = n * size * size
    res = {}
    # every buffer is allocated up front and the kernels are launched through the C ABI directly, right behind the
    # L2-flush memset, so that no host-side launch latency sits inside the event pair
    out = torch.empty_like(img)
    dimg = torch.zeros_like(img)
    dgrid = torch.empty_like(grid)
    for name in ("fwd", "bwd"):
        times = []
        for it in range(iters + 3):
            if name == "bwd":
                dimg.zero_()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "fwd":
                F.call("nemar_grid_sample_fwd", F.fptr(img), None, 1, n, 3, size, size, F.fptr(grid), size, size,
                       F.fptr(out), None, None, F.stream())
            else:
                F.call("nemar_grid_sample_bwd", F.fptr(img), None, 1, n, 3, size, size, F.fptr(grid), size, size,
                       F.fptr(dout), None, F.fptr(dimg), None, F.fptr(dgrid), F.stream())
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        bpp = 32 if name == "fwd" else 52
        gbs = px * bpp / (ms / 1e3) / 1e9
        res[name] = {"gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 4), "ms": round(ms, 4),
                     "bytes_per_px": bpp}
    res["config"] = "%dx%d, batch %d, C=3 fp32, grid = identity + N(0, 2 px); L2 flushed between launches" % (size, size, n)
    res["peak_gbs"] = peaks["hbm"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="C2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration: C2 (256^2, the metric's config; default), C4 (512^2, 3-scale D, bilateral), "
                         "C5 (1024^2); --batch/--size/... override its fields")
    ap.add_argument("--batch", type=int, default=None, help="paired samples per GPU")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--conv_engine", type=str, default="auto", choices=["auto", "generic"])
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--multi_resolution", type=int, default=None, help="discriminator scales (C4: 3)")
    ap.add_argument("--lambda_smooth", type=float, default=None, help="STN regulariser weight (C4: 200)")
    ap.add_argument("--alpha", type=float, default=None, help="bilateral alpha of the smoothness term (C4: 1.0)")
    ap.add_argument("--multires_reg", type=int, default=None)
    ap.add_argument("--torch_gpu_reference", type=int, default=1,
                    help="1: also time the reference's algorithm as eager fp32 PyTorch on cuda:0 (ATen/cuDNN kernels) — "
                         "the informal 'reference on Blackwell' bar of BASELINE.md section 3 item 5 (N = 1 only)")
    ap.add_argument("--batch_d", type=int, default=-1, help="-1: the engine's default; 0/1: one discriminator pass per (A, B) pair / per phase")
    ap.add_argument("--stream_overlap", type=int, default=-1, help="-1: the engine's default; 0/1: STN regressor on a second stream")
    ap.add_argument("--cuda_graph", type=int, default=1,
                    help="1 (default): the model captures optimize_parameters in a CUDA graph (its --cuda_graph 1 flag) and the "
                         "timed region replays it; the roofline figures then come from an eager pass of the same steps")
    ap.add_argument("--top", type=int, default=4, help="shapes listed per kernel class in roofline.by_kernel")
    ap.add_argument("--kernel_timing", type=int, default=1, help="time conv launches with CUDA events (roofline)")
    ap.add_argument("--grid_sample_bench", type=int, default=1)
    ap.add_argument("--profile", action="store_true", help="for ncu runs: honour --warmup below 3 (numbers printed under a "
                                                           "profiler are never bench values)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    overridden = False
    for k in ("size", "batch", "multi_resolution", "lambda_smooth", "alpha", "multires_reg"):
        if getattr(args, k) is None:
            setattr(args, k, wl[k])
        elif getattr(args, k) != wl[k]:
            overridden = True
    args.label = None if overridden else wl["label"]
    args.gflop = wl["gflop"] if (args.size == wl["size"] and args.multi_resolution == wl["multi_resolution"]) else None
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
